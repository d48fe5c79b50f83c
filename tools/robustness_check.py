"""Krylov eigensolver on large problems other than the benchmark torus: does the cut estimation hold, does it converge without
retries / restarts, how long does it take?  usage: python tools/robustness_check.py [kind ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import RVGP
from tests.workloads import make_cloud
cases = [("sphere", 1000000, 500, {}), ("manifold5_R32", 200000, 300, dict(n_neighbors=22, explained_variance=0.9)),
         ("torus", 300000, 256, {}), ("moebius", 400000, 200, {})]
only = set(sys.argv[1:])                    # optional: restrict to these cloud kinds
for kind, n, k, kw in cases:
    if only and kind not in only:
        continue
    X = make_cloud(kind, n, 0)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d = RVGP.create_data_object(X, n_eigenpairs=k, verbose=False, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out = []
    for name in ("eig_L", "eig_Lc"):
        st = d.stats[name]
        out.append({a: st.get(a) for a in ("solver", "paired", "degree", "blocks", "restarts", "checks", "cut", "lam_k_est", "cut_retry",
                                           "final_rr_outer", "residual_max", "tol_abs", "converged", "spmm_kernel")})
    ev_L, ev_Lc = d.evals_L, d.evals_Lc
    print("%s n=%d k=%d dim_man=%d: %.2f s  stages %s" % (kind, n, k, d.dim_man, dt, {a: round(b, 3) for a, b in d.timings.items()}))
    print("   true lambda_k(L) %.5g (estimate %.5g)   lambda_k(Lc) %.5g (estimate %.5g)" % (ev_L[-1], out[0]["lam_k_est"] or -1, ev_Lc[-1], out[1]["lam_k_est"] or -1))
    for o in out:
        print("  ", o)
    del d
