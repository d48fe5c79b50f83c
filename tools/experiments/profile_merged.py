import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.profile_spmm import build
from rvgp_b200._cabi import get_handle
A, L, _ = build("torus", 1000000)
h = get_handle(0)
dev = A.indptr.device
def bench(M, X, W, Y, reps=10):
    kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W)
    for _ in range(3): M.spmm(X, Y, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): M.spmm(X, Y, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, M in (("Lc", A), ("L", L)):
    for b in (64,):
        X = torch.randn((M.nrows, b), dtype=torch.float64, device=dev); W = torch.randn_like(X)
        Y0 = torch.empty_like(X); Y1 = torch.empty_like(X)
        M.merged = None; M.d_code = M.d
        t0 = bench(M, X, W, Y0)
        by = M.spmm_bytes(b, True)
        print("%s b=%d v2 plain: %.4f ms frac %.3f" % (name, b, t0, by / t0 / 1e6 / 6534.5))
        for R in (4,):
            mp = M.enable_merged(R)
            for rot in ((False, True) if name == "Lc" else (False,)):
                M.d_code = M.d
                if rot: M.compress_rot2()
                for lpr in (0,):
                    h.set_option("spmm_lpr", lpr)
                    t1 = bench(M, X, W, Y1)
                    print("%s b=%d merged R=%d rot2=%s lpr=%d reuse %.2f: %.4f ms frac %.3f maxdiff %.1e" % (name, b, R, rot, lpr, mp["reuse"], t1, by / t1 / 1e6 / 6534.5, float((Y0 - Y1).abs().max())))
                h.set_option("spmm_lpr", 0)
