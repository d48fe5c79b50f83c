"""Scratch probe: time create_data_object (+ smoothing, fit, transform) stage by stage at a given size."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.workloads import make_cloud

def main():
    kind, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    nb = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    import RVGP
    from rvgp_b200 import eigensolver
    X = make_cloud(kind, n, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d = RVGP.create_data_object(X, n_eigenpairs=k, n_neighbors=nb, verbose=False)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    print("create_data_object %.3f s" % t_create)
    print(json.dumps({a: round(b, 4) for a, b in d.timings.items()}))
    for nm in ("eig_L", "eig_Lc"):
        st = d.stats[nm]
        print(nm, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in st.items()})
    t0 = time.perf_counter(); d.random_vector_field(seed=1); d.smooth_vector_field(t=100); torch.cuda.synchronize()
    print("random+smooth %.3f s" % (time.perf_counter() - t0), d.stats.get("smoothing"))
    np.random.seed(0)
    train_ind = np.random.choice(np.arange(n), size=n // 2)
    t0 = time.perf_counter(); gp = RVGP.fit(d, train_ind=train_ind, noise_variance=0.001); torch.cuda.synchronize()
    print("fit %.3f s  evals %d  solver %s" % (time.perf_counter() - t0, gp._gpr.n_eval, gp.solver))
    mask = np.ones(n, dtype=bool); mask[train_ind] = False
    t0 = time.perf_counter(); m, v = gp.transform(d, mask); torch.cuda.synchronize()
    print("transform %.3f s  n_test %d" % (time.perf_counter() - t0, m.shape[0]))
    print("mean abs err vs field:", float(np.linalg.norm(m - d.vectors[mask], axis=1).mean()))
    print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)

if __name__ == "__main__":
    main()
