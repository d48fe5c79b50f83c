"""Config C5 shape at reduced n: 5-manifold in R^32, n_neighbors=22, vs the oracle (CPU) on the same input."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.workloads import make_cloud
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
X = make_cloud("manifold5_R32", n, 0)
import RVGP
t0 = time.perf_counter()
ev = float(os.environ.get("C5_EXPLAINED_VARIANCE", "0.8"))   # isotropic 5-d neighbourhoods put ~80% of the variance in 4 PCs: use 0.9 at large n
d = RVGP.create_data_object(X, n_neighbors=22, n_eigenpairs=k, explained_variance=ev, verbose=False)
torch.cuda.synchronize()
print("GPU create_data_object %.2f s dim_man %d" % (time.perf_counter() - t0, d.dim_man), {a: round(b, 3) for a, b in d.timings.items()})
print("eig_Lc", {a: d.stats["eig_Lc"][a] for a in ("N", "m", "outer", "filter_launches", "t_filter", "t_dense", "t_host", "spmm_kernel", "residual_max", "converged")})
if "--oracle" in sys.argv:
    from oracle import rvgp_oracle as O
    t0 = time.perf_counter()
    o = O.create_data_object(X, n_neighbors=22, n_eigenpairs=k, explained_variance=ev)
    print("oracle %.2f s dim_man %d" % (time.perf_counter() - t0, o.dim_man), {a: round(b, 2) for a, b in o.timings.items()})
    idx_same = np.array_equal(np.sort(d._graph.knn.cpu().numpy(), 1), o.knn)
    print("knn sets equal:", idx_same, " csr equal:", np.array_equal(d._graph.indices.cpu().numpy(), o.indices))
    print("evals_Lc rel err %.2e   evals_L abs err %.2e" % (np.abs(d.evals_Lc - o.evals_Lc).max() / o.evals_Lc.max(), np.abs(d.evals_L - o.evals_L).max()))
    P = np.einsum("nip,njp->nij", d.gauges, d.gauges); Po = np.einsum("nip,njp->nij", o.gauges, o.gauges)
    print("projector err %.2e" % np.abs(P - Po).max())
