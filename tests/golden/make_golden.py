"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build
container.  Run once here (`python tests/golden/make_golden.py`); the fixtures are committed because
/root/reference does not exist on the GPU box.

Requires oracle/_ref (python oracle/build_ref.py --instrumented).  The instrumented build only adds
an export of the popped geodesic-neighbourhood index sequence (see oracle/build_ref.py).
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from tests.workloads import make_cloud  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (cloud kind, n, kwargs for create_data_object, extras)
    "sphere_n2000_k50": dict(kind="sphere", n=2000, seed=0, k=50, nb=10, fps_spacing=0.05, smooth_t=100.0),
    "torus_n600_k20": dict(kind="torus", n=600, seed=0, k=20, nb=10, fps_spacing=0.1, smooth_t=10.0),
    "flat3torus_R6_n900_k24": dict(kind="flat3torus", n=900, seed=0, k=24, nb=10, fps_spacing=0.2, smooth_t=None),
    "sheet_R20_n500_k16": dict(kind="sheet_R20", n=500, seed=0, k=16, nb=14, fps_spacing=0.2, smooth_t=None),
}


def run_case(name, c, ns):
    X = make_cloud(c["kind"], c["n"], c["seed"])
    ns.ptu.GEO_SEQ.clear()
    with contextlib.redirect_stdout(io.StringIO()):
        d = ns.dataclass.data(X, n_neighbors=c["nb"], n_eigenpairs=c["k"])
    n, D = X.shape
    K = int(c["nb"] * 1.5)
    seq = np.array(ns.ptu.GEO_SEQ[:n], dtype=np.int32)
    assert seq.shape == (n, K + 1)
    # directed kNN sets exactly as the reference asks sklearn for them (geometry.py:103-110)
    from sklearn.neighbors import kneighbors_graph
    A = kneighbors_graph(X, c["nb"], mode="connectivity", metric="minkowski", p=2, include_self=False)
    knn = np.sort(A.tocsr().indices.reshape(n, c["nb"]).astype(np.int32), axis=1)
    # CSR the Cython code saw (pyx:84-103)
    import networkx as nx
    M = nx.adjacency_matrix(d.G, weight="weight")
    M = M.maximum(M.T).tocsr()
    M.sort_indices()
    # full tangent frames (d=D) and Sigma (dataclass.py:35); rerun, as the data object only keeps gauges
    with contextlib.redirect_stdout(io.StringIO()):
        tangents, Sigma = ns.ptu.tangent_frames(X, d.G, D, c["nb"] * 1.5)
    ns.ptu.GEO_SEQ.clear()
    out = dict(
        X=X, knn=knn, indptr=M.indptr.astype(np.int32), indices=M.indices.astype(np.int32),
        geo_seq=seq, Sigma=Sigma, projectors_full=np.einsum("nip,njp->nij", tangents[:, :, :d.dim_man], tangents[:, :, :d.dim_man]),
        dim_man=np.int64(d.dim_man), gauges=d.gauges,
        L_data=d.L.data, L_indices=d.L.indices.astype(np.int32), L_indptr=d.L.indptr.astype(np.int32),
        Lc_data=d.Lc.tobsr((d.dim_man, d.dim_man)).data,
        Lc_indices=d.Lc.tobsr((d.dim_man, d.dim_man)).indices.astype(np.int32),
        Lc_indptr=d.Lc.tobsr((d.dim_man, d.dim_man)).indptr.astype(np.int32),
        evals_L=d.evals_L, evecs_L=d.evecs_L, evals_Lc=d.evals_Lc, evecs_Lc=d.evecs_Lc,
    )
    perm, lambdas = ns.geometry.furthest_point_sampling(X, spacing=c["fps_spacing"])
    out["fps_perm"], out["fps_lambdas"] = perm, lambdas
    perm_n, lambdas_n = ns.geometry.furthest_point_sampling(X, N=40, start_idx=3)
    out["fps_perm_N40"], out["fps_lambdas_N40"] = perm_n, lambdas_n
    d.random_vector_field(seed=1)
    out["random_field_seed1"] = d.vectors.copy()
    if c["smooth_t"] is not None:
        d.smooth_vector_field(t=c["smooth_t"])
        out["smoothed_field"] = np.asarray(d.vectors).copy()
        out["smooth_t"] = np.float64(c["smooth_t"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "n", n, "D", D, "dim_man", d.dim_man, "nnzb", M.nnz, "evals_Lc[:3]", d.evals_Lc[:3],
          "fps", len(perm))


if __name__ == "__main__":
    ns = ref_harness.load(instrumented=True)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        run_case(name, c, ns)
