"""GPU parity tests of the two boundaries that round 1 left untested:

  * the inner FFI-shaped shim ``ptu_dijkstra.tangent_frames`` / ``ptu_dijkstra.connections`` (reference
    RVGP/lib/ptu_dijkstra.pyx:33, :128) called exactly as RVGP/dataclass.py:35,47 calls them, with a networkx graph;
  * the public frame contractions ``project_to_manifold`` / ``express_in_local_frame`` (RVGP/geometry.py:165-176);
  * config C5's shape (5-manifold in R^32, d = 5 blocks, brute-force kNN because D > 15) against the oracle at reduced n.
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from tests.conftest import load_golden, GOLDEN_NB, subspace_angle_max, eigen_clusters

pytestmark = pytest.mark.gpu
CASES = ["sphere_n2000_k50", "torus_n600_k20", "flat3torus_R6_n900_k24", "sheet_R20_n500_k16"]


def _nx_graph(g):
    """The graph object the reference hands to the extension: nx.from_scipy_sparse_array(A + I) (geometry.py:111-112)."""
    import networkx as nx
    n = g["X"].shape[0]
    A = sp.csr_matrix((np.ones(g["indices"].size), g["indices"], g["indptr"]), shape=(n, n))
    return nx.from_scipy_sparse_array(A)


def _ref_R_blocks(g):
    """R_ij in CSR entry order recovered from the golden Lc blocks: (i,i) = deg_i R_ii, (i,j) = -R_ij (geometry.py:38-42)."""
    ip, ix = g["indptr"], g["indices"]
    rows = np.repeat(np.arange(ip.size - 1), np.diff(ip))
    deg = (np.diff(ip) - 1).astype(np.float64)
    scale = np.where(rows == ix, deg[rows], -1.0)
    return g["Lc_data"] / scale[:, None, None]


@pytest.mark.parametrize("case", CASES)
def test_ptu_dijkstra_shim_matches_reference(case):
    import ptu_dijkstra                       # the repo-root drop-in for the reference's CPython extension
    assert os.path.dirname(os.path.abspath(ptu_dijkstra.__file__)) == os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = load_golden(case)
    X = g["X"]
    n, D = X.shape
    G = _nx_graph(g)
    d = int(g["dim_man"])
    tangents, Sigma = ptu_dijkstra.tangent_frames(X, G, D, GOLDEN_NB[case] * 1.5)      # dataclass.py:35: float K, d = D
    assert tangents.shape == (n, D, D) and Sigma.shape == (n, D) and tangents.dtype == np.float64
    np.testing.assert_allclose(Sigma, g["Sigma"], rtol=1e-12, atol=1e-14)
    P = np.einsum("nip,njp->nij", tangents[:, :, :d], tangents[:, :, :d])
    np.testing.assert_allclose(P, g["projectors_full"], atol=1e-10)
    R = ptu_dijkstra.connections(g["gauges"], G, d)                                     # dataclass.py:47
    assert sp.isspmatrix_coo(R) and R.shape == (n * d, n * d)
    Rb = sp.bsr_matrix(R.tocsr(), blocksize=(d, d))
    Rb.sort_indices()
    assert np.array_equal(Rb.indices, g["indices"]) and np.array_equal(Rb.indptr, g["indptr"])
    np.testing.assert_allclose(Rb.data, _ref_R_blocks(g), atol=1e-12)


def test_ptu_dijkstra_shim_errors_and_weighted_graph():
    import networkx as nx
    import ptu_dijkstra
    g = load_golden("torus_n600_k20")
    X, G = g["X"], _nx_graph(g)
    n, D = X.shape
    with pytest.raises(ValueError, match="less than the total number of samples"):       # pyx:68-72
        ptu_dijkstra.tangent_frames(X, G, 2, n)
    with pytest.raises(ValueError, match="larger or equal to the embedding dimension"):  # pyx:73-77
        ptu_dijkstra.tangent_frames(X, G, 3, 2)
    with pytest.raises(ValueError, match="less or equal to the ambient dimension"):      # pyx:78-82
        ptu_dijkstra.tangent_frames(X, G, 4, 15)
    with pytest.raises(ValueError, match="less or equal to the ambient dimension"):      # pyx:156-160
        ptu_dijkstra.connections(g["gauges"], G, 4)
    Xflat = X.copy()
    Xflat[:, 2] = 0.0                                                                     # planar cloud: rank 2 < d = 3
    with pytest.raises(RuntimeError, match="does not span"):                              # pyx:119-123
        ptu_dijkstra.tangent_frames(Xflat, G, 3, 15)
    W = nx.Graph()
    W.add_weighted_edges_from([(0, 1, 0.5), (1, 2, 2.0), (2, 0, 1.0)])
    with pytest.raises(NotImplementedError, match="non-unit edge weights"):               # refused, not silently unit-weighted
        ptu_dijkstra.tangent_frames(np.eye(3), W, 1, 1)


@pytest.mark.parametrize("case", CASES)
def test_frame_contractions_public_api(case):
    """project_to_manifold / express_in_local_frame (geometry.py:165-176) against the reference's einsum definitions and,
    through random_vector_field's recipe (dataclass.py:91-103), against the golden field of the unmodified reference."""
    import RVGP
    from RVGP.geometry import project_to_manifold, express_in_local_frame
    assert RVGP.geometry.project_to_manifold is project_to_manifold
    g = load_golden(case)
    gauges = g["gauges"]
    n, D, d = gauges.shape
    rng = np.random.default_rng(7)
    x = rng.normal(size=(n, D))
    xl = rng.normal(size=(n, d))
    loc = express_in_local_frame(x, gauges)
    assert isinstance(loc, np.ndarray) and loc.shape == (n, d) and loc.dtype == np.float64
    np.testing.assert_allclose(loc, np.einsum("bij,bi->bj", gauges, x), rtol=0, atol=1e-14 * np.abs(x).max() * D)
    amb = express_in_local_frame(xl, gauges, reverse=True)
    assert amb.shape == (n, D)
    np.testing.assert_allclose(amb, np.einsum("bji,bi->bj", gauges, xl), rtol=0, atol=1e-14 * d * np.abs(xl).max())
    proj = project_to_manifold(x, gauges)
    coeffs = np.einsum("bij,bi->bj", gauges, x)
    np.testing.assert_allclose(proj, np.einsum("bj,bij->bi", coeffs, gauges), rtol=0, atol=1e-13)
    np.testing.assert_allclose(project_to_manifold(proj, gauges), proj, atol=1e-13)      # idempotent
    np.random.seed(1)
    v = np.random.uniform(low=-0.5, high=0.5, size=(n, D))
    v = project_to_manifold(v, gauges)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    np.testing.assert_allclose(v, g["random_field_seed1"], atol=1e-12)                   # golden of the unmodified reference


def test_c5_shape_matches_oracle():
    """Config C5 at reduced n (the oracle needs seconds): 5-manifold in R^32, n_neighbors = 22, d = 5 blocks."""
    import RVGP
    from oracle import rvgp_oracle as O
    from tests.workloads import make_cloud
    n, k = 6000, 40
    X = make_cloud("manifold5_R32", n, 0)
    d = RVGP.create_data_object(X, n_neighbors=22, n_eigenpairs=k, verbose=False)
    o = O.create_data_object(X, n_neighbors=22, n_eigenpairs=k)
    assert d.dim_man == o.dim_man == 5
    assert np.array_equal(np.sort(d._graph.knn.cpu().numpy(), 1), o.knn)                 # brute-force path (D > 15): same sets
    assert np.array_equal(d._graph.indices.cpu().numpy(), o.indices)
    P = np.einsum("nip,njp->nij", d.gauges, d.gauges)
    Po = np.einsum("nip,njp->nij", o.gauges, o.gauges)
    assert np.abs(P - Po).max() < 1e-10
    np.testing.assert_allclose(d.evals_Lc, o.evals_Lc, rtol=1e-8, atol=1e-10)
    assert np.abs(d.evals_L - o.evals_L).max() < 1e-8 * o.evals_L.max()
    for name in ("L", "Lc"):
        ev = getattr(o, "evals_" + name)
        Ur, Uo = getattr(o, "evecs_" + name), getattr(d, "evecs_" + name)
        cl = eigen_clusters(ev, rtol=1e-6)
        # the last cluster may be cut by k: only complete clusters define a subspace
        for s in cl[:-1]:
            assert subspace_angle_max(Uo[:, s], Ur[:, s]) < 1e-6, (name, s)
