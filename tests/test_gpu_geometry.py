"""GPU parity tests (through the C-ABI) for the geometry front end: K1 FPS, K2 kNN, K3 CSR, K4 heap-Dijkstra,
K5 gauges, K6 dimension, K7/K8 connections + Lc, and the full data object against the golden vectors of the
unmodified reference."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests.conftest import load_golden, subspace_angle_max, eigen_clusters, GOLDEN_NB

pytestmark = pytest.mark.gpu
CASES = ["sphere_n2000_k50", "torus_n600_k20", "flat3torus_R6_n900_k24", "sheet_R20_n500_k16"]


def _dev():
    return torch.device("cuda", 0)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(_dev())


@pytest.mark.parametrize("case", CASES)
def test_knn_bit_exact(case):
    from rvgp_b200 import geometry as geo
    from oracle import rvgp_oracle as O
    g = load_golden(case)
    nb = GOLDEN_NB[case]
    idx, d2 = geo.knn_device(_t(g["X"]), nb, return_d2=True)
    idx = idx.cpu().numpy()
    assert idx.dtype == np.int32
    assert np.array_equal(np.sort(idx, 1), g["knn"])                 # sets == sklearn (reference call)
    assert np.array_equal(idx, O.knn_exact(g["X"], nb))              # order == oracle restatement, bit exact
    d2 = d2.cpu().numpy()
    ref = ((g["X"][:, None, :] - g["X"][idx]) ** 2)
    acc = np.zeros(idx.shape)
    for j in range(g["X"].shape[1]):
        acc = acc + ref[:, :, j]
    assert np.array_equal(d2, acc)                                    # squared distances bit exact (no FMA)


def test_knn_ragged_and_sharded():
    from rvgp_b200 import geometry as geo
    from oracle import rvgp_oracle as O
    rng = np.random.default_rng(5)
    X = rng.normal(size=(1237, 5))
    X[100] = X[7]                      # exact duplicate -> distance 0 tie handled by index
    X[200] = X[7]
    full = geo.knn_device(_t(X), 12).cpu().numpy()
    assert np.array_equal(full, O.knn_exact(X, 12))
    a = geo.knn_device(_t(X), 12, q_begin=0, q_count=600).cpu().numpy()
    b = geo.knn_device(_t(X), 12, q_begin=600, q_count=637).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b]), full)               # row sharding does not change results
    with pytest.raises(ValueError):
        geo.knn_device(_t(X[:10]), 10)
    k22 = geo.knn_device(_t(X), 22).cpu().numpy()
    assert np.array_equal(k22, O.knn_exact(X, 22))


@pytest.mark.parametrize("case", CASES)
def test_csr_and_heap_sequences_bit_exact(case):
    from rvgp_b200 import geometry as geo
    g = load_golden(case)
    nb = GOLDEN_NB[case]
    indptr, indices = geo.knn_to_csr_device(_t(g["knn"]))
    assert np.array_equal(indptr.cpu().numpy(), g["indptr"])
    assert np.array_equal(indices.cpu().numpy(), g["indices"])
    seq, counts = geo.geodesic_neighbourhoods_device(indptr, indices, int(nb * 1.5))
    assert int(counts.min()) == int(nb * 1.5) + 1
    assert np.array_equal(seq.cpu().numpy(), g["geo_seq"])            # pop ORDER equals the reference heap's


def test_heap_small_components_stale_tail():
    """Components smaller than K+1: the reference keeps the previous source's tail (pyx:350)."""
    from rvgp_b200 import geometry as geo
    from oracle import rvgp_oracle as O
    rng = np.random.default_rng(0)
    # three far-apart clusters of 12, 13 and 40 points with nb=10 -> K=15 > component size for two of them
    X = np.concatenate([rng.normal(size=(40, 3)), rng.normal(size=(12, 3)) + 100, rng.normal(size=(13, 3)) - 100])
    knn = O.knn_exact(X, 10)
    ip, ix = O.symmetrize_csr(knn)
    seq_ref, cnt_ref = O.geodesic_neighbourhoods(ip, ix, 15)
    seq, cnt = geo.geodesic_neighbourhoods_device(_t(ip), _t(ix), 15)
    assert np.array_equal(cnt.cpu().numpy(), cnt_ref) and cnt_ref.min() < 16
    assert np.array_equal(seq.cpu().numpy(), seq_ref)


@pytest.mark.parametrize("case", CASES)
def test_tangent_frames_dimension_connections(case):
    from rvgp_b200 import geometry as geo
    g = load_golden(case)
    X = g["X"]
    n, D = X.shape
    T, S = geo.tangent_frames_device(_t(X), _t(g["geo_seq"]), D)
    np.testing.assert_allclose(S.cpu().numpy(), g["Sigma"], rtol=1e-12, atol=1e-14)
    var_exp = geo.explained_variance_device(S)
    d = int(np.where(var_exp >= 0.8)[0][0] + 1)
    assert d == int(g["dim_man"])
    G = geo.slice_frames_device(T, d).cpu().numpy()
    P = np.einsum("nip,njp->nij", G, G)
    np.testing.assert_allclose(P, g["projectors_full"], atol=1e-10)
    # connections + Lc from the REFERENCE gauges: the polar factor is unique, so blocks must match LAPACK's
    Lc = geo.connections_device(_t(g["gauges"]), _t(g["indptr"]), _t(g["indices"])).cpu().numpy()
    np.testing.assert_allclose(Lc, g["Lc_data"], atol=1e-12)
    # invariants with our own gauges
    Lc2, R2 = geo.connections_device(_t(G), _t(g["indptr"]), _t(g["indices"]), want_R=True)
    R2 = R2.cpu().numpy()
    assert np.abs(np.einsum("eij,ekj->eik", R2, R2) - np.eye(d)).max() < 1e-12
    A = sp.bsr_matrix((Lc2.cpu().numpy(), g["indices"], g["indptr"]), shape=(n * d, n * d)).tocsr()
    assert abs(A - A.T).max() < 1e-12


def test_tangent_frames_rank_deficient_raises():
    from rvgp_b200 import geometry as geo
    X = np.zeros((50, 3))
    X[:, :2] = np.random.default_rng(0).normal(size=(50, 2))           # exactly planar -> S[2] = 0
    from oracle import rvgp_oracle as O
    ip, ix = O.symmetrize_csr(O.knn_exact(X, 10))
    seq, _ = geo.geodesic_neighbourhoods_device(_t(ip), _t(ix), 15)
    with pytest.raises(RuntimeError, match="does not span"):
        geo.tangent_frames_device(_t(X), seq, 3)
    with pytest.raises(RuntimeError):
        O.tangent_frames(X, ip, ix, 3, 15)


@pytest.mark.parametrize("case", CASES)
def test_fps_bit_exact(case):
    from rvgp_b200.geometry import furthest_point_sampling
    g = load_golden(case)
    spacing = {"sphere_n2000_k50": 0.05, "torus_n600_k20": 0.1}.get(case, 0.2)
    perm, lam = furthest_point_sampling(g["X"], spacing=spacing)
    assert perm.dtype == np.int32
    assert np.array_equal(perm, g["fps_perm"])
    np.testing.assert_allclose(lam, g["fps_lambdas"], rtol=1e-12)
    perm, lam = furthest_point_sampling(g["X"], N=40, start_idx=3)
    assert np.array_equal(perm, g["fps_perm_N40"])
    np.testing.assert_allclose(lam, g["fps_lambdas_N40"], rtol=1e-12)
    p0, l0 = furthest_point_sampling(g["X"], spacing=0.0)
    assert l0 is None and np.array_equal(p0, np.arange(len(g["X"])))
    p1, _ = furthest_point_sampling(g["X"], stop_crit=spacing)          # README alias
    assert np.array_equal(p1, g["fps_perm"])


@pytest.mark.parametrize("case", CASES)
def test_data_object_matches_reference(case):
    from rvgp_b200.dataclass import data
    g = load_golden(case)
    X = g["X"]
    n, D = X.shape
    k = len(g["evals_Lc"])
    d = data(X, n_neighbors=GOLDEN_NB[case], n_eigenpairs=k, verbose=False)
    assert d.dim_man == int(g["dim_man"]) and d.n == n
    hi = 2.0 * (np.diff(g["indptr"]).max() - 1)
    np.testing.assert_allclose(d.evals_Lc, g["evals_Lc"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(d.evals_L, g["evals_L"], rtol=1e-8, atol=1e-8 * hi)
    assert d.evecs_Lc.shape == (n * D, k) and d.evecs_L.shape == (n, k)
    for s in eigen_clusters(g["evals_Lc"])[:-1]:
        assert subspace_angle_max(d.evecs_Lc[:, s], g["evecs_Lc"][:, s]) < 1e-6
    for s in eigen_clusters(g["evals_L"])[:-1]:
        assert subspace_angle_max(d.evecs_L[:, s], g["evecs_L"][:, s]) < 1e-6
    # host-object attributes
    Lref = sp.csr_matrix((g["L_data"], g["L_indices"], g["L_indptr"]), shape=(n, n))
    assert abs(d.L - Lref).max() == 0
    dm = d.dim_man
    assert d.Lc.shape == (n * dm, n * dm) and d.R.shape == (n * dm, n * dm)
    P = np.einsum("nip,njp->nij", d.gauges, d.gauges)
    np.testing.assert_allclose(P, g["projectors_full"], atol=1e-10)
    ev = np.sort(np.linalg.eigvalsh(d.Lc.toarray()))[:k] if n * dm <= 2000 else None
    if ev is not None:
        np.testing.assert_allclose(ev, g["evals_Lc"], rtol=1e-8, atol=1e-10)
    assert d.G.number_of_nodes() == n
    # random field: same host RNG stream, projection on the device
    d.random_vector_field(seed=1)
    np.testing.assert_allclose(d.vectors, g["random_field_seed1"], atol=1e-12)
    if "smoothed_field" in g:
        d.smooth_vector_field(t=float(g["smooth_t"]))
        np.testing.assert_allclose(d.vectors, g["smoothed_field"], atol=1e-9)
        np.testing.assert_allclose(np.linalg.norm(d.vectors, axis=1), 1.0, atol=1e-9)


@pytest.mark.parametrize("kind,n,D", [("torus", 20000, 3), ("sphere", 5000, 3), ("plane2", 6000, 2), ("line1", 5000, 1), ("clustered", 8000, 3)])
def test_knn_grid_bit_identical_to_brute(kind, n, D):
    """Grid-accelerated kNN (D <= 3) must be bit-identical to the brute-force kernel, including exact ties."""
    from rvgp_b200 import geometry as geo
    from tests.workloads import make_cloud
    rng = np.random.default_rng(0)
    if kind in ("torus", "sphere"):
        X = make_cloud(kind, n, 0)
    elif kind == "clustered":
        X = np.concatenate([rng.normal(scale=0.01, size=(n // 2, 3)), rng.normal(scale=1.0, size=(n // 2, 3)) + 5.0])
        X[10] = X[11]                                               # exact duplicate
    else:
        X = rng.uniform(size=(n, D))
        X = np.round(X * 50) / 50 if kind == "plane2" else X        # lattice -> many exact distance ties
    Xd = _t(X)
    for k in (10, 22):
        a, ad = geo.knn_device(Xd, k, return_d2=True, method="brute")
        b, bd = geo.knn_device(Xd, k, return_d2=True, method="grid")
        assert torch.equal(a, b) and torch.equal(ad, bd)
    part = geo.knn_device(Xd, 10, q_begin=1000, q_count=777, method="grid")
    assert torch.equal(part, geo.knn_device(Xd, 10, method="brute")[1000:1777])


def test_full_decomposition_when_n_eigenpairs_is_none():
    """n_eigenpairs=None -> k = N (geometry.py:68-71): the block solver degenerates to one exact Rayleigh-Ritz."""
    from rvgp_b200.dataclass import data
    from tests.workloads import make_cloud
    X = make_cloud("sphere", 300, 3)
    d = data(X, verbose=False)
    assert d.evecs_L.shape == (300, 300) and d.evecs_Lc.shape == (900, 600)
    ev = np.linalg.eigvalsh(d.Lc.toarray())
    np.testing.assert_allclose(d.evals_Lc, ev, rtol=1e-9, atol=1e-10)
    evL = np.linalg.eigvalsh(d.L.toarray())
    np.testing.assert_allclose(d.evals_L, evL, rtol=1e-9, atol=1e-10)
    d2 = data(X, n_eigenpairs=250, verbose=False)            # k close to N
    np.testing.assert_allclose(d2.evals_L, evL[:250], rtol=1e-8, atol=1e-9)


def test_geodesic_source_ranges_equal_full_run():
    """Multi-GPU form of K4 (rvgp_geodesic_neighbourhoods_range + rvgp_geodesic_fix_stale): sources are independent, so
    computing them in contiguous ranges and running the stale-tail pass once on the assembled arrays reproduces the
    single call bit for bit -- including the reference's stale-tail quirk across a range boundary (pyx:350)."""
    from rvgp_b200 import geometry as geo
    from rvgp_b200._cabi import get_handle, I64
    from oracle import rvgp_oracle as O
    rng = np.random.default_rng(0)
    X = np.concatenate([rng.normal(size=(40, 3)), rng.normal(size=(12, 3)) + 100, rng.normal(size=(13, 3)) - 100,
                        rng.normal(size=(300, 3)) * 3 + 30])
    ip, ix = O.symmetrize_csr(O.knn_exact(X, 10))
    n, K = len(X), 15
    ipd, ixd = _t(ip), _t(ix)
    seq_full, cnt_full = geo.geodesic_neighbourhoods_device(ipd, ixd, K)
    seq_ref, cnt_ref = O.geodesic_neighbourhoods(ip, ix, K)
    assert np.array_equal(seq_full.cpu().numpy(), seq_ref) and cnt_ref.min() < K + 1
    h = get_handle(0)
    maxrow = int(np.diff(ip).max())
    wsb = h.query("rvgp_geodesic_workspace_bytes", h._h, int(n), K, maxrow)
    ws = torch.empty(wsb, dtype=torch.uint8, device=_dev())
    seq = torch.zeros((n, K + 1), dtype=torch.int32, device=_dev())
    cnt = torch.zeros(n, dtype=torch.int32, device=_dev())
    allflags = 0
    for s0, s1 in ((0, 45), (45, 46), (46, 46), (46, n)):          # boundaries inside the short components, an empty range
        fl = torch.zeros(1, dtype=torch.int32, device=_dev())
        h.call("rvgp_geodesic_neighbourhoods_range", ipd, ixd, int(n), K, maxrow, int(s0), int(s1 - s0), seq, cnt, fl, ws, I64(wsb))
        allflags |= int(fl.item())
    assert allflags & 1
    fl = torch.tensor([allflags], dtype=torch.int32, device=_dev())
    h.call("rvgp_geodesic_fix_stale", int(n), K, seq, cnt, fl)
    assert np.array_equal(cnt.cpu().numpy(), cnt_ref)
    assert np.array_equal(seq.cpu().numpy(), seq_ref)
    with pytest.raises(ValueError):
        h.call("rvgp_geodesic_neighbourhoods_range", ipd, ixd, int(n), K, maxrow, int(n - 3), 10, seq, cnt, fl, ws, I64(wsb))
