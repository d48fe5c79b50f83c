"""Pins oracle/ (the CPU restatement) to golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU-only; no product code involved."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import rvgp_oracle as O
from tests.conftest import subspace_angle_max, eigen_clusters


def test_knn_sets(golden):
    X = golden["X"]
    assert np.array_equal(O.knn_sklearn(X, golden["nb"]), golden["knn"])
    if X.shape[1] <= 15:   # KD-tree path: exact sequential squared distances
        assert np.array_equal(np.sort(O.knn_exact(X, golden["nb"]), 1), golden["knn"])


def test_csr(golden):
    indptr, indices = O.symmetrize_csr(golden["knn"])
    assert np.array_equal(indptr, golden["indptr"])
    assert np.array_equal(indices, golden["indices"])


def test_heap_sequences(golden):
    K = int(golden["nb"] * 1.5)
    seq, counts = O.geodesic_neighbourhoods(golden["indptr"], golden["indices"], K)
    assert counts.min() == K + 1
    assert np.array_equal(seq, golden["geo_seq"])        # sequence equality, not just sets


def test_tangent_frames_and_dim(golden):
    X = golden["X"]
    D = X.shape[1]
    T, S = O.tangent_frames(X, golden["indptr"], golden["indices"], D, golden["nb"] * 1.5)
    np.testing.assert_allclose(S, golden["Sigma"], rtol=1e-10, atol=1e-13)
    dim, _ = O.manifold_dimension(S.copy(), 0.8)
    assert dim == int(golden["dim_man"])
    P = np.einsum("nip,njp->nij", T[:, :, :dim], T[:, :, :dim])
    np.testing.assert_allclose(P, golden["projectors_full"], atol=1e-9)


def test_laplacians_from_reference_gauges(golden):
    indptr, indices = golden["indptr"], golden["indices"]
    L = O.laplacian(indptr, indices)
    Lref = sp.csr_matrix((golden["L_data"], golden["L_indices"], golden["L_indptr"]), shape=L.shape)
    assert abs(L - Lref).max() == 0.0
    R = O.connections(golden["gauges"], indptr, indices)
    Lc = O.connection_laplacian(indptr, indices, R)
    d = int(golden["dim_man"])
    Lcref = sp.bsr_matrix((golden["Lc_data"], golden["Lc_indices"], golden["Lc_indptr"]), shape=Lc.shape)
    assert abs(Lc - Lcref).max() < 1e-12


def test_spectrum(golden):
    n = golden["X"].shape[0]
    k = len(golden["evals_L"])
    d = int(golden["dim_man"])
    L = sp.csr_matrix((golden["L_data"], golden["L_indices"], golden["L_indptr"]), shape=(n, n))
    ev, U = O.spectrum(L, k)
    np.testing.assert_allclose(ev, golden["evals_L"], rtol=1e-8, atol=1e-8 * 20)
    Lc = sp.bsr_matrix((golden["Lc_data"], golden["Lc_indices"], golden["Lc_indptr"]), shape=(n * d, n * d))
    evc, Uc = O.spectrum(Lc, k)
    np.testing.assert_allclose(evc, golden["evals_Lc"], rtol=1e-8, atol=1e-10)
    Phi = O.lift_eigenvectors(Uc, golden["gauges"])
    cl = eigen_clusters(golden["evals_Lc"])
    for s in cl[:-1]:       # the cluster straddling index k is an arbitrary slice
        assert subspace_angle_max(Phi[:, s], golden["evecs_Lc"][:, s]) < 1e-6


def test_fps(golden):
    name = golden["name"]
    spacing = {"sphere_n2000_k50": 0.05, "torus_n600_k20": 0.1}.get(name, 0.2)
    perm, lam = O.furthest_point_sampling(golden["X"], spacing=spacing)
    assert perm.dtype == np.int32
    assert np.array_equal(perm, golden["fps_perm"])
    np.testing.assert_allclose(lam, golden["fps_lambdas"], rtol=1e-12)
    perm, lam = O.furthest_point_sampling(golden["X"], N=40, start_idx=3)
    assert np.array_equal(perm, golden["fps_perm_N40"])
    np.testing.assert_allclose(lam, golden["fps_lambdas_N40"], rtol=1e-12)


def test_random_field_and_smoothing(golden):
    X = golden["X"]
    v = O.random_vector_field(X.shape[0], X.shape[1], golden["gauges"], seed=1)
    np.testing.assert_allclose(v, golden["random_field_seed1"], atol=1e-13)
    if "smoothed_field" in golden and X.shape[0] <= 700:
        n = X.shape[0]
        d = int(golden["dim_man"])
        L = sp.csr_matrix((golden["L_data"], golden["L_indices"], golden["L_indptr"]), shape=(n, n))
        Lc = sp.bsr_matrix((golden["Lc_data"], golden["Lc_indices"], golden["Lc_indptr"]), shape=(n * d, n * d))
        t = float(golden["smooth_t"])
        s = O.smooth_vector_field(golden["random_field_seed1"], golden["gauges"], t, L, Lc, dense=True)
        np.testing.assert_allclose(s, golden["smoothed_field"], atol=1e-11)
        s2 = O.smooth_vector_field(golden["random_field_seed1"], golden["gauges"], t, L, Lc, dense=False)
        np.testing.assert_allclose(s2, golden["smoothed_field"], atol=1e-9)


def test_full_pipeline_c1():
    from tests.conftest import load_golden
    g = load_golden("sphere_n2000_k50")
    o = O.create_data_object(g["X"], n_eigenpairs=50)
    assert o.dim_man == 2 and o.L.nnz == 24854 and o.Lc.nnz == 99416
    np.testing.assert_allclose(o.evals_Lc, g["evals_Lc"], rtol=1e-8)
    np.testing.assert_allclose(o.evals_L, g["evals_L"], rtol=1e-8, atol=2e-7)


def test_vectorfield_features_oracle_pinned_to_reference():
    """oracle.compute_vectorfield_features == the unmodified reference function (examples/eeg_example/eeg_utils.py:46-80,
    golden vectors from tests/golden/make_golden_eeg.py), bit for bit."""
    from tests.conftest import load_golden
    g = load_golden("eeg_features")
    for k in (5, 3):
        div, curl = O.compute_vectorfield_features(g["positions"], g["vectors"], k=k)
        assert np.array_equal(div, g["div_k%d" % k]) and np.array_equal(curl, g["curl_k%d" % k])
