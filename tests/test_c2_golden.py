"""Parity at BASELINE config C2 size (35 000-point torus, k = 200) against tests/golden/torus_c2_n35000_k200.npz, produced by
the UNMODIFIED reference (tests/golden/make_golden_c2.py; 171 s on the build container).

CPU (not gpu): the oracle's front end (kNN, CSR, L, gauges, Lc blocks) against the golden -- the ARPACK stages are too slow
for the CPU suite.  GPU: the whole CUDA data object -- eigenvalues to 1e-8 relative, subspace angles of EVERY eigenvalue
cluster below 1e-6 through a seeded random sketch, exact principal angles on the stored clusters (geometry.py:66-80)."""
import hashlib
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN_DIR, eigen_clusters, subspace_angle_max
from tests.workloads import make_cloud

PATH = os.path.join(GOLDEN_DIR, "torus_c2_n35000_k200.npz")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden():
    return dict(np.load(PATH))


def _sketch_matrix(nrows, r, seed):
    return np.random.default_rng(seed).standard_normal((nrows, r)) / np.sqrt(nrows)


def sketch_angle_estimates(evals, U, sketch_ref, r, seed):
    """Estimated sin(max principal angle) per eigenvalue cluster from the r x r sketches Omega^T P Omega of the two
    projectors (P = U_C U_C^T / N for eigenvectors scaled by sqrt(N)):  E||Omega^T D Omega||_F^2 = r (r + 1) ||D||_F^2 / N^2
    for D = P_ref - P_ours, and ||D||_F^2 = 2 sum sin^2(theta_i)."""
    N = U.shape[0]
    S = _sketch_matrix(N, r, seed).T @ U
    out = []
    for s in eigen_clusters(evals):
        Mr = sketch_ref[:, s] @ sketch_ref[:, s].T / N
        Mo = S[:, s] @ S[:, s].T / N
        dF = N * np.linalg.norm(Mr - Mo) / np.sqrt(r * (r + 1))
        out.append(dF / np.sqrt(2.0))
    return np.array(out)


def test_oracle_front_end_matches_c2_golden():
    from oracle import rvgp_oracle as O
    g = _golden()
    n, nb = int(g["n"]), int(g["nb"])
    X = make_cloud("torus", n, 0)
    knn = O.knn_sklearn(X, nb)
    assert _sha(knn) == str(g["knn_sha"])
    indptr, indices = O.symmetrize_csr(knn)
    assert indices.size == int(g["nnz_L"])
    assert _sha(indices) == str(g["L_indices_sha"]) and _sha(indptr) == str(g["L_indptr_sha"])
    L = O.laplacian(indptr, indices)
    assert _sha(L.data) == str(g["L_data_sha"])
    tangents, Sigma = O.tangent_frames(X, indptr, indices, 3, nb * 1.5)
    dim, _ = O.manifold_dimension(Sigma, 0.8)
    assert dim == int(g["dim_man"])
    gauges = np.ascontiguousarray(tangents[:, :, :dim])
    P = np.einsum("nip,njp->nij", gauges[::500], gauges[::500])
    assert np.abs(P - g["gauges_projector_sample"]).max() < 1e-10
    R = O.connections(gauges, indptr, indices)
    Lc = O.connection_laplacian(indptr, indices, R)
    assert Lc.data.shape[0] * dim * dim == int(g["nnz_Lc"])
    # blocks depend on the SVD's sign / ordering conventions only through U V^T, which is unique
    assert np.abs(Lc.data[g["Lc_block_sample_idx"]] - g["Lc_block_sample"]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("solver", ["auto", "krylov"])
def test_cuda_data_object_matches_c2_golden(solver, monkeypatch):
    """solver = "krylov" forces the filtered block Lanczos eigensolver (krylov.py; the default only from n = 100 000 on), so
    that BOTH eigensolvers are pinned to the reference's ARPACK output at the largest size the reference can run."""
    import RVGP
    import torch
    monkeypatch.setenv("RVGP_EIGSOLVER", solver)
    g = _golden()
    n, k, nb = int(g["n"]), int(g["k"]), int(g["nb"])
    X = make_cloud("torus", n, 0)
    d = RVGP.create_data_object(X, n_neighbors=nb, n_eigenpairs=k, verbose=False)
    assert d.dim_man == int(g["dim_man"])
    assert ("Lanczos" in str(d.stats["eig_L"].get("solver"))) == (solver == "krylov")
    assert ("Lanczos" in str(d.stats["eig_Lc"].get("solver"))) == (solver == "krylov")
    gr = d._graph
    assert _sha(np.sort(gr.knn.cpu().numpy(), axis=1).astype(np.int32)) == str(g["knn_sha"])            # bit-exact sets
    assert _sha(gr.indices.cpu().numpy().astype(np.int32)) == str(g["L_indices_sha"])
    assert _sha(gr.indptr.cpu().numpy().astype(np.int32)) == str(g["L_indptr_sha"])
    L = d.L
    assert L.nnz == int(g["nnz_L"]) and _sha(L.data) == str(g["L_data_sha"])
    P = np.einsum("nip,njp->nij", d.gauges[::500], d.gauges[::500])
    assert np.abs(P - g["gauges_projector_sample"]).max() < 1e-10
    Lc = d.Lc
    assert Lc.data.size == int(g["nnz_Lc"])
    # Lc is only defined up to the sign of each gauge vector (SVD convention: LAPACK there, Jacobi here): block (i,j) becomes
    # D_i R_ij D_j with diagonal +-1 matrices, so entries agree in ABSOLUTE value; signs are covered by the spectra below
    assert np.abs(np.abs(Lc.data[g["Lc_block_sample_idx"]]) - np.abs(g["Lc_block_sample"])).max() < 1e-9
    r, seed = int(g["sketch_r"]), int(g["sketch_seed"])
    for name in ("L", "Lc"):
        ev_ref = g["evals_" + name]
        ev = getattr(d, "evals_" + name)
        scale = max(abs(ev_ref).max(), 1e-300)
        # relative 1e-8 (north_star); the zero eigenvalue of L is compared on the scale of the spectrum
        assert np.abs(ev - ev_ref).max() <= 1e-8 * scale, (name, np.abs(ev - ev_ref).max())
        np.testing.assert_allclose(ev[1:], ev_ref[1:], rtol=1e-8, atol=1e-12)
        U = getattr(d, "evecs_" + name)
        est = sketch_angle_estimates(ev_ref, U, g["sketch_" + name], r, seed)
        assert est.max() < 1e-6, (name, est.max(), int(est.argmax()))
        ci = 0
        while "cluster_%s_%d_range" % (name, ci) in g:
            a, b = [int(v) for v in g["cluster_%s_%d_range" % (name, ci)]]
            ang = subspace_angle_max(U[:, a:b], g["cluster_%s_%d_vecs" % (name, ci)])
            assert ang < 1e-6, (name, ci, a, b, ang)
            ci += 1
        assert ci >= 2
    torch.cuda.empty_cache()
