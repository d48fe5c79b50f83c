"""CPU tests of the HOST LOGIC of the filtered block Lanczos eigensolver (rvgp_b200/krylov.py) against SciPy's ARPACK -- the
reference's own back end (geometry.py:73) -- through tests/fake_cabi.py (a NumPy emulation of the C-ABI calls; test
infrastructure, see its header).  Covers: the real and the complex (paired) block algebra, thick restarts, the
Lanczos-estimate -> true-residual hand-over, the final Rayleigh-Ritz through ChFSI, and the retry when the cut was placed too low."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

from tests import fake_cabi
from tests.workloads import make_cloud
from oracle import rvgp_oracle as O


def _laplacian(n, seed=0):
    X = make_cloud("torus", n, seed)
    ip, ix = O.symmetrize_csr(O.knn_sklearn(X, 10))
    L = O.laplacian(ip, ix)
    return L, 2.0 * (np.diff(ip).max() - 1)


def _check(A, evals, evecs, k, ref, tol_abs):
    ev = evals.numpy()
    U = evecs.numpy()
    np.testing.assert_allclose(ev, ref[:k], rtol=1e-8, atol=1e-10)
    res = np.linalg.norm(A @ U - U * ev, axis=0)
    assert res.max() <= tol_abs * 1.0001, res.max()
    assert np.abs(U.T @ U - np.eye(k)).max() < 1e-9


@pytest.mark.parametrize("cap_cols,fast_accept", [(None, "1"), (None, "0"), (224, "1")])   # 224 columns: forces thick restarts
def test_krylov_real_matches_arpack(monkeypatch, cap_cols, fast_accept):
    from rvgp_b200.krylov import krylov_eigenpairs
    fake_cabi.install_eigensolver(monkeypatch)
    monkeypatch.setenv("RVGP_KRYLOV_FAST_ACCEPT", fast_accept)
    L, hi = _laplacian(3000)
    k = 40
    ref = np.sort(spla.eigsh(L, k=90, which="SM", return_eigenvectors=False))
    A = fake_cabi.FakeBsr(L)
    st = {}
    evals, evecs = krylov_eigenpairs(A, k, hi, cut=1.05 * ref[int(1.5 * k) + 16], lam_k=ref[k - 1], block=32, stats=st,
                                     cap_cols=cap_cols)
    _check(L, evals, evecs, k, ref, 1e-12 * hi)
    assert st["converged"] and st["krylov_converged"]
    assert st["final_rr_outer"] == (0 if fast_accept == "1" else 1)   # 0: Ritz block accepted as it is; 1: one A-space Rayleigh-Ritz, no polishing
    assert (st["restarts"] > 0) == (cap_cols is not None)
    # the point of the method: far fewer column-degrees than subspace iteration needs (~ m * 27 / g per column)
    from rvgp_b200.eigensolver import smallest_eigenpairs
    A2 = fake_cabi.FakeBsr(L)
    st2 = {}
    smallest_eigenpairs(A2, k, hi, stats=st2)
    assert st["filter_col_degrees"] < 0.6 * st2["filter_col_degrees"], (st["filter_col_degrees"], st2["filter_col_degrees"])


def test_krylov_recovers_from_a_cut_below_lambda_k(monkeypatch):
    from rvgp_b200.krylov import krylov_eigenpairs
    fake_cabi.install_eigensolver(monkeypatch)
    L, hi = _laplacian(3000)
    k = 40
    ref = np.sort(spla.eigsh(L, k=60, which="SM", return_eigenvectors=False))
    st = {}
    evals, evecs = krylov_eigenpairs(fake_cabi.FakeBsr(L), k, hi, cut=0.6 * ref[k - 1], lam_k=0.4 * ref[k - 1], block=32, stats=st)
    _check(L, evals, evecs, k, ref, 1e-12 * hi)
    assert st.get("cut_retries", 0) >= 1 or st["final_rr_outer"] > 1       # it noticed (retry) or ChFSI finished the job


@pytest.mark.parametrize("fast_accept", ["1", "0"])      # Ritz block accepted directly / ChFSI Rayleigh-Ritz hand-over
def test_krylov_paired_matches_arpack(monkeypatch, fast_accept):
    """Complex-Hermitian operator in real 2x2-block storage (every block a scaled rotation): eigenvalues come in exact pairs
    and the solver works on half the columns (eigensolver.py paired mode)."""
    from rvgp_b200.krylov import krylov_eigenpairs
    fake_cabi.install_eigensolver(monkeypatch)
    monkeypatch.setenv("RVGP_KRYLOV_FAST_ACCEPT", fast_accept)
    L, hi = _laplacian(1500, seed=1)
    n = L.shape[0]
    rng = np.random.default_rng(0)
    Lc = sp.coo_matrix(L)
    # magnetic Laplacian: off-diagonal entries -exp(i phi_ij), phi antisymmetric -> Hermitian, PSD, spectrum below 2 max degree
    phi = {}
    data = np.empty(Lc.nnz, dtype=np.complex128)
    for e, (i, j, v) in enumerate(zip(Lc.row, Lc.col, Lc.data)):
        if i == j:
            data[e] = v
        else:
            key = (min(i, j), max(i, j))
            if key not in phi:
                phi[key] = rng.uniform(-0.3, 0.3)
            data[e] = v * np.exp(1j * (phi[key] if i < j else -phi[key]))
    Hc = sp.csr_matrix((data, (Lc.row, Lc.col)), shape=(n, n))
    # real storage: z = x + i y at rows (2i, 2i+1); H z -> [[Re, -Im], [Im, Re]] blocks
    Ar = sp.bmat([[Hc.real, -Hc.imag], [Hc.imag, Hc.real]]).tocsr()
    perm = np.empty(2 * n, dtype=np.int64)
    perm[0::2] = np.arange(n)
    perm[1::2] = n + np.arange(n)
    Ar = Ar[perm][:, perm].tocsr()
    k = 40
    ref = np.sort(spla.eigsh(Ar, k=80, which="SM", return_eigenvectors=False))
    assert np.allclose(ref[0::2], ref[1::2])
    st = {}
    evals, evecs = krylov_eigenpairs(fake_cabi.FakeBsr(Ar, d=2), k, hi, cut=1.05 * ref[int(1.5 * k) + 16], lam_k=ref[k - 1],
                                     paired=True, block=16, stats=st)
    _check(Ar, evals, evecs, k, ref, 1e-12 * hi)
    assert st["converged"] and st["paired"]
    assert st["final_rr_outer"] == (0 if fast_accept == "1" else 1)
    U = evecs.numpy()
    JU = np.empty_like(U[:, 0::2])
    JU[0::2] = -U[1::2, 0::2]
    JU[1::2] = U[0::2, 0::2]
    assert np.allclose(U[:, 1::2], JU)                   # columns come as (v, J v) pairs
