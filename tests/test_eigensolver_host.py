"""CPU tests of the eigensolver's HOST algebra (rvgp_b200/eigensolver.py): the pieces that run in NumPy between kernel launches
-- Hermitian reconstruction from lower-triangle tiles, (shifted) Cholesky factors, the per-column Chebyshev degree rule, and
the complex <-> real-pair identities the paired mode relies on."""
import numpy as np
import torch

from rvgp_b200 import eigensolver as E


def _rot90(V):
    out = np.empty_like(V)
    out[0::2] = -V[1::2]
    out[1::2] = V[0::2]
    return out


def test_hermitian_from_lower_tiles_and_complex_identities():
    rng = np.random.default_rng(0)
    n, m = 40, 6
    V = rng.normal(size=(2 * n, m))
    W = rng.normal(size=(2 * n, m))
    Z = V[0::2] + 1j * V[1::2]                       # column c of the real (2n x m) block IS one complex n-vector
    JV = _rot90(V)
    assert np.allclose(JV[0::2] + 1j * JV[1::2], 1j * Z)                           # J = multiplication by i
    G = Z.conj().T @ Z
    assert np.allclose(V.T @ V, G.real) and np.allclose(JV.T @ V, G.imag)          # V^H V = V^T V + i (J V)^T V
    # what the device hands back: only tiles touching the lower triangle are defined; poison the rest
    Gr = np.tril(V.T @ V) + np.triu(np.full((m, m), np.nan), 1)
    Gi = np.tril(JV.T @ V) + np.triu(np.full((m, m), np.nan), 1)
    H = E._herm_from_lower(torch.from_numpy(Gr), torch.from_numpy(Gi))
    assert np.allclose(H, G) and np.allclose(H, H.conj().T) and np.all(np.diag(H).imag == 0)
    # apply: V C = V Re(C) + (J V) Im(C)
    C = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
    out = V @ C.real + JV @ C.imag
    assert np.allclose(out[0::2] + 1j * out[1::2], Z @ C)
    # the pair (v, J v) spans a J-invariant plane: both are eigenvectors of any J-commuting operator
    a, b = rng.normal(size=(n, n)), rng.normal(size=(n, n))
    Hc = (a + a.T) + 1j * (b - b.T)
    A = np.zeros((2 * n, 2 * n))
    A[0::2, 0::2], A[0::2, 1::2], A[1::2, 0::2], A[1::2, 1::2] = Hc.real, -Hc.imag, Hc.imag, Hc.real
    assert np.allclose(A, A.T) and np.allclose(A @ JV, _rot90(A @ V))
    ev = np.linalg.eigvalsh(A)
    assert np.allclose(ev[0::2], ev[1::2]) and np.allclose(ev[0::2], np.linalg.eigvalsh(Hc))
    del W


def test_shifted_cholesky_real_and_complex():
    rng = np.random.default_rng(1)
    m = 12
    B = rng.normal(size=(m, 3 * m)) + 1j * rng.normal(size=(m, 3 * m))
    G = B @ B.conj().T
    R, shifted = E._chol_upper_shifted_c(G)
    assert not shifted and np.allclose(R.conj().T @ R, G) and np.allclose(R, np.triu(R))
    Rinv = E._tri_inv_upper(R)
    assert np.allclose(R @ Rinv, np.eye(m))
    # numerically singular Gram matrix (a repeated column): the shifted factor exists and is flagged
    Bs = B.copy()
    Bs[1] = Bs[0]
    Gs = Bs @ Bs.conj().T
    Gs[1, 1] = Gs[0, 0] * (1 - 1e-15)
    Rs, shifted = E._chol_upper_shifted_c(Gs - 1e-9 * np.eye(m))
    assert shifted and np.all(np.isfinite(Rs))
    Gr = G.real + np.eye(m)
    Rr, sh = E._chol_upper_shifted(Gr)
    assert not sh and np.allclose(Rr.T @ Rr, Gr)
    _, sh2 = E._chol_upper_shifted(Gr - (np.linalg.eigvalsh(Gr)[0] + 1e-9) * np.eye(m))
    assert sh2


def test_degree_rule():
    k, m = 6, 10
    theta = np.linspace(0.01, 0.1, m)
    hi, lo, tol = 30.0, 0.0, 1e-11
    a_cut = theta[-1]
    res = np.full(m, 1e-3)
    res[:2] = 1e-13                                  # converged columns are not filtered again
    d = E._next_degrees(theta, res, k, tol, a_cut, hi, lo, 1e6)
    assert d.dtype == np.int64 and np.all(d[:2] == 0) and np.all(d[2:k] >= 8)
    assert np.all(np.diff(d[2:k]) >= 0)              # columns closer to the cut need more degrees
    assert np.all(d[k:] <= d[:k].max())              # buffer columns never cost more than the wanted ones
    # the cap keeps the filtered block numerically full rank: amplification of the lowest direction <= cond_max
    e, c = 0.5 * (hi - a_cut), 0.5 * (hi + a_cut)
    g = np.arccosh(np.maximum((c - theta) / e, 1.0))
    g0 = np.arccosh((c - lo) / e)
    assert np.all(d * (g0 - g) <= np.log(1e6) * 1.0001 + (g0 - g))
    # smaller residuals need fewer degrees
    d2 = E._next_degrees(theta, res * 1e-4, k, tol, a_cut, hi, lo, 1e6)
    assert np.all(d2[2:k] <= d[2:k])


def test_lapack_thread_context_is_reentrant_and_restores():
    from threadpoolctl import threadpool_info
    before = [(d["internal_api"], d["num_threads"]) for d in threadpool_info()]
    with E._lapack_ctx():
        with E._lapack_ctx():
            np.linalg.eigh(np.eye(8))
        inside = [d["num_threads"] for d in threadpool_info() if d["user_api"] == "blas"]
    after = [(d["internal_api"], d["num_threads"]) for d in threadpool_info()]
    assert before == after and all(1 <= t <= 4 for t in inside)
