"""Heat-diffusion smoothing on the GPU (mirror of the reference's ``RVGP/smoothing.py``).

Reference: ``scalar_diffusion`` (smoothing.py:8-34) forms the DENSE Pade ``expm(-t A) @ x`` (:13-15) and
``vector_diffusion`` (:37-64) combines one connection-Laplacian diffusion with two scalar ones.  Here the
action ``exp(-t A) x`` is evaluated without ever forming a matrix:

  * eigenbasis part  U exp(-t Lambda) U^T x      (north_star: "reuses the eigenbasis"; K10 dgemm kernels)
  * complement part  exp(-t A) (x - U U^T x)     by a Chebyshev expansion of exp(-t lambda) on
    [lambda_k, hi] driven by the fused block-SpMM (K9) -- only when exp(-t lambda_k) is not negligible
    (SURVEY.md H5), so the result equals the reference's dense expm to ~1e-12 at every size.
"""
import math

import numpy as np
import torch

from ._cabi import get_handle, I64
from .eigensolver import BsrMatrix, _dgemm


def _axpy(h, a, X, Y):
    h.call("rvgp_axpy_f64", I64(X.shape[0]), int(X.shape[1]), float(a), X, I64(X.stride(0)), Y, I64(Y.stride(0)))


def cheb_expm_action(A, x, t, lo, hi, tol=1e-14, stats=None):
    """exp(-t A) x for a symmetric BsrMatrix with spectrum (of the relevant invariant subspace) in [lo, hi].
    x: (N, c) cuda f64, c <= 64.  Chebyshev series with modified-Bessel coefficients; every term is one
    fused SpMM launch (three-term recurrence) plus one axpy."""
    from scipy.special import ive
    h = get_handle(x.device.index)
    z = 0.5 * t * (hi - lo)
    # f(lam) = exp(-t lo) * [ive(0,z) + 2 sum_j (-1)^j ive(j,z) T_j(y)],  y = (2 lam - lo - hi)/(hi - lo)
    jmax = int(min(20000, 40 + 1.2 * z + 12 * math.sqrt(max(z, 1.0))))
    coef = ive(np.arange(jmax + 1), z)
    coef[1:] *= 2.0 * (-1.0) ** np.arange(1, jmax + 1)
    coef *= math.exp(-t * lo)
    keep = np.nonzero(np.abs(coef) > tol * np.abs(coef).max())[0]
    deg = int(keep.max()) if keep.size else 0
    out = torch.zeros_like(x)
    # the three-term recurrence ROTATES its buffers, so from the third term on it writes into the one that held T_0: that must
    # never be the caller's x (round-2 finding: vector_diffusion(method="matrix_exp", normalise=True) computed |x| from an x the
    # connection-Laplacian diffusion had already overwritten)
    Tprev = x.clone() if deg >= 2 else x.contiguous()
    _axpy(h, coef[0], Tprev, out)
    if deg >= 1:
        s = 2.0 / (hi - lo)
        Tcur = torch.empty_like(x)
        A.spmm(Tprev, Tcur, alpha=s, beta=-(hi + lo) / (hi - lo), h=h)          # T_1 = y(A) x
        _axpy(h, coef[1], Tcur, out)
        Tnext = torch.empty_like(x)
        for j in range(2, deg + 1):
            A.spmm(Tcur, Tnext, alpha=2 * s, beta=-2 * (hi + lo) / (hi - lo), gamma=-1.0, W=Tprev, h=h)
            _axpy(h, coef[j], Tnext, out)
            Tprev, Tcur, Tnext = Tcur, Tnext, Tprev
    if stats is not None:
        stats["cheb_degree"] = deg
    return out


def spectral_action(evals, U, x, t, h=None):
    """U exp(-t Lambda) U^T x with unit-norm eigenvector columns U (N, k); x (N, c)."""
    h = h or get_handle(x.device.index)
    N, k = U.shape
    c = x.shape[1]
    coef = torch.empty((k, c), dtype=torch.float64, device=x.device)
    split = max(1, min(32, N // 4096))
    ws = torch.empty(split * k * c, dtype=torch.float64, device=x.device)
    _dgemm(h, k, c, N, U, U.stride(0), 0, x, x.stride(0), 0, coef, coef.stride(0), split_k=split, ws=ws)   # U^T x
    proj_coef = coef.clone()
    damp = torch.exp(-t * evals).reshape(k, 1)
    coef = (coef * damp).contiguous()
    out = torch.empty_like(x)
    _dgemm(h, N, c, k, U, U.stride(0), 1, coef, coef.stride(0), 0, out, out.stride(0))                        # U (..)
    return out, proj_coef


def diffuse(A, x, t, evals=None, U=None, hi=None, tol=1e-12, stats=None, lo=0.0):
    """exp(-t A) x.  With an eigenbasis (evals ascending, U unit-norm): spectral part + Chebyshev action on the
    deflated remainder when exp(-t*evals[-1]) > tol.  Without: Chebyshev on [0, hi]."""
    h = get_handle(x.device.index)
    x = x.contiguous()
    if evals is None:
        return cheb_expm_action(A, x, t, float(lo), hi, stats=stats)
    out, pc = spectral_action(evals, U, x, t, h)
    lam_k = float(evals[-1].item())
    trunc = math.exp(-t * max(lam_k, 0.0))
    if stats is not None:
        stats["truncation_bound"] = trunc
    if trunc > tol:
        # remainder r = x - U U^T x lives in the invariant subspace with spectrum in [lam_k, hi]
        N, k = U.shape
        r = torch.empty_like(x)
        _dgemm(h, N, x.shape[1], k, U, U.stride(0), 1, pc, pc.stride(0), 0, r, r.stride(0), alpha=-1.0)
        _axpy(h, 1.0, x, r)
        out2 = cheb_expm_action(A, r, t, lam_k, hi, tol=max(1e-16, 1e-2 * tol / trunc), stats=stats)
        _axpy(h, 1.0, out2, out)
    return out


def vector_diffusion_device(x_local, t, Lc, L, eig_Lc=None, eig_L=None, hi=None, normalise=True, stats=None, lo=(0.0, 0.0)):
    """smoothing.py:37-64 on device tensors.  x_local (n, d) local coordinates; Lc, L BsrMatrix;
    eig_* = (evals, unit-norm evecs) or None."""
    h = get_handle(x_local.device.index)
    n, d = x_local.shape
    out = diffuse(Lc, x_local.reshape(n * d, 1), t, *(eig_Lc or (None, None)), hi=hi, stats=stats, lo=lo[0]).reshape(n, d)
    if normalise:
        x_abs = torch.empty((n, 1), dtype=torch.float64, device=x_local.device)
        h.call("rvgp_row_norms_f64", I64(n), int(d), x_local.contiguous(), x_abs)
        both = torch.cat([x_abs, torch.ones_like(x_abs)], dim=1).contiguous()      # [|x|, 1]
        res = diffuse(L, both, t, *(eig_L or (None, None)), hi=hi, lo=lo[1])
        out_abs = res[:, 0].contiguous()
        ind = res[:, 1].contiguous()
        out = out.contiguous()
        h.call("rvgp_renorm_rows_f64", I64(n), int(d), out, out_abs, ind)
    return out


# ---- reference-named entry points (host arrays / scipy matrices in, numpy out) -----------------------------
def _bsr_from_scipy(M, dev):
    from scipy import sparse
    if sparse.isspmatrix_bsr(M) and M.blocksize[0] == M.blocksize[1]:
        d = M.blocksize[0]
        B = M
    else:
        d = 1
        B = sparse.csr_matrix(M)
    B.sort_indices()
    vals = torch.from_numpy(np.ascontiguousarray(B.data, dtype=np.float64).reshape(-1, d, d)).to(dev)
    A = BsrMatrix(M.shape[0] // d, d, torch.from_numpy(B.indptr.astype(np.int32)).to(dev),
                  torch.from_numpy(B.indices.astype(np.int32)).to(dev), vals)
    C = sparse.csr_matrix(M)
    absrow = np.asarray(abs(C).sum(1)).reshape(-1)
    hi = float(absrow.max())
    # Gershgorin LOWER bound (symmetric matrices): min_i (a_ii - sum_{j != i} |a_ij|), clamped at 0.  It is 0 for a full graph
    # Laplacian, but positive for principal sub-matrices (eeg_utils.interpolate_timepoint diffuses on the sub-graph of the
    # training nodes): starting the Chebyshev expansion of exp(-t A) at lo keeps its accuracy RELATIVE to exp(-t lo) ||x||
    # instead of ||x||, which matters when the result is renormalised afterwards (smoothing.py:56-61)
    diag = C.diagonal()
    lo = float(max(0.0, (2.0 * diag - absrow).min())) if C.shape[0] else 0.0
    return A, hi, lo


def scalar_diffusion(x, t, method="matrix_exp", par=None):
    """smoothing.py:8-34.  method="matrix_exp": par is a scipy sparse matrix; "spectral": par=(evals, evecs)
    with unit-norm eigenvector columns."""
    from .geometry import to_device_f64
    xd = to_device_f64(x)
    if xd.dim() == 1:
        xd = xd.unsqueeze(1)
    if method == "matrix_exp":
        A, hi, lo = _bsr_from_scipy(par, xd.device)
        return cheb_expm_action(A, xd, float(t), lo, hi).cpu().numpy()
    if method == "spectral":
        assert isinstance(par, (list, tuple)) and len(par) == 2, \
            "For spectral method, par must be a tuple of eigenvalues, eigenvectors!"
        evals, evecs = to_device_f64(par[0]), to_device_f64(par[1])
        out, _ = spectral_action(evals, evecs, xd, float(t))
        return out.cpu().numpy()
    raise NotImplementedError


def vector_diffusion(x, t, Lc, L=None, method="spectral", normalise=True):
    """smoothing.py:37-64 (host arrays in, numpy out; all arithmetic on the device)."""
    from .geometry import to_device_f64
    x = np.asarray(x)
    n, d = x.shape[0], x.shape[1]
    nd = Lc[1].shape[0] if method == "spectral" else Lc.shape[0]
    assert (n * d % nd) == 0, "Data dimension must be an integer multiple of the dimensions of the connection Laplacian!"
    if normalise:
        assert L is not None, "Need Laplacian for normalised diffusion!"
    xd = to_device_f64(x)
    if method == "matrix_exp":
        A_c, hi_c, lo_c = _bsr_from_scipy(Lc, xd.device)
        A_s, hi_s, lo_s = _bsr_from_scipy(L, xd.device) if normalise else (None, 0.0, 0.0)
        out = vector_diffusion_device(xd, float(t), A_c, A_s, hi=max(hi_c, hi_s), normalise=normalise, lo=(lo_c, lo_s))
    elif method == "spectral":
        assert len(Lc) == 2, "Lc must be a tuple of eigenvalues, eigenvectors!"
        eig_c = (to_device_f64(Lc[0]), to_device_f64(Lc[1]))
        eig_s = (to_device_f64(L[0]), to_device_f64(L[1])) if normalise else None
        out = _vector_diffusion_spectral_only(xd, float(t), eig_c, eig_s, normalise)
    else:
        raise NotImplementedError
    return out.cpu().numpy()


def _vector_diffusion_spectral_only(xd, t, eig_c, eig_s, normalise):
    """Pure eigenbasis diffusion (what the reference's dead "spectral" branch intends, smoothing.py:17-32)."""
    h = get_handle(xd.device.index)
    n, d = xd.shape
    out, _ = spectral_action(eig_c[0], eig_c[1], xd.reshape(n * d, 1).contiguous(), t, h)
    out = out.reshape(n, d).contiguous()
    if normalise:
        x_abs = torch.empty((n, 1), dtype=torch.float64, device=xd.device)
        h.call("rvgp_row_norms_f64", I64(n), int(d), xd.contiguous(), x_abs)
        both = torch.cat([x_abs, torch.ones_like(x_abs)], dim=1).contiguous()
        res, _ = spectral_action(eig_s[0], eig_s[1], both, t, h)
        h.call("rvgp_renorm_rows_f64", I64(n), int(d), out, res[:, 0].contiguous(), res[:, 1].contiguous())
    return out
