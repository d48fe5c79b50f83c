"""TEST INFRASTRUCTURE ONLY: a NumPy emulation of the handful of librvgp_b200.so entry points the GP host code calls,
so that the HOST LOGIC of rvgp_b200/gp_general.py, kernels.py and main.py (argument order, layout flags, strides, the
hand-written adjoints' algebra, the L-BFGS-B plumbing) can be checked on a machine without a GPU (`-m "not gpu"`).

It is installed by monkeypatching inside tests/test_gp_general_host.py and nowhere else; the product has no CPU path
(rvgp_b200/_cabi.py raises without the library / a CUDA device).  Each emulated call follows the contract written in
include/rvgp_b200.h: raw "pointer + leading dimension" semantics are reproduced with torch.as_strided on CPU tensors, and
rvgp_potrf_f64 poisons the strict upper triangle (the header calls it scratch) so that host code relying on it fails here.
"""
import numpy as np
import torch

NB = 64


def _mat(t, rows, cols, ld):
    """rows x cols row-major view starting at t's first element with leading dimension ld."""
    if rows == 0 or cols == 0:
        return torch.empty((rows, cols), dtype=torch.float64)
    return torch.as_strided(t, (int(rows), int(cols)), (int(ld), 1))


def _op(t, rows, cols, ld, kmajor_is_cols):
    """Logical (rows x cols) operand: element (i, j) at t[i*ld + j] when kmajor_is_cols else t[j*ld + i]."""
    if kmajor_is_cols:
        return _mat(t, rows, cols, ld)
    return _mat(t, cols, rows, ld).t()


class FakeHandle:
    sm_count = 148
    _h = None

    def __init__(self):
        self.calls = {}
        self.launches = 0

    def sync_stream(self):
        pass

    def query(self, name, *a):
        if name == "rvgp_potrf_workspace_bytes":
            return 8
        if name == "rvgp_coldot_workspace_bytes":
            return 8
        if name == "rvgp_rbf_adjoint_workspace_bytes":
            return 8
        if name == "rvgp_fps_workspace_bytes":
            return 8
        raise NotImplementedError(name)

    def call(self, name, *a):
        self.calls[name] = self.calls.get(name, 0) + 1
        self.launches += 1
        return getattr(self, name)(*a)

    # ---- K10 ---------------------------------------------------------------------------------------------------
    def rvgp_dgemm_f64(self, m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, scale_k, C, ldc, split_k, ws):
        assert split_k == 1 or (ws is not None and ws.numel() >= split_k * m * n)
        Am = _op(A, m, k, lda, a_kmajor == 1)
        Bm = _op(B, k, n, ldb, b_kmajor == 0)
        if scale_k is not None:
            Am = Am * scale_k[:k][None, :]
        _mat(C, m, n, ldc).copy_(alpha * (Am @ Bm))
        return 0

    def rvgp_dgemm_acc_f64(self, m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, beta, C, ldc):
        Am = _op(A, m, k, lda, a_kmajor == 1)
        Bm = _op(B, k, n, ldb, b_kmajor == 0)
        Cm = _mat(C, m, n, ldc)
        Cm.copy_(alpha * (Am @ Bm) + beta * Cm)
        return 0

    def rvgp_coldot_f64(self, nrows, ncols, A, lda, B, ldb, out, ws):
        Am = _mat(A, nrows, ncols, lda)
        Bm = _mat(B, nrows, ncols, ldb) if B is not None else torch.ones_like(Am)
        out[:ncols].copy_((Am * Bm).sum(0))
        return 0

    def rvgp_colscale_f64(self, nrows, ncols, A, lda, s):
        Am = _mat(A, nrows, ncols, lda)
        Am.mul_(s[:ncols][None, :])
        return 0

    def rvgp_gather_rows_f64(self, nrows, ncols, inp, ldin, perm, block, out, ldout):
        r = torch.arange(nrows)
        src = perm.to(torch.int64)[r // block] * block + (r % block)
        nsrc = int(src.max().item()) + 1 if nrows else 0
        _mat(out, nrows, ncols, ldout).copy_(_mat(inp, nsrc, ncols, ldin)[src])
        return 0

    # ---- K14 ---------------------------------------------------------------------------------------------------
    def rvgp_potrf_f64(self, A, lda, n, flag, ws, wsb):
        Am = _mat(A, n, n, lda)
        flag.zero_()
        sym = torch.tril(Am) + torch.tril(Am, -1).t()             # only the lower triangle is read
        try:
            L = torch.linalg.cholesky(sym)
        except Exception:
            flag.fill_(1)
            L = torch.full_like(sym, float("nan"))
        Am.copy_(torch.tril(L) + torch.triu(torch.full_like(L, float("nan")), 1))     # upper triangle = scratch
        return 0

    def rvgp_trsm_f64(self, L, ldl, n, B, ldb, nrhs, trans, ws, scratch):
        assert scratch.numel() >= NB * nrhs
        Lm = torch.tril(_mat(L, n, n, ldl))
        Bm = _mat(B, n, nrhs, ldb)
        X = torch.linalg.solve_triangular(Lm.t() if trans else Lm, Bm.clone(), upper=bool(trans))
        Bm.copy_(X)
        return 0

    def rvgp_add_diag_f64(self, A, lda, n, v):
        _mat(A, n, n, lda).diagonal().add_(v)
        return 0

    def rvgp_logdiag_sum_f64(self, A, lda, n, out):
        out[0] = torch.log(_mat(A, n, n, lda).diagonal()).sum()
        return 0

    def rvgp_kdiag_f64(self, X, ldx, n, k, S, out):
        Xm = _mat(X, n, k, ldx)
        out[:n].copy_((Xm * Xm * S[:k][None, :]).sum(1))
        return 0

    def rvgp_axpy_f64(self, nrows, ncols, a, X, ldx, Y, ldy):
        _mat(Y, nrows, ncols, ldy).add_(a * _mat(X, nrows, ncols, ldx))
        return 0

    # ---- K17 ---------------------------------------------------------------------------------------------------
    def rvgp_rbf_from_dot_f64(self, m, n, P, ldp, xa2, xb2, variance, lengthscale, Kout, ldk):
        r2 = (xa2[:m, None] + xb2[None, :n] - 2.0 * _mat(P, m, n, ldp)) / lengthscale ** 2
        _mat(Kout, m, n, ldk).copy_(variance * torch.exp(-0.5 * r2))
        return 0

    def rvgp_rbf_adjoint_f64(self, m, n, Gbar, ldg, P, ldp, xa2, xb2, variance, lengthscale, Hout, ldh, rowsum, sums, ws, wsb):
        r2 = (xa2[:m, None] + xb2[None, :n] - 2.0 * _mat(P, m, n, ldp)) / lengthscale ** 2
        Hm = _mat(Gbar, m, n, ldg) * (variance * torch.exp(-0.5 * r2))
        sums[0] = Hm.sum()
        sums[1] = (Hm * r2).sum()
        if rowsum is not None:
            rowsum[:m].copy_(Hm.sum(1))
        if Hout is not None:
            _mat(Hout, m, n, ldh).copy_(Hm)
        return 0

    def rvgp_rbf_dx_f64(self, m, k, HX, ldhx, rowsum, XA, ldx, lengthscale, out, ldo):
        _mat(out, m, k, ldo).copy_((_mat(HX, m, k, ldhx) - rowsum[:m, None] * _mat(XA, m, k, ldx)) / lengthscale ** 2)
        return 0

    def rvgp_scale_shift_f64(self, nrows, ncols, alpha, beta, A, lda):
        Am = _mat(A, nrows, ncols, lda)
        Am.mul_(alpha)
        Am.diagonal().add_(beta)
        return 0

    # ---- K1 (N given) ------------------------------------------------------------------------------------------
    def rvgp_fps_f64(self, X, n, D, N, spacing, start_idx, perm, lambdas, count, ws, wsb):
        assert N > 0
        from sklearn.metrics import pairwise_distances
        Dm = pairwise_distances(_mat(X, n, D, D).numpy())
        ds = Dm[start_idx].copy()
        perm[0] = start_idx
        lambdas[0] = 0.0
        for i in range(1, N):
            idx = int(np.argmax(ds))
            perm[i] = idx
            lambdas[i] = ds[idx]
            ds = np.minimum(ds, Dm[idx])
        count[0] = N
        return 0


def install(monkeypatch):
    """Route every get_handle / to_device_f64 of the GP host modules to the CPU emulation.  Returns the FakeHandle."""
    import rvgp_b200._cabi as cabi
    import rvgp_b200.geometry as geo
    fh = FakeHandle()

    def fake_get_handle(device=None):
        return fh

    def fake_to_device(x, device=None):
        if isinstance(x, torch.Tensor):
            return x.to(dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))

    import importlib
    for modname in ("rvgp_b200._cabi", "rvgp_b200.gp", "rvgp_b200.gp_general", "rvgp_b200.kernels", "rvgp_b200.main",
                    "rvgp_b200.fps", "rvgp_b200.geometry"):
        mod = importlib.import_module(modname)
        if hasattr(mod, "get_handle"):
            monkeypatch.setattr(mod, "get_handle", fake_get_handle)
        if hasattr(mod, "to_device_f64"):
            monkeypatch.setattr(mod, "to_device_f64", fake_to_device)
    return fh
