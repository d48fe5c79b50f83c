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


# ---- eigensolver entry points (tests/test_krylov_host.py) -----------------------------------------------------------
def _fake_fill_uniform(self, nrows, ncols, A, lda, seed, col_offset, row_offset):
    g = np.random.default_rng([int(seed), int(col_offset), int(row_offset)])
    _mat(A, nrows, ncols, lda).copy_(torch.from_numpy(g.uniform(-1.0, 1.0, (int(nrows), int(ncols)))))
    return 0


def _fake_rot90(self, nnodes, ncols, V, ldv, out, ldo):
    Vm, Om = _mat(V, 2 * nnodes, ncols, ldv), _mat(out, 2 * nnodes, ncols, ldo)
    Om[0::2] = -Vm[1::2]
    Om[1::2] = Vm[0::2]
    return 0


def _fake_resid_sq(self, nrows, ncols, W, ldw, V, ldv, theta, out, ws):
    Wm = _mat(W, nrows, ncols, ldw)
    Vm = _mat(V, nrows, ncols, ldv) if V is not None else torch.ones_like(Wm)
    out[:ncols].copy_(((Wm - Vm * theta[:ncols][None, :]) ** 2).sum(0))
    return 0


def _fake_dgemm_lower(self, m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, split_k, ws):
    Am = _op(A, m, k, lda, a_kmajor == 1)
    Bm = _op(B, k, n, ldb, b_kmajor == 0)
    full = alpha * (Am @ Bm)
    # only tiles touching the lower triangle are defined: poison the strict upper part beyond the diagonal tiles
    _mat(C, m, n, ldc).copy_(torch.tril(full) + torch.triu(torch.full_like(full, float("nan")), 1))
    return 0


FakeHandle.rvgp_fill_uniform_f64 = _fake_fill_uniform
FakeHandle.rvgp_rot90_nodes_f64 = _fake_rot90


def _fake_pair_panel(self, nnodes, ncols, V, ldv, out, ldo):
    Vm, Om = _mat(V, 2 * nnodes, ncols, ldv), _mat(out, 2 * nnodes, 2 * ncols, ldo)
    Om[:, :ncols] = Vm
    Om[0::2, ncols:] = -Vm[1::2]
    Om[1::2, ncols:] = Vm[0::2]
    return 0


def _fake_pair_combine(self, nnodes, ncols, T, ldt, alpha, beta, W, ldw):
    Tm, Wm = _mat(T, 2 * nnodes, 2 * ncols, ldt), _mat(W, 2 * nnodes, ncols, ldw)
    v = Tm[:, :ncols].clone()
    v[0::2] -= Tm[1::2, ncols:]
    v[1::2] += Tm[0::2, ncols:]
    if beta != 0.0:
        Wm.mul_(beta).add_(alpha * v)
    else:
        Wm.copy_(alpha * v)
    return 0


FakeHandle.rvgp_pair_panel_f64 = _fake_pair_panel
FakeHandle.rvgp_pair_combine_f64 = _fake_pair_combine
FakeHandle.rvgp_resid_sq_f64 = _fake_resid_sq
FakeHandle.rvgp_dgemm_lower_f64 = _fake_dgemm_lower


class FakeBsr:
    """Duck-type of eigensolver.BsrMatrix over a SciPy matrix: the operator side of the C ABI (fused SpMM filter, products)."""
    mma = None
    row_offset = 0
    vals = 1

    def __init__(self, M, d=1):
        import scipy.sparse as sp
        self.M = sp.csr_matrix(M)
        self.d = int(d)
        self.nrows = self.M.shape[0]
        self.nbrows = self.nrows // self.d
        self.indptr = torch.zeros(1, dtype=torch.int32)          # only its .device is read
        self.col_degrees = 0

    def spmm_bytes(self, ncols, fused=False):
        return int(self.M.nnz * 12 + 8 * self.nrows * ncols * (3 if fused else 2))

    def matmat(self, X, out=None, h=None):
        Y = torch.from_numpy(self.M @ X.numpy())
        if out is None:
            return Y
        out.copy_(Y)
        return out

    def cheb_filter(self, Vp, w0, w1, ncols, degree, lo_spec, lo_cut, hi, h=None, w2=None):
        """The scaled three-term recurrence of rvgp_cheb_filter_f64 (csrc/spmm.cu), in place on the panel Vp."""
        if degree <= 0:
            return
        A = self.M
        V = Vp.numpy().copy()
        e, c = 0.5 * (hi - lo_cut), 0.5 * (hi + lo_cut)
        s1 = e / (lo_spec - c)
        tau, sig = 2.0 / s1, s1
        Y = (A @ V - c * V) * (s1 / e)
        X = V
        for _ in range(2, degree + 1):
            sn = 1.0 / (tau - sig)
            Y, X = (A @ Y - c * Y) * (2.0 * sn / e) - sig * sn * X, Y
            sig = sn
        Vp.copy_(torch.from_numpy(Y))
        self.col_degrees += degree * ncols


def install_eigensolver(monkeypatch):
    """CPU emulation for eigensolver.py / krylov.py: fake handle, and no-op CUDA events / synchronisation."""
    import importlib
    fh = FakeHandle()
    for modname in ("rvgp_b200.eigensolver", "rvgp_b200.krylov"):
        mod = importlib.import_module(modname)
        monkeypatch.setattr(mod, "get_handle", lambda device=None: fh)

    class _Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 0.0

    monkeypatch.setattr(torch.cuda, "Event", _Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    return fh


class FakeShardedBsr(FakeBsr):
    """Row-sharded duck-type of distributed.ShardedBsr for the gloo tests: this rank owns rows [r0, r1) of the global SciPy
    matrix; every product all-gathers the block vector (the real operator exchanges only the halo rows -- covered by
    tests/test_distributed_cpu.py::test_sharded_halo_spmm_gloo -- the eigensolvers above it see the same interface)."""
    spmm_kernel_name = "fake + gloo all-gather"

    def __init__(self, M, r0, r1, counts, comm, d=1):
        import scipy.sparse as sp
        super().__init__(sp.csr_matrix(M)[r0:r1], d=d)
        self.row_offset = int(r0)
        self.counts, self.comm = list(counts), comm

    def _full(self, X):
        return self.comm.allgather_rows(X.contiguous(), self.counts).numpy()

    def matmat(self, X, out=None, h=None):
        Y = torch.from_numpy(self.M @ self._full(X))
        if out is None:
            return Y
        out.copy_(Y)
        return out

    def cheb_filter(self, Vp, w0, w1, ncols, degree, lo_spec, lo_cut, hi, h=None, w2=None):
        if degree <= 0:
            return
        A = self.M
        e, c = 0.5 * (hi - lo_cut), 0.5 * (hi + lo_cut)
        s1 = e / (lo_spec - c)
        tau, sig = 2.0 / s1, s1
        X = Vp.clone()
        Y = torch.from_numpy((A @ self._full(X) - c * X.numpy()) * (s1 / e))
        for _ in range(2, degree + 1):
            sn = 1.0 / (tau - sig)
            Yn = torch.from_numpy((A @ self._full(Y) - c * Y.numpy()) * (2.0 * sn / e) - sig * sn * X.numpy())
            Y, X = Yn, Y
            sig = sn
        Vp.copy_(Y)
        self.col_degrees += degree * ncols
