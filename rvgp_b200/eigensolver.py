"""Smallest-k eigenpairs of the (connection) Laplacian on the GPU.

Replaces ``scipy.sparse.linalg.eigsh(A, k, which="SM")`` = ARPACK (reference RVGP/geometry.py:66-80).
Method: Chebyshev-filtered subspace iteration (block method, so the exactly paired eigenvalues of
``Lc`` on orientable 2-manifolds are handled as a block).  The hot loop is the fused block-SpMM
``rvgp_cheb_filter_f64`` (one launch per polynomial degree per column panel); orthogonalisation and
Rayleigh-Ritz use the FP64 ``rvgp_dgemm_f64`` kernel.  Only the m x m projected problems
(m = k + buffer <= a few hundred) are factorised on the host with LAPACK, in the same way ARPACK itself
hands its small tridiagonal/Hessenberg problems to LAPACK.

All device memory is torch tensors; all device math is librvgp_b200.so.
"""
import math
import time

import numpy as np
import torch

from ._cabi import get_handle, I64, U64
from . import _nvtx

_TPC = [None, 0, 0]


def _lapack_ctx(big=False):
    """Context for the host LAPACK sections (the m x m projected problems): a FEW BLAS threads.  Measured on the B200 box
    (16 host cores, tools/host_lapack_probe.py, profiles/r01d_host_lapack.txt): one outer iteration's Cholesky + triangular
    inverse + projection + eigh of a 640 x 640 problem takes 205 ms with OpenBLAS's default 16 threads, 51 ms with 4 and
    86 ms with 1 (torchrun's OMP_NUM_THREADS=1).  Threads per rank: host cores / (4 * ranks on the node), clamped to [1, 4];
    ``big`` (the >= 1000-dimensional eigh of the Krylov solver's checks): host cores / ranks, clamped to [1, 16].
    RVGP_HOST_THREADS overrides both."""
    import contextlib
    import os
    if _TPC[0] is None:
        try:
            from threadpoolctl import ThreadpoolController
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
            forced = int(os.environ.get("RVGP_HOST_THREADS", "0") or 0)
            cores = os.cpu_count() or 1
            n = forced or max(1, min(4, cores // (4 * max(1, lw))))
            nbig = forced or max(1, min(16, cores // max(1, lw)))
            _TPC[0], _TPC[1], _TPC[2] = ThreadpoolController(), n, nbig
        except Exception:
            _TPC[0] = False
    if not _TPC[0]:
        return contextlib.nullcontext()
    return _TPC[0].limit(limits=_TPC[2] if big else _TPC[1], user_api="blas")


class BsrMatrix:
    """Device block-CSR matrix with d x d FP64 blocks (vals None => unit-weight graph Laplacian pattern)."""

    def __init__(self, nbrows, d, indptr, indices, vals=None, max_offdiag=None):
        self.nbrows, self.d = int(nbrows), int(d)
        self.indptr, self.indices, self.vals = indptr, indices, vals
        self.nrows = self.nbrows * self.d
        self.nnzb = int(indices.numel())
        self.max_offdiag = max_offdiag   # max number of off-diagonal blocks in a row (Gershgorin)
        self.d_code, self.indices_k, self.vals_k = self.d, indices, vals     # what the kernels are given

    def compress_rot2(self, rtol=1e-12, h=None):
        """Switch the kernels to ROT2 storage (d == 2, every block a scaled 2x2 orthogonal matrix).  Returns True on
        success; the plain arrays stay available for everything else."""
        if self.d != 2 or self.vals is None or self.nnzb == 0:
            return False
        h = h or get_handle(self.indptr.device.index)
        dev = self.indptr.device
        ab = torch.empty((self.nnzb, 2), dtype=torch.float64, device=dev)
        idxf = torch.empty(self.nnzb, dtype=torch.int32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        h.call("rvgp_bsr_compress_rot2", I64(self.nnzb), self.vals, self.indices, ab, idxf, bad, float(rtol))
        if int(bad.item()) != 0:
            return False
        self.d_code, self.indices_k, self.vals_k = -2, idxf, ab
        return True

    def spmm_bytes(self, ncols, fused=False):
        """Algorithmic HBM bytes of one SpMM launch (SURVEY.md 8d / DESIGN.md K9)."""
        d = self.d
        mat = self.nnzb * ((8 * d * d if self.vals is not None else 0) + 4) + 4 * (self.nbrows + 1)
        return mat + 8 * self.nrows * ncols * (3 if fused else 2)

    def spmm(self, X, Y, alpha=1.0, beta=0.0, gamma=0.0, W=None, h=None):
        h = h or get_handle(X.device.index)
        ncols = X.shape[1]
        aligned = (ncols % 2 == 0 and X.stride(0) % 2 == 0 and Y.stride(0) % 2 == 0 and X.data_ptr() % 16 == 0
                   and Y.data_ptr() % 16 == 0 and (W is None or (W.stride(0) % 2 == 0 and W.data_ptr() % 16 == 0)))
        if self.d_code == -2 and aligned:
            dc, ix, vl = self.d_code, self.indices_k, self.vals_k
        else:
            dc, ix, vl = self.d, self.indices, self.vals
        if self._mma_ok(ncols, X, Y, W):
            mp = self.mma
            h.call("rvgp_bsr_spmm_mma_f64", self.nbrows, self.d, mp["kptr"], mp["kcols"], mp["afrag"],
                   X, I64(X.stride(0)), W, I64(W.stride(0) if W is not None else 0), Y, I64(Y.stride(0)),
                   int(ncols), float(alpha), float(beta), float(gamma))
            return Y
        h.call("rvgp_bsr_spmm_f64", self.nbrows, dc, self.indptr, ix, vl,
               X, I64(X.stride(0)), W, I64(W.stride(0) if W is not None else 0), Y, I64(Y.stride(0)),
               int(ncols), float(alpha), float(beta), float(gamma))
        return Y

    # ---- FP64-MMA row-group variant (K9 v4, spmm_mma.cu) ----------------------------------------------------
    mma = None
    mma_rowmajor = False      # experiment: also route single products / row-major filters through the (slower) row-major MMA kernel

    def build_mma_plan(self, h=None, compact=True, keep_full=False):
        """k-step plan of the MMA SpMM: groups of 8/d block rows x k-steps of 4/d union columns.  For d == 2 the
        fragments are compacted to (a, b) + sign bits when every block is a scaled rotation / reflection (always true for
        the connection Laplacian); ``keep_full`` keeps the 32-double fragments as well (row-major kernel)."""
        if self.d not in (1, 2) or self.nbrows == 0:
            return None
        if getattr(self, "_mma_plan", None) is not None:
            return self._mma_plan
        h = h or get_handle(self.indptr.device.index)
        dev = self.indptr.device
        R, T = 8 // self.d, 4 // self.d
        mp = self.build_merge_plan(R, h=h)
        gptr = mp["gptr"]
        ns = (gptr[1:] - gptr[:-1] + (T - 1)) // T
        kptr = torch.zeros(gptr.numel(), dtype=torch.int32, device=dev)
        kptr[1:] = torch.cumsum(ns, 0).to(torch.int32)
        nk = int(kptr[-1].item())
        kcols = torch.empty(max(1, nk * T), dtype=torch.int32, device=dev)
        afrag = torch.empty(max(1, nk * 32), dtype=torch.float64, device=dev)
        h.call("rvgp_bsr_mma_pack", self.nbrows, self.d, self.indptr, self.indices, self.vals, gptr, mp["uent"], kptr,
               kcols, afrag)
        plan = dict(kptr=kptr, kcols=kcols, afrag=afrag, ksteps=nk, reuse=mp["reuse"],
                    fill=self.nnzb / max(1, nk * R * T), rotc=0, kcols_c=None, afrag_c=None)
        self.__dict__.get("_mplans", {}).pop(R, None)      # the union lists are only needed for packing
        if compact and self.d == 2 and self.vals is not None and nk > 0:
            kcols_c = kcols.clone()
            afrag_c = torch.empty(nk * 16, dtype=torch.float64, device=dev)
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            h.call("rvgp_bsr_mma_rotc", I64(nk), afrag, kcols_c, afrag_c, bad, 1e-12)
            if int(bad.item()) == 0:
                plan.update(rotc=1, kcols_c=kcols_c, afrag_c=afrag_c)
                if not keep_full:
                    plan["afrag"] = None                   # C4: 0.86 GB
        self._mma_plan = plan
        return plan

    mma_pattern = None        # PATTERN plan of the scalar unit-weight Laplacian on the native MMA kernel (spmm_mma.cu, AMODE 2)

    def enable_mma_pattern(self, on=True, h=None):
        """Scalar unit-weight Laplacian (d == 1, vals None): run the Chebyshev filter on the FP64-MMA native kernel as L (x) I_2.
        A row-major (n x B) panel is already in that kernel's layout, and no matrix values are streamed (k-step column words carry
        the row masks; the diagonal is row length - 1)."""
        if not on or self.d != 1 or self.vals is not None or self.nbrows < 64:
            self.mma_pattern = None
            return None
        if self.mma_pattern is None:
            h = h or get_handle(self.indptr.device.index)
            dev = self.indptr.device
            mp = self.build_merge_plan(4, h=h)
            gptr = mp["gptr"]
            ns = (gptr[1:] - gptr[:-1] + 1) // 2
            kptr = torch.zeros(gptr.numel(), dtype=torch.int32, device=dev)
            kptr[1:] = torch.cumsum(ns, 0).to(torch.int32)
            nk = int(kptr[-1].item())
            kcols = torch.empty(max(1, nk * 2), dtype=torch.int32, device=dev)
            deg = torch.empty(self.nbrows, dtype=torch.int32, device=dev)
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            h.call("rvgp_bsr_mma_pack_pattern", self.nbrows, self.indptr, gptr, mp["uent"], kptr, kcols, deg, bad)
            self.__dict__.get("_mplans", {}).pop(4, None)
            if int(bad.item()) != 0:
                return None
            self.mma_pattern = dict(kptr=kptr, kcols=kcols, deg=deg, ksteps=nk, reuse=mp["reuse"])
        return self.mma_pattern

    def _mma_pattern_ok(self, ncols, Vp, w0, w1):
        if self.mma_pattern is None or ncols % 32:
            return False
        return all(t.stride(0) % 2 == 0 and t.data_ptr() % 16 == 0 for t in (Vp, w0, w1))

    def enable_mma(self, on=True, h=None, rowmajor=False):
        """Route cheb_filter (d == 2: node-contiguous panels) through the FP64-MMA kernel whenever shapes / alignment
        allow it; ``rowmajor`` additionally routes single products through the row-major MMA kernel (experiment)."""
        if on and rowmajor and getattr(self, "_mma_plan", None) is not None and self._mma_plan["afrag"] is None:
            self._mma_plan = None
        self.mma = self.build_mma_plan(h=h, keep_full=rowmajor) if on else None
        self.mma_rowmajor = bool(on and rowmajor and self.mma is not None)
        return self.mma

    def _mma_ok(self, ncols, *tensors):
        if self.mma is None or not self.mma_rowmajor or ncols % 16:
            return False
        return all(t is None or (t.stride(0) % 4 == 0 and t.data_ptr() % 32 == 0) for t in tensors)

    def _mma_native_ok(self, ncols, Vp, w0, w1, w2):
        if self.mma is None or self.d != 2 or w2 is None or ncols % 16:
            return False
        if Vp.stride(0) % 2 or Vp.data_ptr() % 16:
            return False
        return all(t.stride(0) == ncols and t.data_ptr() % 16 == 0 and t.shape[1] == ncols for t in (w0, w1, w2))

    def to_native(self, X, out=None, h=None):
        """Row-major (2 n x b) block -> node-contiguous (n x 2b) panel of the native MMA kernel."""
        h = h or get_handle(X.device.index)
        b = X.shape[1]
        if out is None:
            out = torch.empty((self.nbrows, 2 * b), dtype=torch.float64, device=X.device)
        h.call("rvgp_panel_native_f64", 1, self.nbrows, int(b), X, I64(X.stride(0)), out, I64(out.stride(0)))
        return out

    def from_native(self, Xn, out=None, h=None):
        h = h or get_handle(Xn.device.index)
        b = Xn.shape[1] // 2
        if out is None:
            out = torch.empty((self.nrows, b), dtype=torch.float64, device=Xn.device)
        h.call("rvgp_panel_native_f64", 0, self.nbrows, int(b), out, I64(out.stride(0)), Xn, I64(Xn.stride(0)))
        return out

    def spmm_native(self, Xn, Yn, alpha=1.0, beta=0.0, gamma=0.0, Wn=None, reverse=False, h=None):
        """Y = alpha A X + beta X + gamma W on node-contiguous panels (d == 2, MMA plan enabled, alpha != 0)."""
        h = h or get_handle(Xn.device.index)
        mp = self.mma
        rotc = mp["rotc"]
        h.call("rvgp_bsr_spmm_mma_native_f64", self.nbrows, mp["kptr"], mp["kcols_c"] if rotc else mp["kcols"],
               mp["afrag_c"] if rotc else mp["afrag"], int(rotc), Xn, I64(Xn.stride(0)), Wn,
               I64(Wn.stride(0) if Wn is not None else 0), Yn, I64(Yn.stride(0)), int(Xn.shape[1] // 2), float(alpha),
               float(beta), float(gamma), int(bool(reverse)))
        return Yn

    def spmm_pattern(self, X, Y, alpha=1.0, beta=0.0, gamma=0.0, W=None, h=None):
        """Y = alpha L X + beta X + gamma W for ROW-MAJOR scalar panels through the PATTERN plan (ncols % 32 == 0)."""
        h = h or get_handle(X.device.index)
        mp = self.mma_pattern
        h.call("rvgp_bsr_spmm_mma_native_f64", self.nbrows, mp["kptr"], mp["kcols"], mp["deg"], 2, X, I64(X.stride(0)), W,
               I64(W.stride(0) if W is not None else 0), Y, I64(Y.stride(0)), int(X.shape[1] // 2), float(alpha), float(beta),
               float(gamma), 0)
        return Y

    # ---- row-group merge plan (union column lists; input of the MMA k-step plan) --------------------------------
    def build_merge_plan(self, R=4, h=None):
        """Union column lists of groups of R consecutive block rows (rvgp_bsr_merge_plan)."""
        plans = self.__dict__.setdefault("_mplans", {})
        if R in plans:
            return plans[R]
        h = h or get_handle(self.indptr.device.index)
        dev = self.indptr.device
        ngroups = (self.nbrows + R - 1) // R
        wsb = h.query("rvgp_bsr_merge_plan_workspace_bytes", int(self.nbrows), int(R))
        ws = torch.empty(max(8, wsb), dtype=torch.uint8, device=dev)
        gptr = torch.empty(ngroups + 1, dtype=torch.int32, device=dev)
        h.call("rvgp_bsr_merge_plan", self.nbrows, self.indptr, self.indices, int(R), gptr, None, ws, I64(wsb))
        total = int(gptr[-1].item())
        uent = torch.empty((max(1, total), 2), dtype=torch.int32, device=dev)
        h.call("rvgp_bsr_merge_plan", self.nbrows, self.indptr, self.indices, int(R), gptr, uent, ws, I64(wsb))
        plans[R] = dict(R=int(R), gptr=gptr, uent=uent, total=total, reuse=self.nnzb / max(1, total))
        return plans[R]

    row_offset = 0          # global row of local row 0 (non-zero only for row-sharded operators)

    def cheb_filter(self, Vp, w0, w1, ncols, degree, lo_spec, lo_cut, hi, h=None, w2=None):
        """Degree-`degree` Chebyshev filter of the panel Vp in place; the whole recurrence runs inside one C call.
        With a third contiguous work panel ``w2`` and an MMA plan (d == 2) the recurrence runs on node-contiguous panels."""
        h = h or get_handle(Vp.device.index)
        aligned = (ncols % 2 == 0 and Vp.stride(0) % 2 == 0 and w0.stride(0) % 2 == 0 and Vp.data_ptr() % 16 == 0
                   and w0.data_ptr() % 16 == 0 and w1.data_ptr() % 16 == 0)
        if self._mma_native_ok(ncols, Vp, w0, w1, w2):
            mp = self.mma
            rotc = mp["rotc"]
            h.call("rvgp_cheb_filter_mma_f64", self.nbrows, self.d, mp["kptr"], mp["kcols_c"] if rotc else mp["kcols"],
                   mp["afrag_c"] if rotc else mp["afrag"], int(rotc), Vp, I64(Vp.stride(0)), w0, w1, w2, I64(ncols),
                   int(ncols), int(degree), float(lo_spec), float(lo_cut), float(hi))
            return
        if self._mma_pattern_ok(ncols, Vp, w0, w1):
            mp = self.mma_pattern
            h.call("rvgp_cheb_filter_mma_f64", self.nbrows, 1, mp["kptr"], mp["kcols"], mp["deg"], 2, Vp, I64(Vp.stride(0)), w0, w1,
                   None, I64(w0.stride(0)), int(ncols), int(degree), float(lo_spec), float(lo_cut), float(hi))
            return
        if self.d_code == -2 and aligned:
            dc, ix, vl = self.d_code, self.indices_k, self.vals_k
        else:
            dc, ix, vl = self.d, self.indices, self.vals
        if self._mma_ok(ncols, Vp, w0, w1):
            mp = self.mma
            h.call("rvgp_cheb_filter_mma_f64", self.nbrows, self.d, mp["kptr"], mp["kcols"], mp["afrag"], 0,
                   Vp, I64(Vp.stride(0)), w0, w1, None, I64(w0.stride(0)), int(ncols), int(degree), float(lo_spec),
                   float(lo_cut), float(hi))
            return
        h.call("rvgp_cheb_filter_f64", self.nbrows, dc, self.indptr, ix, vl, Vp, I64(Vp.stride(0)),
               w0, w1, I64(w0.stride(0)), int(ncols), int(degree), float(lo_spec), float(lo_cut), float(hi))

    def matmat(self, X, out=None, h=None):
        """out = A @ X for any number of columns (panels of <= 64)."""
        if out is None:
            out = torch.empty_like(X)
        for c0 in range(0, X.shape[1], 64):
            c1 = min(X.shape[1], c0 + 64)
            self.spmm(X[:, c0:c1], out[:, c0:c1], h=h)
        return out


def _dgemm(h, m, n, k, A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, alpha=1.0, scale_k=None, split_k=1, ws=None):
    h.call("rvgp_dgemm_f64", int(m), int(n), I64(k), float(alpha), A, I64(lda), int(a_kmajor), B, I64(ldb),
           int(b_kmajor), scale_k, C, I64(ldc), int(split_k), ws)


class _Dense:
    """Tall-skinny FP64 products on (N x m) block vectors."""

    def __init__(self, h, N, m, device, comm=None):
        self.h, self.N, self.m, self.comm = h, N, m, comm
        tiles = math.ceil(m / 128) ** 2
        self.split = max(1, min(64, (2 * h.sm_count) // tiles, N // 2048 if N >= 4096 else 1))
        self.red_rows = N
        self.ws = torch.empty(self.split * m * m, dtype=torch.float64, device=device)
        nred = h.query("rvgp_coldot_workspace_bytes", I64(N), int(m)) // 8
        self.red_ws = torch.empty(max(1, nred), dtype=torch.float64, device=device)
        self.small = torch.empty(m, dtype=torch.float64, device=device)

    def gram(self, V, W, out, sym=False):
        """out (m1 x m2) = V^T W.  sym=True: the result is symmetric, only its lower-triangle tiles are computed
        (read it back with ``sym_to_host``)."""
        m1, m2 = V.shape[1], W.shape[1]
        if sym:
            self.h.call("rvgp_dgemm_lower_f64", int(m1), int(m2), I64(self.N), 1.0, V, I64(V.stride(0)), 0, W,
                        I64(W.stride(0)), 0, out, I64(out.stride(0)), int(self.split), self.ws)
        else:
            _dgemm(self.h, m1, m2, self.N, V, V.stride(0), 0, W, W.stride(0), 0, out, out.stride(0),
                   split_k=self.split, ws=self.ws)
        if self.comm is not None:
            self.comm.allreduce_(out)            # row-sharded: sum of the ranks' partial Gram matrices
        return out

    @staticmethod
    def sym_to_host(Gd):
        G = np.tril(Gd.cpu().numpy())
        return G + np.tril(G, -1).T

    def apply(self, V, Cm, out):
        """out (N x m2) = V (N x m1) @ Cm (m1 x m2)."""
        m1, m2 = Cm.shape
        _dgemm(self.h, self.N, m2, m1, V, V.stride(0), 1, Cm, Cm.stride(0), 0, out, out.stride(0))
        return out

    def coldot(self, A, B):
        out = self.small[: A.shape[1]]
        self.h.call("rvgp_coldot_f64", I64(self.N), int(A.shape[1]), A, I64(A.stride(0)), B, I64(B.stride(0)),
                    out, self.red_ws)
        if self.comm is not None:
            self.comm.allreduce_(out)
        return out

    def resid_sq(self, W, V, theta):
        out = self.small[: V.shape[1]]
        self.h.call("rvgp_resid_sq_f64", I64(self.N), int(V.shape[1]), W, I64(W.stride(0)), V, I64(V.stride(0)),
                    theta, out, self.red_ws)
        if self.comm is not None:
            self.comm.allreduce_(out)
        return out

    def colscale(self, A, s):
        self.h.call("rvgp_colscale_f64", I64(self.N), int(A.shape[1]), A, I64(A.stride(0)), s)


def smallest_eigenpairs(A, k, upper_bound, lower_bound=0.0, tol=1e-12, nex=None, seed=0, max_outer=80,
                        cond_max=1e6, deg0=20, panel=None, stats=None, verbose=False, comm=None,
                        refine_bound=True, init_fn=None, hi_override=None, m_exact=None, orthonormal_start=False):
    """Smallest k eigenpairs of the symmetric PSD BsrMatrix ``A``.

    upper_bound: a rigorous upper bound of the spectrum (2 * max degree for (connection) Laplacians).
    tol: residual tolerance relative to upper_bound, ||A x - theta x|| <= tol * upper_bound.
    Returns (evals (k,) float64 cuda, evecs (N, k) float64 cuda, unit-norm columns, ascending).
    """
    dev = A.indptr.device
    h = get_handle(dev.index)
    N = A.nrows                       # LOCAL rows (== global rows on a single GPU)
    Nglob = N
    if comm is not None and comm.world > 1:
        t = torch.tensor([N], dtype=torch.int64, device=dev)
        comm.allreduce_(t)
        Nglob = int(t.item())
    else:
        comm = None
    k = int(min(k, Nglob))
    if nex is None:
        nex = max(16, int(math.ceil(0.2 * k)))
    m = min(Nglob, k + nex)
    if panel is None:
        panel = 64 if m >= 128 else 32
        import os
        if getattr(A, "vals", 1) is None and m >= 512 and (comm is None or os.environ.get("RVGP_SHARDED_L_PANEL") == "128"):
            # scalar pattern-mode Laplacian: 0.57 vs 0.52 of HBM peak with 128-column panels (measured on one GPU).  Row-sharded
            # runs keep 64 until measured: RVGP_SHARDED_L_PANEL=128 opts in (8 GPUs: 9 791 steps of 0.09 ms are latency-bound,
            # half as many 128-column steps should cost less; DESIGN.md "next")
            panel = 128
    if m_exact is not None:
        m = int(min(Nglob, max(k, m_exact)))     # block handed over by krylov.krylov_eigenpairs: exactly its columns
    elif m < Nglob:
        m = min(Nglob, ((m + panel - 1) // panel) * panel)
    hi = float(upper_bound)
    if hi_override is not None:
        hi = float(hi_override)
    elif refine_bound and Nglob > 4 * m:
        # Gershgorin (2 * max degree) overestimates lambda_max of kNN-graph Laplacians by ~1.5x; the filter degree scales
        # with sqrt(hi), so a 24-step Lanczos bound pays for itself many times over
        hi = min(hi, 1.01 * lanczos_upper_bound(A, comm=comm, h=h))
    lo_spec = float(lower_bound)
    tol_abs = tol * float(upper_bound)
    import os as _os
    cond_max = float(_os.environ.get("RVGP_EIG_COND_MAX", cond_max))     # policy knob (profiles/r01d_eig_policy_emulation.txt)
    st = stats if stats is not None else {}
    st.update(dict(N=N, k=k, m=m, panel=panel, spmm_launches=0, filter_launches=0, filter_col_degrees=0, outer=0,
                   t_filter=0.0, t_dense=0.0, t_host=0.0, spmm_bytes_fused=int(A.spmm_bytes(panel, fused=True)),
                   spmm_bytes_plain=int(A.spmm_bytes(panel, fused=False)), d=A.d, world=(comm.world if comm else 1),
                   hi=hi, hi_gershgorin=float(upper_bound)))

    B1 = torch.empty((N, m), dtype=torch.float64, device=dev)
    B2 = torch.empty((N, m), dtype=torch.float64, device=dev)
    w0 = torch.empty((N, panel), dtype=torch.float64, device=dev)
    w1 = torch.empty((N, panel), dtype=torch.float64, device=dev)
    # third work panel: the MMA filter (d == 2) rotates three node-contiguous panels (BsrMatrix.cheb_filter)
    w2 = None
    if isinstance(A, BsrMatrix) and A.mma is not None and A.d == 2:
        w2 = torch.empty((N, panel), dtype=torch.float64, device=dev)
    st["spmm_kernel"] = "mma_native" if w2 is not None else ("mma_native_pattern" if getattr(A, "mma_pattern", None) is not None else "gather")
    Gd = torch.empty((m, m), dtype=torch.float64, device=dev)
    Hd = torch.empty((m, m), dtype=torch.float64, device=dev)
    Cd = torch.empty((m, m), dtype=torch.float64, device=dev)
    theta_d = torch.empty(m, dtype=torch.float64, device=dev)
    dense = _Dense(h, N, m, dev, comm)

    V, W = B1, B2
    h.call("rvgp_fill_uniform_f64", I64(N), int(m), V, I64(V.stride(0)), U64(seed), I64(0), I64(A.row_offset))
    if init_fn is not None:
        init_fn(V)            # overwrite (some of) the random columns with an informed guess of the invariant subspace

    deg = np.full(m, deg0, dtype=np.int64)
    a_cut = lo_spec + 0.3 * (hi - lo_spec)
    theta = res = None
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev2 = torch.cuda.Event(enable_timing=True)

    for it in range(max_outer):
        ev0.record()
        _nvtx.push("eig:filter")
        # ---- polynomial filter: one fused SpMM launch per degree per panel -------------------------
        for p0 in range(0, m, panel):
            p1 = min(m, p0 + panel)
            dg = int(deg[p0:p1].max())
            if dg <= 0:
                continue
            if w2 is not None and p1 - p0 == panel:
                A.cheb_filter(V[:, p0:p1], w0, w1, p1 - p0, dg, lo_spec, float(a_cut), hi, h=h, w2=w2)
            else:
                A.cheb_filter(V[:, p0:p1], w0, w1, p1 - p0, dg, lo_spec, float(a_cut), hi, h=h)
            st["spmm_launches"] += dg
            st["filter_launches"] += dg
            st["filter_col_degrees"] += dg * (p1 - p0)
        ev1.record()
        _nvtx.pop()
        _nvtx.push("eig:orthonormalise+rayleigh_ritz")
        # ---- orthonormalise: column scaling + Cholesky-QR, then Rayleigh-Ritz with the second
        #      Cholesky folded into the projected problem ---------------------------------------------
        # CholeskyQR passes until one succeeds WITHOUT a diagonal shift (shifted CholeskyQR3: a failed /
        # shifted pass still reduces cond(V) by orders of magnitude, so the next pass is safe).
        # orthonormal_start (hand-over from krylov.py): the start block is already orthonormal to ~1e-13 and is not filtered in
        # the first sweep, so the explicit pass is skipped there -- the Rayleigh-Ritz below solves the GENERALISED problem with
        # G = V^T V anyway (2 Gram products and 1 block multiply less: ~40 % of the hand-over)
        for _pass in range(0 if (orthonormal_start and it == 0) else 4):
            nrm = dense.coldot(V, V)
            inv = torch.rsqrt(nrm)
            dense.colscale(V, inv)
            Gd.zero_()
            dense.gram(V, V, Gd, sym=True)
            G = dense.sym_to_host(Gd)
            t0 = time.perf_counter()
            with _lapack_ctx():
                Rinv, shifted = _cholqr_factor(G, dev)
            st["t_host"] += time.perf_counter() - t0
            Cd.copy_(torch.from_numpy(np.ascontiguousarray(Rinv)))
            dense.apply(V, Cd, W)                          # W = V R^-1   (nearly orthonormal)
            V, W = W, V
            st["cholqr_passes"] = st.get("cholqr_passes", 0) + 1
            if not shifted:
                break
        A.matmat(V, out=W, h=h)                            # W = A V
        st["spmm_launches"] += math.ceil(m / 64)
        Gd.zero_(); Hd.zero_()
        dense.gram(V, V, Gd, sym=True)
        dense.gram(V, W, Hd, sym=True)               # V^T A V is symmetric
        G = dense.sym_to_host(Gd)
        Hm = dense.sym_to_host(Hd)
        t0 = time.perf_counter()
        with _lapack_ctx():
            theta, Cm = _rayleigh_ritz_small(G, Hm, dev)
        st["t_host"] += time.perf_counter() - t0
        Cd.copy_(torch.from_numpy(np.ascontiguousarray(Cm)))
        theta_d.copy_(torch.from_numpy(theta))
        dense.apply(V, Cd, W)                              # Ritz vectors
        V, W = W, V
        A.matmat(V, out=W, h=h)                            # A * Ritz vectors, for true residuals
        st["spmm_launches"] += math.ceil(m / 64)
        res = torch.sqrt(dense.resid_sq(W, V, theta_d)).cpu().numpy()
        ev2.record()
        _nvtx.pop()
        torch.cuda.synchronize(dev)
        st["t_filter"] += ev0.elapsed_time(ev1) * 1e-3
        st["t_dense"] += ev1.elapsed_time(ev2) * 1e-3
        st["outer"] = it + 1

        nconv = int((res[:k] <= tol_abs).sum())
        a_cut = float(theta[-1]) if m < Nglob else a_cut
        if verbose:
            print("  [eig] it %d  cut=%.6g  conv=%d/%d  maxres=%.3e  theta_k=%.9g" %
                  (it, a_cut, nconv, k, res[:k].max(), theta[k - 1]))
        if nconv == k or m >= Nglob:
            break
        a_cut = max(a_cut, lo_spec + 1e-12 * (hi - lo_spec) + theta[k - 1] * (1 + 1e-9))
        deg = _next_degrees(theta, res, k, tol_abs, a_cut, hi, lo_spec, cond_max)

    if not isinstance(A, BsrMatrix) and hasattr(A, "spmm_kernel_name"):
        st["spmm_kernel"] = A.spmm_kernel_name
    st["residual_max"] = float(res[:k].max())
    st["converged"] = bool((res[:k] <= tol_abs).all())
    st["tol_abs"] = tol_abs
    _warn_unconverged(st, max_outer)
    evals = theta_d[:k].clone()
    evecs = V[:, :k].contiguous()     # copy out so the (N x m) work buffers can be freed
    return evals, evecs


class EigenConvergenceWarning(RuntimeWarning):
    """The block eigensolver stopped at max_outer without meeting the residual tolerance (scipy's eigsh raises
    ArpackNoConvergence in the same situation, geometry.py:73)."""


def _warn_unconverged(st, max_outer):
    if not st["converged"]:
        import warnings
        warnings.warn("eigensolver stopped after %d outer iterations with max residual %.3e > tolerance %.3e; the returned "
                      "pairs are NOT converged" % (st.get("outer", max_outer), st["residual_max"], st["tol_abs"]),
                      EigenConvergenceWarning, stacklevel=3)


def lanczos_upper_bound(A, steps=24, seed=12345, comm=None, h=None):
    """Safeguarded upper bound of the spectrum of the symmetric operator A from a `steps`-step Lanczos run
    (Zhou & Li 2011: theta_max + |beta_k|).  One single-column SpMM, two dot products and two axpys per step."""
    dev = A.indptr.device
    h = h or get_handle(dev.index)
    N = A.nrows
    dense = _Dense(h, N, 1, dev, comm)
    v = torch.empty((N, 1), dtype=torch.float64, device=dev)
    vp = torch.zeros((N, 1), dtype=torch.float64, device=dev)
    w = torch.empty((N, 1), dtype=torch.float64, device=dev)
    h.call("rvgp_fill_uniform_f64", I64(N), 1, v, I64(1), U64(seed), I64(0), I64(A.row_offset))
    nrm = math.sqrt(float(dense.coldot(v, v).item()))
    dense.colscale(v, torch.full((1,), 1.0 / nrm, dtype=torch.float64, device=dev))
    alphas, betas = [], []
    beta = 0.0
    for j in range(steps):
        A.matmat(v, out=w, h=h)
        alpha = float(dense.coldot(w, v).item())
        h.call("rvgp_axpy_f64", I64(N), 1, -alpha, v, I64(1), w, I64(1))
        if j > 0:
            h.call("rvgp_axpy_f64", I64(N), 1, -beta, vp, I64(1), w, I64(1))
        beta = math.sqrt(max(float(dense.coldot(w, w).item()), 0.0))
        alphas.append(alpha)
        betas.append(beta)
        if beta < 1e-14 * max(1.0, abs(alpha)):
            break
        vp, v, w = v, w, vp
        dense.colscale(v, torch.full((1,), 1.0 / beta, dtype=torch.float64, device=dev))
    T = np.diag(alphas) + np.diag(betas[:-1], 1) + np.diag(betas[:-1], -1)
    theta = np.linalg.eigvalsh(T)
    return float(theta[-1] + abs(betas[-1]))



# ---- the m x m projected problems -------------------------------------------------------------------------------------------
# Small blocks (m < 192) are factorised with host LAPACK (a few BLAS threads, _lapack_ctx); larger ones on the GPU through
# torch.linalg (cuSOLVER): measured on the B200 box (tools/eigh_probe.py, profiles/r02f_eigh_probe.txt) a 1280 x 1280 real eigh
# takes 20 ms there against 86 ms on 16 host cores and 266 ms on one -- and under torchrun every rank has only a few cores, so at
# 8 GPUs the host factorisations were the serial floor of the eigensolver.  RVGP_HOST_EIGH=1 forces the host path.
import os as _os_mod
_HOST_EIGH = _os_mod.environ.get("RVGP_HOST_EIGH", "0") == "1"
_DEVICE_MIN = 192


def _on_device(m, dev):
    return (not _HOST_EIGH) and m >= _DEVICE_MIN and dev is not None and dev.type == "cuda"


def _cholqr_factor(G, dev=None):
    """Upper Cholesky factor R of the Hermitian Gram matrix G (G = R^H R; shifted on breakdown) and R^-1.
    Returns (Rinv (host), shifted)."""
    m = G.shape[0]
    if _on_device(m, dev):
        Gt = torch.from_numpy(np.ascontiguousarray(G)).to(dev)
        L, info = torch.linalg.cholesky_ex(Gt)
        shifted = False
        if int(info.item()) != 0:
            shifted = True
            shift = 1e-13 * m
            eye = torch.eye(m, dtype=Gt.dtype, device=dev)
            while True:
                L, info = torch.linalg.cholesky_ex(Gt + shift * eye)
                if int(info.item()) == 0:
                    break
                shift *= 100.0
                if shift > 1.0:
                    raise np.linalg.LinAlgError("Gram matrix is not positive definite even with a unit shift")
        eye = torch.eye(m, dtype=Gt.dtype, device=dev)
        Rinv = torch.linalg.solve_triangular(L.mH, eye, upper=True)          # R = L^H
        return Rinv.cpu().numpy(), shifted
    R, shifted = (_chol_upper_shifted_c(G) if np.iscomplexobj(G) else _chol_upper_shifted(G))
    return _tri_inv_upper(R), shifted


def _rayleigh_ritz_small(G, Hm, dev=None):
    """Generalised Hermitian eigenproblem H y = theta G y of the (nearly orthonormal) block: R = chol(G) upper,
    eigh(R^-H H R^-1), coefficients R^-1 Y.  Returns (theta ascending (host), Cm (host))."""
    m = G.shape[0]
    if _on_device(m, dev):
        Gt = torch.from_numpy(np.ascontiguousarray(G)).to(dev)
        Ht = torch.from_numpy(np.ascontiguousarray(Hm)).to(dev)
        L = torch.linalg.cholesky(Gt)                                          # G = L L^H, R = L^H
        X = torch.linalg.solve_triangular(L, Ht, upper=False)                   # L^-1 H
        Hp = torch.linalg.solve_triangular(L, X.mH, upper=False).mH             # L^-1 H L^-H = R^-H H R^-1
        Hp = 0.5 * (Hp + Hp.mH)
        theta, Y = torch.linalg.eigh(Hp)
        Cm = torch.linalg.solve_triangular(L.mH, Y, upper=True)                 # R^-1 Y
        return theta.cpu().numpy(), Cm.cpu().numpy()
    R2 = np.linalg.cholesky(G).conj().T
    R2inv = _tri_inv_upper(R2)
    Hp = R2inv.conj().T @ Hm @ R2inv
    Hp = 0.5 * (Hp + Hp.conj().T)
    theta, Y = np.linalg.eigh(Hp)
    return theta, R2inv @ Y


def _chol_upper_shifted(G):
    """Upper Cholesky factor of the (unit-diagonal) Gram matrix; on breakdown retry with a growing diagonal
    shift (Fukaya et al., shifted CholeskyQR).  Returns (R, shifted)."""
    try:
        return np.linalg.cholesky(G).T, False
    except np.linalg.LinAlgError:
        pass
    m = G.shape[0]
    shift = 1e-13 * m
    while True:
        try:
            return np.linalg.cholesky(G + shift * np.eye(m)).T, True
        except np.linalg.LinAlgError:
            shift *= 100.0
            if shift > 1.0:
                raise


def _tri_inv_upper(R):
    import scipy.linalg
    return scipy.linalg.solve_triangular(R, np.eye(R.shape[0]), lower=False)


def _next_degrees(theta, res, k, tol_abs, a_cut, hi, lo_spec, cond_max, deg_cap=6000):
    """Per-column Chebyshev degree for the next sweep (ChASE-style degree optimisation).

    Column i with Ritz value theta_i is amplified by exp(g_i) per degree relative to the damped interval
    [a_cut, hi], g_i = acosh((c - theta_i)/e).  The degree is what brings its residual a decade below
    tol, capped so the filtered block stays numerically full rank (relative growth of the lowest
    eigen-direction <= cond_max)."""
    e = 0.5 * (hi - a_cut)
    c = 0.5 * (hi + a_cut)
    g = np.arccosh(np.maximum((c - theta) / e, 1.0))
    g0 = math.acosh(max((c - lo_spec) / e, 1.0))
    with np.errstate(divide="ignore"):
        need = np.where(res > tol_abs,
                        np.maximum(np.log(np.maximum(10.0 * res / tol_abs, 1.0)) / np.maximum(g, 1e-12), 8.0), 0.0)
    cap = math.log(cond_max) / np.maximum(g0 - g, 1e-12)
    d = np.ceil(np.minimum(need * 1.05 + 1.0, cap))
    d[res <= tol_abs] = 0
    if k < len(d):
        d[k:] = np.minimum(d[k:], d[:k].max())
    return np.minimum(d, deg_cap).astype(np.int64)


# =====================================================================================================================
# Paired mode: d = 2 connection Laplacian whose blocks are ALL scaled rotations (after geometry.orient_gauges_device).
# Such a matrix commutes with the per-node quarter turn J, i.e. it is an n x n complex-Hermitian operator on
# z_i = x_i + i y_i (real storage unchanged: column c of a (2n x b) block IS one complex vector).  Every eigenvalue is an
# exactly double eigenvalue of the real matrix (eigenvectors v and J v), so the block method needs HALF the columns:
#   filter            unchanged (the SpMM is complex-linear as it stands), on b/2 columns
#   Gram  V^H W       = V^T W + i (J V)^T W        two real lower-triangle Grams of half the width  (1/2 the flops)
#   apply V C         = V Re(C) + (J V) Im(C)      two real dgemms of half the width                  (1/2 the flops)
#   projected problem   complex Hermitian (m/2) x (m/2) on the host (LAPACK zheevd / zpotrf), as in the real mode
# The caller expands the result: eigenvalue theta_j -> (theta_j, theta_j), eigenvectors (v_j, J v_j).
# =====================================================================================================================
def _herm_from_lower(Gr_d, Gi_d):
    """Complex Hermitian matrix from the lower-triangle tiles of its real (symmetric) and imaginary (antisymmetric) parts."""
    Gr = np.tril(Gr_d.cpu().numpy())
    Gr = Gr + np.tril(Gr, -1).T
    Gi = np.tril(Gi_d.cpu().numpy(), -1)
    Gi = Gi - Gi.T
    return Gr + 1j * Gi


def _chol_upper_shifted_c(G):
    """Complex version of _chol_upper_shifted: R upper with G = R^H R."""
    try:
        return np.linalg.cholesky(G).conj().T, False
    except np.linalg.LinAlgError:
        pass
    m = G.shape[0]
    shift = 1e-13 * m
    while True:
        try:
            return np.linalg.cholesky(G + shift * np.eye(m)).conj().T, True
        except np.linalg.LinAlgError:
            shift *= 100.0
            if shift > 1.0:
                raise


def smallest_eigenpairs_paired(A, k, upper_bound, lower_bound=0.0, tol=1e-12, nex=None, seed=0, max_outer=80,
                               cond_max=1e6, deg0=20, panel=None, stats=None, verbose=False, comm=None,
                               refine_bound=True, init_fn=None, hi_override=None, m_exact=None, orthonormal_start=False):
    """Smallest k eigenpairs of a d = 2 block matrix ``A`` that commutes with J (all blocks scaled rotations), through its
    complex-Hermitian form.  Same contract as ``smallest_eigenpairs``: returns (evals (k,), evecs (N, k)) with unit-norm
    columns, ascending; columns 2j and 2j+1 are (v_j, J v_j) of the j-th complex eigenpair."""
    dev = A.indptr.device
    h = get_handle(dev.index)
    assert A.d == 2
    N = A.nrows                       # LOCAL real rows
    nn = N // 2                       # local nodes
    Nglob = N
    if comm is not None and comm.world > 1:
        t = torch.tensor([N], dtype=torch.int64, device=dev)
        comm.allreduce_(t)
        Nglob = int(t.item())
    else:
        comm = None
    nglob = Nglob // 2                # dimension of the complex problem
    k = int(min(k, Nglob))
    kc = (k + 1) // 2
    nexc = max(8, int(math.ceil(0.2 * kc))) if nex is None else max(1, (int(nex) + 1) // 2)
    mc = min(nglob, kc + nexc)
    if panel is None:
        panel = 64 if mc >= 128 else 32
    if m_exact is not None:
        mc = int(min(nglob, max(kc, m_exact)))
    elif mc < nglob:
        mc = min(nglob, ((mc + panel - 1) // panel) * panel)
    hi = float(upper_bound)
    if hi_override is not None:
        hi = float(hi_override)
    elif refine_bound and Nglob > 8 * mc:
        hi = min(hi, 1.01 * lanczos_upper_bound(A, comm=comm, h=h))
    lo_spec = float(lower_bound)
    tol_abs = tol * float(upper_bound)
    import os as _os
    cond_max = float(_os.environ.get("RVGP_EIG_COND_MAX", cond_max))     # policy knob (profiles/r01d_eig_policy_emulation.txt)
    st = stats if stats is not None else {}
    st.update(dict(N=N, k=k, m=mc, paired=True, panel=panel, spmm_launches=0, filter_launches=0, filter_col_degrees=0,
                   outer=0, t_filter=0.0, t_dense=0.0, t_host=0.0, spmm_bytes_fused=int(A.spmm_bytes(panel, fused=True)),
                   spmm_bytes_plain=int(A.spmm_bytes(panel, fused=False)), d=A.d, world=(comm.world if comm else 1),
                   hi=hi, hi_gershgorin=float(upper_bound)))

    B1 = torch.empty((N, mc), dtype=torch.float64, device=dev)
    B2 = torch.empty((N, mc), dtype=torch.float64, device=dev)
    JV = torch.empty((N, mc), dtype=torch.float64, device=dev)
    w0 = torch.empty((N, panel), dtype=torch.float64, device=dev)
    w1 = torch.empty((N, panel), dtype=torch.float64, device=dev)
    w2 = None
    if isinstance(A, BsrMatrix) and A.mma is not None:
        w2 = torch.empty((N, panel), dtype=torch.float64, device=dev)
    st["spmm_kernel"] = "mma_native" if w2 is not None else "gather"
    Gr = torch.empty((mc, mc), dtype=torch.float64, device=dev)
    Gi = torch.empty((mc, mc), dtype=torch.float64, device=dev)
    Hr = torch.empty((mc, mc), dtype=torch.float64, device=dev)
    Hi = torch.empty((mc, mc), dtype=torch.float64, device=dev)
    Cr = torch.empty((mc, mc), dtype=torch.float64, device=dev)
    Ci = torch.empty((mc, mc), dtype=torch.float64, device=dev)
    theta_d = torch.empty(mc, dtype=torch.float64, device=dev)
    dense = _Dense(h, N, mc, dev, comm)

    def rot90(X, out):
        h.call("rvgp_rot90_nodes_f64", I64(nn), int(X.shape[1]), X, I64(X.stride(0)), out, I64(out.stride(0)))
        return out

    def gram_c(X, JX, Y, outr, outi):
        """X^H Y (Hermitian expected): lower tiles of X^T Y and (J X)^T Y."""
        outr.zero_(); outi.zero_()
        dense.gram(X, Y, outr, sym=True)
        dense.gram(JX, Y, outi, sym=True)
        return _herm_from_lower(outr, outi)

    def apply_c(X, JX, Cm, out):
        """out = X Cm for a complex (mc x mc) host matrix: X Re(Cm) + (J X) Im(Cm)."""
        Cr.copy_(torch.from_numpy(np.ascontiguousarray(Cm.real)))
        Ci.copy_(torch.from_numpy(np.ascontiguousarray(Cm.imag)))
        dense.apply(X, Cr, out)
        h.call("rvgp_dgemm_acc_f64", int(N), int(mc), I64(mc), 1.0, JX, I64(JX.stride(0)), 1, Ci, I64(Ci.stride(0)), 0, 1.0,
               out, I64(out.stride(0)))
        return out

    V, W = B1, B2
    h.call("rvgp_fill_uniform_f64", I64(N), int(mc), V, I64(V.stride(0)), U64(seed), I64(0), I64(A.row_offset))
    if init_fn is not None:
        init_fn(V)

    deg = np.full(mc, deg0, dtype=np.int64)
    a_cut = lo_spec + 0.3 * (hi - lo_spec)
    theta = res = None
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev2 = torch.cuda.Event(enable_timing=True)

    for it in range(max_outer):
        ev0.record()
        _nvtx.push("eig:filter")
        for p0 in range(0, mc, panel):
            p1 = min(mc, p0 + panel)
            dg = int(deg[p0:p1].max())
            if dg <= 0:
                continue
            if w2 is not None and p1 - p0 == panel:
                A.cheb_filter(V[:, p0:p1], w0, w1, p1 - p0, dg, lo_spec, float(a_cut), hi, h=h, w2=w2)
            else:
                A.cheb_filter(V[:, p0:p1], w0, w1, p1 - p0, dg, lo_spec, float(a_cut), hi, h=h)
            st["spmm_launches"] += dg
            st["filter_launches"] += dg
            st["filter_col_degrees"] += dg * (p1 - p0)
        ev1.record()
        _nvtx.pop()
        _nvtx.push("eig:orthonormalise+rayleigh_ritz")
        # ---- complex CholeskyQR (shifted CholeskyQR3 on breakdown), then Rayleigh-Ritz ---------------------------------
        for _pass in range(0 if (orthonormal_start and it == 0) else 4):
            nrm = dense.coldot(V, V)
            inv = torch.rsqrt(nrm)
            dense.colscale(V, inv)
            rot90(V, JV)
            G = gram_c(V, JV, V, Gr, Gi)
            t0 = time.perf_counter()
            with _lapack_ctx():
                Rinv, shifted = _cholqr_factor(G, dev)
            st["t_host"] += time.perf_counter() - t0
            apply_c(V, JV, Rinv, W)
            V, W = W, V
            st["cholqr_passes"] = st.get("cholqr_passes", 0) + 1
            if not shifted:
                break
        rot90(V, JV)
        A.matmat(V, out=W, h=h)                            # W = A V
        st["spmm_launches"] += math.ceil(mc / 64)
        G = gram_c(V, JV, V, Gr, Gi)
        Hm = gram_c(V, JV, W, Hr, Hi)                      # V^H A V
        t0 = time.perf_counter()
        with _lapack_ctx():
            theta, Cm = _rayleigh_ritz_small(G, Hm, dev)
        st["t_host"] += time.perf_counter() - t0
        theta_d.copy_(torch.from_numpy(np.ascontiguousarray(theta)))
        apply_c(V, JV, Cm, W)                              # Ritz vectors
        V, W = W, V
        A.matmat(V, out=W, h=h)                            # A * Ritz vectors, for true residuals
        st["spmm_launches"] += math.ceil(mc / 64)
        res = torch.sqrt(dense.resid_sq(W, V, theta_d)).cpu().numpy()
        ev2.record()
        _nvtx.pop()
        torch.cuda.synchronize(dev)
        st["t_filter"] += ev0.elapsed_time(ev1) * 1e-3
        st["t_dense"] += ev1.elapsed_time(ev2) * 1e-3
        st["outer"] = it + 1

        nconv = int((res[:kc] <= tol_abs).sum())
        a_cut = float(theta[-1]) if mc < nglob else a_cut
        if verbose:
            print("  [eig paired] it %d  cut=%.6g  conv=%d/%d  maxres=%.3e  theta_k=%.9g" %
                  (it, a_cut, nconv, kc, res[:kc].max(), theta[kc - 1]))
        if nconv == kc or mc >= nglob:
            break
        a_cut = max(a_cut, lo_spec + 1e-12 * (hi - lo_spec) + theta[kc - 1] * (1 + 1e-9))
        deg = _next_degrees(theta, res, kc, tol_abs, a_cut, hi, lo_spec, cond_max)

    if not isinstance(A, BsrMatrix) and hasattr(A, "spmm_kernel_name"):
        st["spmm_kernel"] = A.spmm_kernel_name
    st["residual_max"] = float(res[:kc].max())
    st["converged"] = bool((res[:kc] <= tol_abs).all())
    st["tol_abs"] = tol_abs
    _warn_unconverged(st, max_outer)
    # expand every complex pair (theta_j, v_j) to the two real eigenpairs (theta_j, v_j), (theta_j, J v_j)
    del W, B1, B2                                        # V keeps the buffer it points to
    Vk = V[:, :kc]
    rot90(Vk, JV[:, :kc])
    evecs = torch.empty((N, 2 * kc), dtype=torch.float64, device=dev)
    evecs[:, 0::2] = Vk
    evecs[:, 1::2] = JV[:, :kc]
    evals = theta_d[:kc].repeat_interleave(2)[:k].clone()
    if 2 * kc != k:
        evecs = evecs[:, :k].contiguous()
    return evals, evecs
