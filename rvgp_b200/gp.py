"""Device-side GP regression algebra for the spectral (rank <= k) vector-field kernel.

Replaces what GPflow's GPR does through TensorFlow for the reference (RVGP/main.py:55-58,77,80,111):
``log_marginal_likelihood`` (Cholesky of K + sigma^2 I, triangular solve, log-det), its gradient (reverse-mode AD
there, analytic here) and ``predict_f`` (base_conditional).  Two solvers with identical mathematics:

  dense    what GPflow does: M x M Gram (K13 dgemm with the spectral density folded in), blocked Cholesky and
           triangular solves (K14).  FP64-FLOP-bound, O(M^3); for M up to ~60k on one B200.
  lowrank  K = Phi S Phi^T has rank <= k, so every quantity follows from G = Phi^T Phi (k x k), b = Phi^T y,
           y^T y (K15: one tall-skinny Gram pass over the M rows) and k x k Cholesky/solves per evaluation.
           This is the only form that exists at config C4 (M = 1.2 M -> an 11.5 TB Gram matrix).

All O(M) and O(k^3) arithmetic runs in librvgp_b200.so kernels; only k-vector glue and the 4-scalar chain rule
stay on the host next to SciPy's L-BFGS-B (as SURVEY.md section 1 prescribes).
"""
import math

import numpy as np
import torch

from ._cabi import get_handle, I64, RvgpError, RVGP_ERR_NOT_SPD
from .eigensolver import _dgemm

LOG2PI = math.log(2.0 * math.pi)


def _f64(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


class _Chol:
    """In-place blocked Cholesky of an n x n device matrix + solves against it."""

    def __init__(self, h, A, n):
        self.h, self.A, self.n = h, A, n
        wsb = h.query("rvgp_potrf_workspace_bytes", int(n))
        self.ws = torch.empty(max(8, wsb), dtype=torch.uint8, device=A.device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=A.device)
        h.call("rvgp_potrf_f64", A, I64(A.stride(0)), int(n), self.flag, self.ws, I64(wsb))
        self._scratch = None

    def check(self):
        if int(self.flag.item()) & 1:
            raise RvgpError(RVGP_ERR_NOT_SPD, "Cholesky decomposition was not successful. The input might not be valid.")

    def solve(self, B, trans):
        """L X = B (trans 0) / L^T X = B (trans 1), in place on B (n x nrhs, contiguous rows)."""
        nrhs = B.shape[1]
        if self._scratch is None or self._scratch.numel() < 64 * nrhs:
            self._scratch = torch.empty(64 * nrhs, dtype=torch.float64, device=B.device)
        self.h.call("rvgp_trsm_f64", self.A, I64(self.A.stride(0)), int(self.n), B, I64(B.stride(0)), int(nrhs),
                    int(trans), self.ws, self._scratch)
        return B

    def logdiag_sum(self):
        out = torch.empty(1, dtype=torch.float64, device=self.A.device)
        self.h.call("rvgp_logdiag_sum_f64", self.A, I64(self.A.stride(0)), int(self.n), out)
        return out


class DeviceGPR:
    """GP regression with kernel K = Phi diag(S) Phi^T + noise I on rows X (M, k), single output column Y (M, 1)."""

    def __init__(self, X, Y, solver="auto", comm=None):
        """comm (distributed.Comm, world > 1): X / Y are THIS RANK's training rows; the rank-k path needs exactly one
        all-reduce of [Phi y]^T [Phi y] ((k+1)^2 doubles), after which every rank holds the same G, b, y^T y and runs the same
        host optimisation.  The dense M x M path cannot be row-sharded: callers pass comm only with lowrank."""
        self.X, self.Y = X.contiguous(), Y.contiguous().reshape(-1, 1)
        self.M_local, self.k = self.X.shape
        self.M = self.M_local
        self.dev = X.device
        self.h = get_handle(self.dev.index)
        self.comm = comm if (comm is not None and comm.world > 1) else None
        if self.comm is not None:
            t = torch.tensor([self.M_local], dtype=torch.int64, device=self.dev)
            self.comm.allreduce_(t)
            self.M = int(t.item())
        if solver == "auto":
            solver = "lowrank" if self.M >= 2 * self.k else "dense"
        if self.comm is not None and solver != "lowrank":
            raise ValueError("row-sharded GP needs the rank-k solver (global M >= 2 k)")
        self.solver = solver
        self.n_eval = 0
        if solver == "lowrank":
            self._init_lowrank()

    # ---- K15: tall-skinny Gram pass --------------------------------------------------------------------
    def _init_lowrank(self):
        h, M, k = self.h, self.M, self.k
        Ml = self.M_local
        XY = torch.cat([self.X, self.Y], dim=1).contiguous()          # (M_local, k+1): one pass gives G, b and y^T y
        kk = k + 1
        split = max(1, min(64, Ml // 2048))
        ws = torch.empty(split * kk * kk, dtype=torch.float64, device=self.dev)
        Gd = torch.zeros((kk, kk), dtype=torch.float64, device=self.dev)
        if Ml > 0:
            _dgemm(h, kk, kk, Ml, XY, XY.stride(0), 0, XY, XY.stride(0), 0, Gd, Gd.stride(0), split_k=split, ws=ws)
        if self.comm is not None:
            self.comm.allreduce_(Gd)                                  # the ONE collective of the row-sharded fit
        Gh = Gd.cpu().numpy()
        Gh = 0.5 * (Gh + Gh.T)
        self.G, self.b, self.yy = Gh[:k, :k].copy(), Gh[:k, k].copy(), float(Gh[k, k])
        self._Gdiag = np.diag(self.G).copy()
        self._small = k <= 64
        if not self._small:
            # K15b: the whole evaluation is one C call on resident data (rvgp_gp_lowrank_eval_f64); per evaluation the host
            # uploads k + 1 doubles and reads back 2 + 2k
            self._Gd = _f64(self.G, self.dev)
            self._bd = _f64(self.b, self.dev)
            self._par_h = torch.empty(k + 1, dtype=torch.float64).pin_memory()
            self._par_d = torch.empty(k + 1, dtype=torch.float64, device=self.dev)
            self._out_h = torch.empty(2 + 2 * k, dtype=torch.float64).pin_memory()
            self._out_d = torch.empty(2 + 2 * k, dtype=torch.float64, device=self.dev)
            self._ev_wsb = int(self.h.query("rvgp_gp_lowrank_eval_workspace_bytes", int(k)))
            self._ev_ws = torch.empty(self._ev_wsb, dtype=torch.uint8, device=self.dev)
        if self._small:
            # K16: everything an evaluation needs stays resident; one fused launch per evaluation
            self._Gd = _f64(self.G, self.dev)
            self._bd = _f64(self.b, self.dev)
            self._scal = _f64(np.array([self.yy, float(M)]), self.dev)
            self._par_h = torch.empty(k + 1, dtype=torch.float64).pin_memory()
            self._par_d = torch.empty(k + 1, dtype=torch.float64, device=self.dev)
            self._out_d = torch.empty(2 + 2 * k + k * k, dtype=torch.float64, device=self.dev)

    def _small_eval(self, S, noise, want_predict):
        k = self.k
        self._par_h[:k] = torch.from_numpy(np.ascontiguousarray(S))
        self._par_h[k] = noise
        self._par_d.copy_(self._par_h, non_blocking=True)
        self.h.sync_stream()
        self.h.call("rvgp_gp_lowrank_small_f64", 1, int(k), self._Gd, I64(0), self._bd, I64(0), self._scal[0:1],
                    self._scal[1:2], self._par_d, I64(0), self._par_d[k:k + 1], self._out_d, I64(self._out_d.numel()),
                    int(want_predict))
        return self._out_d.cpu().numpy()

    # ---- log marginal likelihood and gradient ------------------------------------------------------------
    def lml_and_grads(self, S, noise, grads=True):
        self.n_eval += 1
        if self.solver == "lowrank":
            return self._lml_lowrank(np.asarray(S, dtype=np.float64), float(noise), grads)
        return self._lml_dense(np.asarray(S, dtype=np.float64), float(noise), grads)

    def _lowrank_factor(self, S, noise):
        k = self.k
        rs = np.sqrt(S)
        B = np.eye(k) + (rs[:, None] * self.G * rs[None, :]) / noise
        Bd = _f64(B, self.dev)
        ch = _Chol(self.h, Bd, k)
        return rs, ch

    def _lml_lowrank(self, S, noise, grads):
        k, M = self.k, self.M
        if getattr(self, "_small", False):
            o = self._small_eval(S, noise, False)
            if not np.isfinite(o[0]):
                raise RvgpError(RVGP_ERR_NOT_SPD, "Cholesky decomposition was not successful. The input might not be valid.")
            return (float(o[0]), o[2:2 + k].copy(), float(o[1])) if grads else float(o[0])
        rs = np.sqrt(S)
        bt = rs * self.b
        self._par_h[:k] = torch.from_numpy(np.ascontiguousarray(S))
        self._par_h[k] = noise
        self._par_d.copy_(self._par_h, non_blocking=True)
        self.h.sync_stream()
        self.h.call("rvgp_gp_lowrank_eval_f64", int(k), self._Gd, self._bd, self._par_d, self._out_d, self._ev_ws, I64(self._ev_wsb))
        self._out_h.copy_(self._out_d)                          # the one synchronising read-back of the evaluation
        o = self._out_h.numpy()
        if o[0] != 0.0 or not np.isfinite(o[1]):
            raise RvgpError(RVGP_ERR_NOT_SPD, "Cholesky decomposition was not successful. The input might not be valid.")
        logdet_half = float(o[1])
        z = o[2:2 + k].copy()
        qs_h = o[2 + k:2 + 2 * k].copy()
        quad = (self.yy - (bt @ z) / noise) / noise
        lml = float(-0.5 * quad - 0.5 * M * LOG2PI - 0.5 * M * math.log(noise) - logdet_half)
        if not grads:
            return lml
        c = rs * z
        Gc = self.G @ c
        u = (self.b - Gc / noise) / noise
        wdiag = (self._Gdiag - qs_h / noise) / noise
        dS = 0.5 * u ** 2 - 0.5 * wdiag
        aa = (self.yy - 2.0 * (self.b @ c) / noise + (c @ Gc) / noise ** 2) / noise ** 2
        tr_inv = (M - (S * wdiag).sum()) / noise
        return lml, dS, 0.5 * aa - 0.5 * tr_inv

    def _dense_factor(self, S, noise):
        h, M, k = self.h, self.M, self.k
        Sd = _f64(S, self.dev)
        Ky = torch.empty((M, M), dtype=torch.float64, device=self.dev)
        # K13: Gram (X * S) X^T
        _dgemm(h, M, M, k, self.X, self.X.stride(0), 1, self.X, self.X.stride(0), 1, Ky, Ky.stride(0), scale_k=Sd)
        h.call("rvgp_add_diag_f64", Ky, I64(Ky.stride(0)), int(M), float(noise))
        return _Chol(h, Ky, M)      # K14

    def _lml_dense(self, S, noise, grads):
        h, M, k = self.h, self.M, self.k
        ch = self._dense_factor(S, noise)
        alpha = self.Y.clone()
        ch.solve(alpha, 0)
        ws = torch.empty(max(1, h.query("rvgp_coldot_workspace_bytes", I64(M), int(max(k, 1))) // 8),
                         dtype=torch.float64, device=self.dev)
        s1 = torch.empty(1, dtype=torch.float64, device=self.dev)
        h.call("rvgp_coldot_f64", I64(M), 1, alpha, I64(1), alpha, I64(1), s1, ws)
        logdet_half = ch.logdiag_sum()
        ch.check()
        lml = float(-0.5 * float(s1.item()) - 0.5 * M * LOG2PI - float(logdet_half.item()))
        if not grads:
            return lml
        a = alpha.clone()
        ch.solve(a, 1)                                   # Ky^-1 y
        Z = self.X.clone()
        ch.solve(Z, 0)                                   # L^-1 Phi
        wd = torch.empty(k, dtype=torch.float64, device=self.dev)
        h.call("rvgp_coldot_f64", I64(M), int(k), Z, I64(Z.stride(0)), Z, I64(Z.stride(0)), wd, ws)
        ud = torch.empty((k, 1), dtype=torch.float64, device=self.dev)
        split = max(1, min(32, M // 4096))
        gws = torch.empty(split * k, dtype=torch.float64, device=self.dev)
        _dgemm(h, k, 1, M, self.X, self.X.stride(0), 0, a, 1, 0, ud, 1, split_k=split, ws=gws)     # Phi^T a
        h.call("rvgp_coldot_f64", I64(M), 1, a, I64(1), a, I64(1), s1, ws)
        wdiag = wd.cpu().numpy()
        u = ud.cpu().numpy()[:, 0]
        dS = 0.5 * u ** 2 - 0.5 * wdiag
        tr_inv = (M - (S * wdiag).sum()) / noise
        return lml, dS, 0.5 * float(s1.item()) - 0.5 * tr_inv

    # ---- prediction (GPR.predict_f, full_cov=False) ------------------------------------------------------
    def predict(self, S, noise, Xnew, chunk=262144):
        S = np.asarray(S, dtype=np.float64)
        noise = float(noise)
        Xnew = Xnew.contiguous()
        Ns = Xnew.shape[0]
        mean = torch.empty((Ns, 1), dtype=torch.float64, device=self.dev)
        var = torch.empty((Ns, 1), dtype=torch.float64, device=self.dev)
        h, k = self.h, self.k
        if self.solver == "lowrank":
            if getattr(self, "_small", False):
                o = self._small_eval(S, noise, True)
                if not np.isfinite(o[0]):
                    raise RvgpError(RVGP_ERR_NOT_SPD, "Cholesky decomposition was not successful.")
                wbar = self._out_d[2 + k:2 + 2 * k].reshape(k, 1).clone()
                Q = self._out_d[2 + 2 * k:2 + 2 * k + k * k].reshape(k, k).clone()
            else:
                rs, ch = self._lowrank_factor(S, noise)
                rhs = _f64((rs * self.b).reshape(k, 1), self.dev)
                ch.solve(rhs, 0); ch.solve(rhs, 1)
                wbar = _f64((rs * rhs.cpu().numpy()[:, 0] / noise).reshape(k, 1), self.dev)
                Q = _f64(np.diag(rs), self.dev)
                ch.solve(Q, 0)                               # L_b^-1 S^1/2
                ch.check()
            ones = torch.ones(k, dtype=torch.float64, device=self.dev)
            for r0 in range(0, Ns, chunk):
                r1 = min(Ns, r0 + chunk)
                Xc = Xnew[r0:r1]
                T = torch.empty((r1 - r0, k), dtype=torch.float64, device=self.dev)
                _dgemm(h, r1 - r0, k, k, Xc, Xc.stride(0), 1, Q, Q.stride(0), 1, T, T.stride(0))
                h.call("rvgp_kdiag_f64", T, I64(T.stride(0)), I64(r1 - r0), int(k), ones, var[r0:r1])
                _dgemm(h, r1 - r0, 1, k, Xc, Xc.stride(0), 1, wbar, 1, 0, mean[r0:r1], 1)
            return mean, var
        # dense: base_conditional.  A = L^-1 Kmn ; fvar = Knn - colsum(A^2) ; fmean = A^T (L^-1 y)
        M = self.M
        ch = self._dense_factor(S, noise)
        alpha = self.Y.clone()
        ch.solve(alpha, 0)
        ch.check()
        Sd = _f64(S, self.dev)
        cchunk = max(1, min(chunk, (1 << 31) // max(M, 1)))
        ws = torch.empty(max(1, h.query("rvgp_coldot_workspace_bytes", I64(M), int(min(cchunk, Ns))) // 8),
                         dtype=torch.float64, device=self.dev)
        for r0 in range(0, Ns, cchunk):
            r1 = min(Ns, r0 + cchunk)
            nc = r1 - r0
            Xc = Xnew[r0:r1]
            Kmn = torch.empty((M, nc), dtype=torch.float64, device=self.dev)
            _dgemm(h, M, nc, k, self.X, self.X.stride(0), 1, Xc, Xc.stride(0), 1, Kmn, Kmn.stride(0), scale_k=Sd)
            ch.solve(Kmn, 0)
            knn = torch.empty(nc, dtype=torch.float64, device=self.dev)
            h.call("rvgp_kdiag_f64", Xc, I64(Xc.stride(0)), I64(nc), int(k), Sd, knn)
            a2 = torch.empty(nc, dtype=torch.float64, device=self.dev)
            h.call("rvgp_coldot_f64", I64(M), int(nc), Kmn, I64(Kmn.stride(0)), Kmn, I64(Kmn.stride(0)), a2, ws)
            h.call("rvgp_axpy_f64", I64(nc), 1, -1.0, a2, I64(1), knn, I64(1))
            var[r0:r1, 0] = knn
            split = max(1, min(32, M // 4096))
            gws = torch.empty(split * nc, dtype=torch.float64, device=self.dev)
            _dgemm(h, nc, 1, M, Kmn, Kmn.stride(0), 0, alpha, 1, 0, mean[r0:r1], 1, split_k=split, ws=gws)
        return mean, var
