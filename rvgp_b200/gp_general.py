"""Device-side GP algebra for ANY kernel object and R output columns: dense GPR and the SGPR collapsed bound.

SURVEY.md 8f rows 1 and 3 (the callers either side of the hot path):
  * ``DenseGPR``   -- gpflow.models.GPR as reached from ``train_gp(kernel='rbf')`` (RVGP/main.py:33-37,55-58): the
    channel-wise baseline, one shared kernel, D output columns.
  * ``DeviceSGPR`` -- gpflow.models.SGPR (Titsias' collapsed bound) as reached from ``train_gp(n_inducing_points=...)``
    (RVGP/main.py:59-67,119-137), inducing points trainable like GPflow's default.

GPflow differentiates these with TensorFlow's reverse mode; here the adjoints are written out by hand
and every O(M^2), O(N Mu) and O(Mu^3) operation is a librvgp_b200.so kernel (K10 dgemm, K14 potrf/trsm, K13/K17 Gram +
adjoint, column reductions).  The kernel object supplies ``_gram``, ``_diag``, ``_adjoint`` and ``_diag_adjoint_uniform``
(rvgp_b200/kernels.py: ManifoldKernel, RBF).  Host side: scalars and the k-vector chain rule only.
"""
import math

import torch

from ._cabi import get_handle, I64
from .eigensolver import _dgemm
from .gp import _Chol, LOG2PI

DEFAULT_JITTER = 1e-6          # gpflow.config.default_jitter()


def _dgemm_acc(h, m, n, k, alpha, A, lda, a_kmajor, B, ldb, b_kmajor, beta, C, ldc):
    h.call("rvgp_dgemm_acc_f64", int(m), int(n), I64(k), float(alpha), A, I64(lda), int(a_kmajor), B, I64(ldb),
           int(b_kmajor), float(beta), C, I64(ldc))


def _scale_shift(h, A, alpha, beta):
    h.call("rvgp_scale_shift_f64", I64(A.shape[0]), int(A.shape[1]), float(alpha), float(beta), A, I64(A.stride(0)))


def _colsq_sum(h, A):
    """sum of squares of a (rows x cols) device matrix: column reduction kernel + a cols-long host-side sum."""
    rows, cols = A.shape
    out = torch.empty(cols, dtype=torch.float64, device=A.device)
    ws = torch.empty(max(1, h.query("rvgp_coldot_workspace_bytes", I64(rows), int(cols)) // 8), dtype=torch.float64,
                     device=A.device)
    h.call("rvgp_coldot_f64", I64(rows), int(cols), A, I64(A.stride(0)), A, I64(A.stride(0)), out, ws)
    return out


def _split_for(K):
    """Deterministic split-K factor for products with a long contraction (K) and a small output."""
    return max(1, min(64, K // 2048)) if K >= 4096 else 1


def _inverse_from_chol(ch, n, dev):
    """(L L^T)^-1 as a full matrix: two triangular solves against the identity."""
    Inv = torch.eye(n, dtype=torch.float64, device=dev)
    ch.solve(Inv, 0)
    ch.solve(Inv, 1)
    return Inv


class DenseGPR:
    """GP regression, kernel object ``kernel``, rows X (M, k), outputs Y (M, R) sharing the kernel."""

    def __init__(self, X, Y, kernel):
        self.X, self.Y = X.contiguous(), Y.contiguous()
        if self.Y.dim() == 1:
            self.Y = self.Y.reshape(-1, 1)
        self.M, self.R = self.Y.shape
        self.kernel = kernel
        self.dev = X.device
        self.h = get_handle(self.dev.index)
        self.n_eval = 0

    def _factor(self, noise):
        Ky = self.kernel._gram(self.X, self.X)
        self.h.call("rvgp_add_diag_f64", Ky, I64(Ky.stride(0)), int(self.M), float(noise))
        ch = _Chol(self.h, Ky, self.M)
        alpha = self.Y.clone()
        ch.solve(alpha, 0)                                   # L^-1 Y
        return ch, alpha

    def lml_and_grads(self, noise, grads=True):
        """GPR.log_marginal_likelihood (gpflow/models/gpr.py) and, with grads, ({kernel parameter: d/dvalue}, d/dnoise)."""
        self.n_eval += 1
        h, M, R = self.h, self.M, self.R
        noise = float(noise)
        ch, alpha = self._factor(noise)
        quad = float(_colsq_sum(h, alpha).sum().item())
        logdet_half = float(ch.logdiag_sum().item())
        ch.check()
        lml = -0.5 * quad - 0.5 * M * R * LOG2PI - R * logdet_half
        if not grads:
            return lml
        a = alpha.clone()
        ch.solve(a, 1)                                       # Ky^-1 Y   (M x R)
        W = _inverse_from_chol(ch, M, self.dev)              # Ky^-1
        # W <- 1/2 (a a^T - R Ky^-1): the adjoint of Ky
        _dgemm_acc(h, M, M, R, 0.5, a, a.stride(0), 1, a, a.stride(0), 1, -0.5 * R, W, W.stride(0))
        dnoise = float(torch.diagonal(W).sum().item())
        kgrads, _ = self.kernel._adjoint(self.X, self.X, W, want_dX=False)
        return lml, kgrads, dnoise

    def predict(self, noise, Xnew, chunk=65536):
        """GPR.predict_f(full_cov=False) = base_conditional: mean (N*, R), variance (N*, R) (one column tiled)."""
        h, M, R = self.h, self.M, self.R
        Xnew = Xnew.contiguous()
        Ns = Xnew.shape[0]
        ch, alpha = self._factor(float(noise))
        ch.check()
        mean = torch.empty((Ns, R), dtype=torch.float64, device=self.dev)
        var = torch.empty(Ns, dtype=torch.float64, device=self.dev)
        cchunk = max(1, min(chunk, (1 << 30) // max(M, 1)))
        for r0 in range(0, Ns, cchunk):
            r1 = min(Ns, r0 + cchunk)
            nc = r1 - r0
            Xc = Xnew[r0:r1]
            Kmn = self.kernel._gram(self.X, Xc)              # (M, nc)
            ch.solve(Kmn, 0)                                 # A = L^-1 Kmn
            var[r0:r1] = self.kernel._diag(Xc) - _colsq_sum(h, Kmn)
            split = _split_for(M)
            ws = torch.empty(split * nc * R, dtype=torch.float64, device=self.dev) if split > 1 else None
            _dgemm(h, nc, R, M, Kmn, Kmn.stride(0), 0, alpha, alpha.stride(0), 0, mean[r0:r1], R, split_k=split, ws=ws)
        return mean, var.reshape(-1, 1).expand(Ns, R).contiguous()


class DeviceSGPR:
    """Sparse GP regression (collapsed bound): data X (N, k), Y (N, R), inducing points Z (Mu, k)."""

    def __init__(self, X, Y, Z, kernel, jitter=DEFAULT_JITTER):
        self.X, self.Y = X.contiguous(), Y.contiguous()
        if self.Y.dim() == 1:
            self.Y = self.Y.reshape(-1, 1)
        self.Z = Z.contiguous().clone()
        self.N, self.R = self.Y.shape
        self.Mu = self.Z.shape[0]
        self.kernel, self.jitter = kernel, float(jitter)
        self.dev = X.device
        self.h = get_handle(self.dev.index)
        self.yy = float(_colsq_sum(self.h, self.Y).sum().item())
        self._cache = {}
        self.n_eval = 0

    def _common(self, noise, Z):
        """L = chol(Kuu + jitter I), Ap = L^-1 Kuf, B = I + Ap Ap^T / noise, LB = chol(B), c = LB^-1 Ap Y / noise
        (sgpr.py: A = Ap / sigma, c as there)."""
        h, N, R, Mu = self.h, self.N, self.R, self.Mu
        Kuu = self.kernel._gram(Z, Z)
        h.call("rvgp_add_diag_f64", Kuu, I64(Kuu.stride(0)), int(Mu), self.jitter)
        chL = _Chol(h, Kuu, Mu)
        Ap = self.kernel._gram(Z, self.X)                    # (Mu, N)
        chL.solve(Ap, 0)
        B = torch.empty((Mu, Mu), dtype=torch.float64, device=self.dev)
        split = _split_for(N)
        ws = torch.empty(split * Mu * Mu, dtype=torch.float64, device=self.dev) if split > 1 else None
        _dgemm(h, Mu, Mu, N, Ap, Ap.stride(0), 1, Ap, Ap.stride(0), 1, B, B.stride(0), alpha=1.0 / noise, split_k=split, ws=ws)
        h.call("rvgp_add_diag_f64", B, I64(B.stride(0)), int(Mu), 1.0)
        Bfull = B.clone()
        trB = float(torch.diagonal(B).sum().item())
        chB = _Chol(h, B, Mu)
        c = torch.empty((Mu, R), dtype=torch.float64, device=self.dev)
        ws2 = torch.empty(split * Mu * R, dtype=torch.float64, device=self.dev) if split > 1 else None
        _dgemm(h, Mu, R, N, Ap, Ap.stride(0), 1, self.Y, self.Y.stride(0), 0, c, c.stride(0), alpha=1.0 / noise,
               split_k=split, ws=ws2)
        chB.solve(c, 0)
        return chL, Ap, Bfull, trB, chB, c

    def elbo_and_grads(self, noise, grads=True, Z=None):
        """SGPR.elbo (gpflow/models/sgpr.py) and, with grads, ({kernel parameter: d/dvalue}, d/dnoise, d/dZ (Mu, k) cuda)."""
        self.n_eval += 1
        h, N, R, Mu = self.h, self.N, self.R, self.Mu
        noise = float(noise)
        Z = self.Z if Z is None else Z.contiguous()
        chL, Ap, Bfull, trB, chB, c = self._common(noise, Z)
        sum_c2 = float(_colsq_sum(h, c).sum().item())
        logdetLB = float(chB.logdiag_sum().item())
        chL.check()
        chB.check()
        kdiag_sum = float(self.kernel._diag(self.X).sum().item())
        trAAT = trB - Mu
        bound = (-0.5 * N * R * LOG2PI - R * logdetLB - 0.5 * N * R * math.log(noise) - 0.5 * self.yy / noise
                 + 0.5 * sum_c2 - 0.5 * R * kdiag_sum / noise + 0.5 * R * trAAT)
        if not grads:
            return bound
        # ---- reverse mode ---------------------------------------------------------------------
        t = c.clone()
        chB.solve(t, 1)                                                      # t = LB^-T c
        Binv = _inverse_from_chol(chB, Mu, self.dev)
        trBinv = float(torch.diagonal(Binv).sum().item())
        beta = self.Y.clone()                                                # beta = (Y - Ap^T t) / noise
        _dgemm_acc(h, N, R, Mu, -1.0 / noise, Ap, Ap.stride(0), 0, t, t.stride(0), 0, 1.0 / noise, beta, beta.stride(0))
        sum_b2 = float(_colsq_sum(h, beta).sum().item())
        E = Binv.clone()
        _scale_shift(h, E, -R / noise, R / noise)                            # (R / noise) (I - B^-1)
        Q = torch.empty((Mu, N), dtype=torch.float64, device=self.dev)
        _dgemm(h, Mu, N, Mu, E, E.stride(0), 1, Ap, Ap.stride(0), 0, Q, Q.stride(0))
        _dgemm_acc(h, Mu, N, R, 1.0, t, t.stride(0), 1, beta, beta.stride(0), 1, 1.0, Q, Q.stride(0))    # + t beta^T
        chL.solve(Q, 1)                                                      # dF/dKuf = L^-T [...]
        Mid = Binv                                                           # -1/2 t t^T - R/2 (B - 2 I + B^-1)
        h.call("rvgp_axpy_f64", I64(Mu), int(Mu), 1.0, Bfull, I64(Bfull.stride(0)), Mid, I64(Mid.stride(0)))
        _scale_shift(h, Mid, -0.5 * R, float(R))
        _dgemm_acc(h, Mu, Mu, R, -0.5, t, t.stride(0), 1, t, t.stride(0), 1, 1.0, Mid, Mid.stride(0))
        chL.solve(Mid, 1)                                                    # L^-T Mid
        Guu = Mid.t().contiguous()                                           # (L^-T Mid)^T = Mid L^-1
        chL.solve(Guu, 1)                                                    # dF/dKuu = L^-T Mid L^-1 (symmetric)
        g1, dZ1 = self.kernel._adjoint(Z, self.X, Q, want_dX=True)
        g2, dZ2 = self.kernel._adjoint(Z, Z, Guu, want_dX=True)
        g3 = self.kernel._diag_adjoint_uniform(self.X, -0.5 * R / noise, self._cache)
        kgrads = {n: g1[n] + g2[n] + g3[n] for n in g1}
        dnoise = (0.5 * (sum_b2 - R * (N - (Mu - trBinv)) / noise) + 0.5 * R * kdiag_sum / noise ** 2
                  - 0.5 * R * trAAT / noise)
        h.call("rvgp_axpy_f64", I64(Mu), int(Z.shape[1]), 2.0, dZ2, I64(dZ2.stride(0)), dZ1, I64(dZ1.stride(0)))
        return bound, kgrads, dnoise, dZ1

    def predict(self, noise, Xnew, chunk=65536):
        """SGPR.predict_f(full_cov=False): mean (N*, R), variance (N*, R) (one column tiled)."""
        h, R, Mu = self.h, self.R, self.Mu
        noise = float(noise)
        Xnew = Xnew.contiguous()
        Ns = Xnew.shape[0]
        chL, Ap, Bfull, trB, chB, c = self._common(noise, self.Z)
        chL.check()
        chB.check()
        del Ap, Bfull
        mean = torch.empty((Ns, R), dtype=torch.float64, device=self.dev)
        var = torch.empty(Ns, dtype=torch.float64, device=self.dev)
        cchunk = max(1, min(chunk, (1 << 30) // max(Mu, 1)))
        for r0 in range(0, Ns, cchunk):
            r1 = min(Ns, r0 + cchunk)
            nc = r1 - r0
            Xc = Xnew[r0:r1]
            T = self.kernel._gram(self.Z, Xc)                # Kus (Mu, nc)
            chL.solve(T, 0)                                  # tmp1
            s1 = _colsq_sum(h, T).clone()
            chB.solve(T, 0)                                  # tmp2
            s2 = _colsq_sum(h, T)
            var[r0:r1] = self.kernel._diag(Xc) + s2 - s1
            _dgemm(h, nc, R, Mu, T, T.stride(0), 0, c, c.stride(0), 0, mean[r0:r1], R)
        return mean, var.reshape(-1, 1).expand(Ns, R).contiguous()
