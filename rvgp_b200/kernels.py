"""Mirror of the reference's ``RVGP/kernels.py``: the spectral vector-field kernel class.

``ManifoldKernel`` keeps the constructor signature, parameter names (nu, kappa, sigma_f), ``eval_S``, ``K`` and
``K_diag`` of the reference (kernels.py:25-67) without GPflow/TensorFlow.  ``K`` / ``K_diag`` accept NumPy arrays,
torch tensors or any ``__dlpack__`` producer (host or CUDA) and return CUDA torch tensors (DLPack-exportable);
the Gram is the FP64 dgemm kernel with the spectral density folded into the operand load (K13) and ``K_diag`` is a
row-wise reduction that never forms the N x N matrix the reference builds (kernels.py:67).
"""
import numpy as np
import torch

from . import params as P
from ._cabi import get_handle, I64
from .eigensolver import _dgemm
from .geometry import to_device_f64


class ManifoldKernel:
    """Matern / squared-exponential kernel on the tangent bundle in the connection-Laplacian eigenbasis."""

    def __init__(self, data, nu=3, kappa=4, sigma_f=1, typ='matern', dtype=None):
        # read at construction time like kernels.py:27 (examples overwrite d.evals_Lc / d.evecs_Lc first)
        self.eigenvalues = np.asarray(data.evals_Lc, dtype=np.float64)
        ev = data._duals["evecs_Lc"] if hasattr(data, "_duals") else None
        self._evecs_dual = ev
        self.eigenvectors = None if ev is not None else np.asarray(data.evecs_Lc)
        nrows = (ev.dev.shape[0] if (ev is not None and ev.dev is not None) else
                 (ev.host.shape[0] if ev is not None else self.eigenvectors.shape[0]))
        self.num_verticies = float(nrows)                      # kernels.py:28 (n*D, the AMBIENT row count)
        self.dtype = dtype
        self.typ = typ
        if typ not in ('se', 'matern'):
            NotImplemented
        if typ == 'matern':
            self.nu = P.Parameter(nu, transform=P.positive(), name='nu')
        self.kappa = P.Parameter(kappa, transform=P.positive(), name='kappa')
        self.sigma_f = P.Parameter(sigma_f, transform=P.positive(), name='sigma_f')

    @property
    def trainable_parameters(self):
        ps = ([self.nu] if self.typ == 'matern' else []) + [self.kappa, self.sigma_f]
        return [p for p in ps if p.trainable]

    def eval_S(self, typ='matern', grads=False):
        """Spectral density (kernels.py:41-53); host k-vector.  grads=True adds dS/d(nu, kappa, sigma_f)."""
        lam = self.eigenvalues
        kappa, sigma_f = self.kappa.value, self.sigma_f.value
        if typ == 'matern':
            nu = self.nu.value
            a = 2.0 * nu / kappa ** 2
            S0 = np.power(lam + a, -nu)
            dl_nu = -np.log(lam + a) - nu * (2.0 / kappa ** 2) / (lam + a)
            dl_kappa = 4.0 * nu ** 2 / (kappa ** 3 * (lam + a))
        elif typ == 'se':
            S0 = np.exp(-0.5 * lam * kappa ** 2)
            dl_nu = np.zeros_like(lam)
            dl_kappa = -lam * kappa
        else:
            raise NotImplementedError(typ)
        Z = S0.sum()
        S = S0 * (self.num_verticies / Z) * sigma_f
        if not grads:
            return S
        p = S0 / Z
        return S, {"nu": S * (dl_nu - (p * dl_nu).sum()), "kappa": S * (dl_kappa - (p * dl_kappa).sum()),
                   "sigma_f": S / sigma_f}

    def K(self, X, X2=None):
        """Kernel function (kernels.py:55-61): (X * S) @ X2^T, returned as a CUDA tensor."""
        Xd = to_device_f64(X)
        X2d = Xd if X2 is None else to_device_f64(X2, Xd.device)
        h = get_handle(Xd.device.index)
        S = torch.from_numpy(self.eval_S(typ=self.typ)).to(Xd.device)
        out = torch.empty((Xd.shape[0], X2d.shape[0]), dtype=torch.float64, device=Xd.device)
        _dgemm(h, Xd.shape[0], X2d.shape[0], Xd.shape[1], Xd, Xd.stride(0), 1, X2d, X2d.stride(0), 1, out,
               out.stride(0), scale_k=S)
        return out

    def K_diag(self, X):
        """Diagonal of K (kernels.py:63-67) as a row-wise reduction."""
        Xd = to_device_f64(X)
        h = get_handle(Xd.device.index)
        S = torch.from_numpy(self.eval_S(typ=self.typ)).to(Xd.device)
        out = torch.empty(Xd.shape[0], dtype=torch.float64, device=Xd.device)
        h.call("rvgp_kdiag_f64", Xd, I64(Xd.stride(0)), I64(Xd.shape[0]), int(Xd.shape[1]), S, out)
        return out

    def __call__(self, X, X2=None, full_cov=True):
        return self.K(X, X2) if full_cov else self.K_diag(X)

    # ---- generic kernel interface used by gp_general.DenseGPR / DeviceSGPR (device tensors in and out) ----------
    def _gram(self, XA, XB, out=None):
        h = get_handle(XA.device.index)
        S = torch.from_numpy(self.eval_S(typ=self.typ)).to(XA.device)
        if out is None:
            out = torch.empty((XA.shape[0], XB.shape[0]), dtype=torch.float64, device=XA.device)
        _dgemm(h, XA.shape[0], XB.shape[0], XA.shape[1], XA, XA.stride(0), 1, XB, XB.stride(0), 1, out, out.stride(0),
               scale_k=S)
        return out

    def _diag(self, X):
        return self.K_diag(X)

    def _chain(self, dS):
        """dF/dS (host k-vector) -> {parameter name: dF/d(constrained value)}."""
        _, dSd = self.eval_S(typ=self.typ, grads=True)
        names = (['nu'] if self.typ == 'matern' else []) + ['kappa', 'sigma_f']
        return {n: float(dS @ dSd[n]) for n in names}

    def _adjoint(self, XA, XB, Gbar, want_dX=False):
        """Reverse mode of K(XA, XB) = (XA * S) XB^T for the adjoint Gbar (A x B): parameter gradients and dF/dXA.
        T = Gbar XB (dgemm), dS_j = sum_a XA[a,j] T[a,j] (column reduction), dXA = T * S (column scaling)."""
        h = get_handle(XA.device.index)
        A, B, k = XA.shape[0], XB.shape[0], XA.shape[1]
        T = torch.empty((A, k), dtype=torch.float64, device=XA.device)
        split, sws = _splitk(B, A * k, XA.device)
        _dgemm(h, A, k, B, Gbar, Gbar.stride(0), 1, XB, XB.stride(0), 0, T, T.stride(0), split_k=split, ws=sws)
        dS = torch.empty(k, dtype=torch.float64, device=XA.device)
        ws = torch.empty(max(1, h.query("rvgp_coldot_workspace_bytes", I64(A), int(k)) // 8), dtype=torch.float64,
                         device=XA.device)
        h.call("rvgp_coldot_f64", I64(A), int(k), XA, I64(XA.stride(0)), T, I64(T.stride(0)), dS, ws)
        grads = self._chain(dS.cpu().numpy())
        if not want_dX:
            return grads, None
        S = torch.from_numpy(self.eval_S(typ=self.typ)).to(XA.device)
        h.call("rvgp_colscale_f64", I64(A), int(k), T, I64(T.stride(0)), S)
        return grads, T

    def _diag_adjoint_uniform(self, X, g, cache):
        """Reverse mode of sum_f g * K_diag(X)[f] with one scalar weight g (the SGPR trace term):
        dS_j = g * sum_f X[f,j]^2; the column sums are data-only and cached by the caller."""
        if "colsq" not in cache:
            h = get_handle(X.device.index)
            k = X.shape[1]
            out = torch.empty(k, dtype=torch.float64, device=X.device)
            ws = torch.empty(max(1, h.query("rvgp_coldot_workspace_bytes", I64(X.shape[0]), int(k)) // 8),
                             dtype=torch.float64, device=X.device)
            h.call("rvgp_coldot_f64", I64(X.shape[0]), int(k), X, I64(X.stride(0)), X, I64(X.stride(0)), out, ws)
            cache["colsq"] = out.cpu().numpy()
        return self._chain(g * cache["colsq"])


def _splitk(K, out_elems, dev):
    """Deterministic split-K factor (and workspace) for products with a long contraction and a small output."""
    split = max(1, min(64, K // 2048)) if K >= 4096 else 1
    return split, (torch.empty(split * out_elems, dtype=torch.float64, device=dev) if split > 1 else None)


def _sq_row_norms(h, X):
    out = torch.empty(X.shape[0], dtype=torch.float64, device=X.device)
    ones = torch.ones(X.shape[1], dtype=torch.float64, device=X.device)
    h.call("rvgp_kdiag_f64", X, I64(X.stride(0)), I64(X.shape[0]), int(X.shape[1]), ones, out)
    return out


class RBF:
    """Stand-in for ``gpflow.kernels.RBF()`` (= SquaredExponential), the channel-wise baseline of
    ``train_gp(kernel='rbf')`` (RVGP/main.py:33-37):  K = variance * exp(-0.5 |x/l - x'/l|^2), K_diag = variance,
    parameters ``variance`` and ``lengthscales`` (both 1.0, positive transform with the default positive minimum at
    construction time).  Inner products on the FP64 dgemm, the elementwise part and its adjoint in csrc/gp_rbf.cu."""

    typ = 'rbf'

    def __init__(self, variance=1.0, lengthscales=1.0):
        self.variance = P.Parameter(variance, transform=P.positive(), name='variance')
        self.lengthscales = P.Parameter(lengthscales, transform=P.positive(), name='lengthscales')

    @property
    def trainable_parameters(self):
        return [p for p in (self.variance, self.lengthscales) if p.trainable]

    def _gram(self, XA, XB, out=None):
        h = get_handle(XA.device.index)
        if out is None:
            out = torch.empty((XA.shape[0], XB.shape[0]), dtype=torch.float64, device=XA.device)
        _dgemm(h, XA.shape[0], XB.shape[0], XA.shape[1], XA, XA.stride(0), 1, XB, XB.stride(0), 1, out, out.stride(0))
        xa2 = _sq_row_norms(h, XA)
        xb2 = xa2 if XB is XA else _sq_row_norms(h, XB)
        h.call("rvgp_rbf_from_dot_f64", int(XA.shape[0]), int(XB.shape[0]), out, I64(out.stride(0)), xa2, xb2,
               float(self.variance.value), float(self.lengthscales.value), out, I64(out.stride(0)))
        return out

    def _diag(self, X):
        return torch.full((X.shape[0],), self.variance.value, dtype=torch.float64, device=X.device)

    def K(self, X, X2=None):
        Xd = to_device_f64(X).contiguous()
        X2d = Xd if X2 is None else to_device_f64(X2, Xd.device).contiguous()
        return self._gram(Xd, X2d)

    def K_diag(self, X):
        return self._diag(to_device_f64(X))

    def __call__(self, X, X2=None, full_cov=True):
        return self.K(X, X2) if full_cov else self.K_diag(X)

    def _adjoint(self, XA, XB, Gbar, want_dX=False):
        """Reverse mode of K(XA, XB) for the adjoint Gbar: {variance, lengthscales} gradients and dF/dXA.
        The inner products are recomputed (one dgemm) instead of being kept from the forward pass; H = Gbar o K overwrites
        that scratch, Gbar is left untouched."""
        h = get_handle(XA.device.index)
        A, B, k = XA.shape[0], XB.shape[0], XA.shape[1]
        dev = XA.device
        Pm = torch.empty((A, B), dtype=torch.float64, device=dev)
        _dgemm(h, A, B, k, XA, XA.stride(0), 1, XB, XB.stride(0), 1, Pm, Pm.stride(0))
        xa2 = _sq_row_norms(h, XA)
        xb2 = xa2 if XB is XA else _sq_row_norms(h, XB)
        wsb = h.query("rvgp_rbf_adjoint_workspace_bytes", int(A), int(B))
        ws = torch.empty(max(8, wsb), dtype=torch.uint8, device=dev)
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        rowsum = torch.empty(A, dtype=torch.float64, device=dev) if want_dX else None
        var, ls = float(self.variance.value), float(self.lengthscales.value)
        h.call("rvgp_rbf_adjoint_f64", int(A), int(B), Gbar, I64(Gbar.stride(0)), Pm, I64(Pm.stride(0)), xa2, xb2, var, ls,
               Pm if want_dX else None, I64(Pm.stride(0)), rowsum, sums, ws, I64(wsb))
        sh = sums.cpu().numpy()
        grads = {"variance": float(sh[0] / var), "lengthscales": float(sh[1] / ls)}
        if not want_dX:
            return grads, None
        HX = torch.empty((A, k), dtype=torch.float64, device=dev)
        split, sws = _splitk(B, A * k, dev)
        _dgemm(h, A, k, B, Pm, Pm.stride(0), 1, XB, XB.stride(0), 0, HX, HX.stride(0), split_k=split, ws=sws)   # H XB
        dX = torch.empty((A, k), dtype=torch.float64, device=dev)
        h.call("rvgp_rbf_dx_f64", I64(A), int(k), HX, I64(HX.stride(0)), rowsum, XA, I64(XA.stride(0)), ls, dX,
               I64(dX.stride(0)))
        return grads, dX

    def _diag_adjoint_uniform(self, X, g, cache):
        return {"variance": float(g * X.shape[0]), "lengthscales": 0.0}
