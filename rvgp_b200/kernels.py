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
