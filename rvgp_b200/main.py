"""Mirror of the reference's ``RVGP/main.py``: ``train_gp`` (exported as ``RVGP.fit``), ``optimize_model_with_scipy``,
``manifold_GPR`` and ``manifold_SGPR`` with ``transform`` / ``predict_f`` -- same names, argument meaning, prints and
quirks (main.py:11-137), GPflow/TensorFlow replaced by rvgp_b200.gp.DeviceGPR (spectral
kernel, rank-k / dense), rvgp_b200.gp_general.DenseGPR (``kernel='rbf'``, main.py:33-37) and
rvgp_b200.gp_general.DeviceSGPR (``n_inducing_points``, main.py:59-67,119-137).

Reproduced on purpose (parity beats intent):
  * ``noise_variance`` is accepted and DROPPED: the likelihood variance starts at GPflow's default 1.0 and is
    trainable with lower bound 1e-6 (main.py:98-100).
  * kernel parameters are created BEFORE ``set_default_positive_minimum(positivity_constraint)`` runs, so the
    first fit of a process uses lower bound 0 and later fits 1e-2 (main.py:30 vs :52).
  * the second output of ``transform`` is a VARIANCE (main.py:113-116).
Accepted superset: ``transform`` also takes NumPy integer arrays, boolean masks and lists (the reference's type
sniffing fails on them, App. B.1).
"""
import numpy as np
import scipy.optimize
import torch
from sklearn.model_selection import train_test_split

from . import params as P
from .geometry import furthest_point_sampling, gather_rows_device, to_device_f64  # noqa: F401
from ._cabi import RvgpError
from . import _nvtx
from .gp import DeviceGPR
from .gp_general import DenseGPR, DeviceSGPR
from .kernels import ManifoldKernel, RBF


class _Gaussian:
    """gpflow.likelihoods.Gaussian: variance Parameter, default 1.0, lower bound 1e-6."""

    def __init__(self, variance=1.0, variance_lower_bound=1e-6):
        self.variance = P.Parameter(variance, transform=P.positive(lower=variance_lower_bound), name='variance')


def _node_rows_device(data, node_ind):
    """evecs_Lc.reshape(n, D*k)[node_ind].reshape(-1, k) on the device (main.py:31,104-106)."""
    Phi = data.device_array("evecs_Lc")
    n = data.n
    D = Phi.shape[0] // n
    idx = torch.from_numpy(np.ascontiguousarray(node_ind, dtype=np.int32)).to(Phi.device)
    out = torch.empty((len(node_ind) * D, Phi.shape[1]), dtype=torch.float64, device=Phi.device)
    from ._cabi import get_handle, I64
    h = get_handle(Phi.device.index)
    h.call("rvgp_gather_rows_f64", I64(out.shape[0]), int(Phi.shape[1]), Phi, I64(Phi.stride(0)), idx, int(D), out,
           I64(out.stride(0)))
    return out


def _as_node_indices(test_ind, n):
    """Node indices as the NumPy-based reference would resolve them (main.py:31,104): negative indices wrap, anything
    outside [-n, n) raises IndexError, a boolean mask must have n entries.  The device gather kernels do no bounds checks,
    so nothing unvalidated may reach them."""
    a = np.asarray(test_ind)
    if a.dtype == bool:
        a = a.reshape(-1)
        if a.size != n:
            raise IndexError("boolean index did not match indexed array along axis 0; size of axis is %d but size of "
                             "corresponding boolean axis is %d" % (n, a.size))
        return np.nonzero(a)[0]
    if a.size and a.dtype.kind not in "iu":
        raise IndexError("arrays used as indices must be of integer (or boolean) type")
    a = a.reshape(-1).astype(np.int64)
    bad = (a < -n) | (a >= n)
    if bad.any():
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (int(a[bad][0]), n))
    return np.where(a < 0, a + n, a)


def _rows_of(data, name, node_ind):
    """``getattr(data, name).reshape(n, -1)[node_ind]`` on the device: (len(node_ind), cols) float64."""
    A = data.device_array(name)
    idx = torch.from_numpy(np.ascontiguousarray(node_ind, dtype=np.int64)).to(A.device)
    return A.reshape(data.n, -1).index_select(0, idx).contiguous()


class _Model:
    """What manifold_GPR and manifold_SGPR share: parameters, the L-BFGS-B objective, ``transform``."""

    inducing = None          # SGPR: DeviceSGPR (its Z is trained with the hyper-parameters, GPflow's default)

    @property
    def trainable_parameters(self):
        ps = list(self.kernel.trainable_parameters)
        if self.likelihood.variance.trainable:
            ps.append(self.likelihood.variance)
        return ps

    def training_loss(self):
        return -self.maximum_log_likelihood_objective()

    # ---- variables as one host vector: [unconstrained scalars..., Z.ravel()] -------------------------------------
    def _pack(self, plist):
        u = np.array([p.unconstrained for p in plist], dtype=np.float64)
        if self.inducing is not None:
            u = np.concatenate([u, self.inducing.Z.cpu().numpy().ravel()])
        return u

    def _unpack(self, v, plist):
        for p, ui in zip(plist, v[:len(plist)]):
            p.unconstrained = float(ui)
        if self.inducing is not None:
            Z = torch.from_numpy(np.ascontiguousarray(v[len(plist):]).reshape(tuple(self.inducing.Z.shape)))
            self.inducing.Z.copy_(Z)

    def _loss_and_grad(self, v, plist):
        v = np.asarray(v, dtype=np.float64)
        self._unpack(v, plist)
        try:
            with np.errstate(all="ignore"):
                f, kgrads, gnoise, dZ = self._evaluate()
                vals = list(kgrads.values()) + [gnoise, f]
                if not np.all(np.isfinite(vals)):
                    raise FloatingPointError("non-finite objective")
        except (ArithmeticError, RvgpError):          # FloatingPointError, ZeroDivisionError (kappa == 0 in pure-Python floats), OverflowError
            # a trial point of the line search left the domain (non-finite density / non-SPD Gram): report a huge
            # loss so L-BFGS-B backs off (TensorFlow would raise here and abort the fit)
            return 1e50, np.zeros(len(v))
        g = []
        for p in plist:
            d = gnoise if p is self.likelihood.variance else kgrads[p.name]
            g.append(-d * p.transform.dforward(p.unconstrained))
        g = np.array(g, dtype=np.float64)
        if dZ is not None:
            g = np.concatenate([g, -dZ.cpu().numpy().ravel()])
        return -f, g

    def transform(self, data, test_ind, as_device=False):
        first = test_ind[0]
        if isinstance(first, (float, np.floating)) or (hasattr(first, "__len__") and np.asarray(test_ind).dtype.kind == "f"):
            test_x = to_device_f64(np.asarray(test_ind))                 # positional encodings given directly (main.py:108-109)
            n_out = test_x.shape[0]
        else:
            # like the reference, ALWAYS rows of evecs_Lc (main.py:104-106), whatever the kernel was trained on
            nodes = _as_node_indices(test_ind, data.n)
            comm = getattr(self, "comm", None)
            if comm is not None and comm.world > 1:
                return self._transform_sharded(data, nodes, comm, as_device)
            test_x = _node_rows_device(data, nodes)
            n_out = len(nodes)
        with _nvtx.stage('transform:predict_f'):
            f_pred_mean, f_pred_std = self.predict_f(test_x)
        if as_device:       # bench.py's device-resident leg: keep the result in HBM
            return f_pred_mean.tensor.reshape(n_out, -1), f_pred_std.tensor.reshape(n_out, -1)
        f_pred_mean = f_pred_mean.numpy().reshape(n_out, -1)
        f_pred_std = f_pred_std.numpy().reshape(n_out, -1)
        return f_pred_mean, f_pred_std


def _rank_chunk(n_items, comm):
    """Contiguous share of range(n_items) for this rank and the share sizes of all ranks."""
    bounds = [(n_items * r) // comm.world for r in range(comm.world + 1)]
    return bounds[comm.rank], bounds[comm.rank + 1], [bounds[r + 1] - bounds[r] for r in range(comm.world)]


def _transform_sharded(self, data, nodes, comm, as_device):
    """Row-sharded ``transform`` (SPMD: every rank calls it with the same arguments): each rank predicts its contiguous share
    of the test nodes, one all-gather returns the complete (n_test, D) mean / variance on every rank."""
    a, b, counts = _rank_chunk(len(nodes), comm)
    test_x = _node_rows_device(data, nodes[a:b])
    with _nvtx.stage('transform:predict_f'):
        m, v = self.predict_f(test_x)
    D = data.device_array("evecs_Lc").shape[0] // data.n
    both = torch.cat([m.tensor.reshape(b - a, D), v.tensor.reshape(b - a, D)], dim=1).contiguous()
    full = comm.allgather_rows(both, counts)
    mean, var = full[:, :D].contiguous(), full[:, D:].contiguous()
    if as_device:
        return mean, var
    return mean.cpu().numpy(), var.cpu().numpy()


_Model._transform_sharded = _transform_sharded


class manifold_GPR(_Model):
    """GP regression model (replaces the gpflow.models.GPR subclass, main.py:98-116)."""

    def __init__(self, data, kernel, mean_function=None, noise_variance=None, likelihood=None, solver="auto", comm=None):
        # like the reference, mean_function / noise_variance / likelihood are ignored (main.py:100)
        X, Y = data
        self.data = (to_device_f64(X), to_device_f64(Y))
        self.kernel = kernel
        self.likelihood = _Gaussian()
        self.comm = comm
        self._spectral = isinstance(kernel, ManifoldKernel) and self.data[1].shape[1] == 1 and solver != "general"
        if self._spectral:
            self._gpr = DeviceGPR(self.data[0], self.data[1], solver=solver, comm=comm)
            self.solver = self._gpr.solver
        else:
            self._gpr = DenseGPR(self.data[0], self.data[1], kernel)
            self.solver = "general"

    def log_marginal_likelihood(self):
        if self._spectral:
            S = self.kernel.eval_S(typ=self.kernel.typ)
            return self._gpr.lml_and_grads(S, self.likelihood.variance.value, grads=False)
        return self._gpr.lml_and_grads(self.likelihood.variance.value, grads=False)

    maximum_log_likelihood_objective = log_marginal_likelihood

    def _evaluate(self):
        if self._spectral:
            S = self.kernel.eval_S(typ=self.kernel.typ)
            if not np.all(np.isfinite(S)) or np.any(S <= 0):
                raise FloatingPointError("non-finite spectral density")
            lml, gS, gnoise = self._gpr.lml_and_grads(S, self.likelihood.variance.value, grads=True)
            if not np.all(np.isfinite(gS)):
                raise FloatingPointError("non-finite LML gradient")
            return lml, self.kernel._chain(gS), gnoise, None
        lml, kgrads, gnoise = self._gpr.lml_and_grads(self.likelihood.variance.value, grads=True)
        return lml, kgrads, gnoise, None

    def predict_f(self, Xnew, full_cov=False):
        """(mean (N*,R), variance (N*,R)) as CUDA tensors with a ``.numpy()``-like host path via .cpu()."""
        if self._spectral:
            S = self.kernel.eval_S(typ=self.kernel.typ)
            mean, var = self._gpr.predict(S, self.likelihood.variance.value, to_device_f64(Xnew))
        else:
            mean, var = self._gpr.predict(self.likelihood.variance.value, to_device_f64(Xnew))
        return _HostView(mean), _HostView(var)


class manifold_SGPR(_Model):
    """Sparse GP regression (replaces the gpflow.models.SGPR subclass, main.py:119-137): Titsias' collapsed bound with
    trainable inducing points; like the reference, mean_function / noise_variance / likelihood are ignored (:121)."""

    def __init__(self, data, kernel, inducing_variable, mean_function=None, noise_variance=None, likelihood=None):
        X, Y = data
        self.data = (to_device_f64(X), to_device_f64(Y))
        self.kernel = kernel
        self.likelihood = _Gaussian()
        self.inducing = DeviceSGPR(self.data[0], self.data[1], to_device_f64(inducing_variable), kernel)
        self.solver = "sgpr"

    @property
    def inducing_variable(self):
        return _HostView(self.inducing.Z)

    def elbo(self):
        return self.inducing.elbo_and_grads(self.likelihood.variance.value, grads=False)

    maximum_log_likelihood_objective = elbo

    def _evaluate(self):
        return self.inducing.elbo_and_grads(self.likelihood.variance.value, grads=True)

    def predict_f(self, Xnew, full_cov=False):
        mean, var = self.inducing.predict(self.likelihood.variance.value, to_device_f64(Xnew))
        return _HostView(mean), _HostView(var)


class _HostView:
    """Wraps a CUDA tensor; ``.numpy()`` copies to the host (what the reference calls on GPflow's outputs)."""

    def __init__(self, t):
        self.tensor = t

    def numpy(self):
        return self.tensor.cpu().numpy()

    def __dlpack__(self, *a, **k):
        return self.tensor.__dlpack__(*a, **k)

    def __dlpack_device__(self):
        return self.tensor.__dlpack_device__()


def optimize_model_with_scipy(model, epochs):
    """gpflow.optimizers.Scipy().minimize(training_loss, trainable_variables, method="l-bfgs-b",
    options={"disp": True, "maxiter": epochs})  (main.py:87-95).  4 scalars on the host; every evaluation's
    linear algebra on the device."""
    plist = model.trainable_parameters
    if not plist and model.inducing is None:
        return model
    u0 = model._pack(plist)
    from .eigensolver import _lapack_ctx
    # a FEW BLAS threads for the k-vector / k x k host glue of every evaluation (measured for k = 500: 0.2 ms per evaluation
    # with 4 threads, 0.6 ms with torchrun's single thread, 0.9 ms with OpenBLAS's default 16 on the B200 host)
    with _lapack_ctx():
        res = scipy.optimize.minimize(lambda u: model._loss_and_grad(u, plist), u0, jac=True, method="L-BFGS-B",
                                      options={"maxiter": epochs})
    model._unpack(res.x, plist)
    model.opt_result = res
    return model


def train_gp(data,
             train_ind=None,
             n_inducing_points=None,
             test_size=0.2,
             kernel=None,
             noise_variance=0.001,
             kernel_lengthscale=None,
             kernel_variance=None,
             epochs=1000,
             positivity_constraint=1e-2,
             seed=0,
             solver="auto"):

    if train_ind is None:
        train_ind = np.arange(data.n)
    train_nodes = _as_node_indices(train_ind, data.n)

    vec = data._duals["vectors"].get_dev(data.device) if hasattr(data, "_duals") else to_device_f64(data.vectors)
    vd = vec.reshape(data.n, -1)
    feature = "evecs_Lc"
    dim = vd.shape[1]
    if kernel is None:
        kernel = ManifoldKernel(data, nu=3 / 2, kappa=5, typ='matern', sigma_f=1.)
    if isinstance(kernel, str) and kernel == 'rbf':
        kernel = RBF()
        feature = "evecs_L"                 # main.py:35: rows of the SCALAR Laplacian's eigenvectors, one row per node
        dim = 1
        print('Using RBF kernel, treating vectors channel-wise.')

    # split training and test set: sklearn on an index array gives the same rows as splitting the arrays
    # (main.py:40-45); bit-exact host RNG, only indices go to the GPU
    tr, te = train_test_split(np.arange(len(train_nodes)), test_size=test_size, random_state=seed)
    # multi-GPU (data.sharded: SPMD, every rank makes the same calls): the rank-k spectral GP is sharded over TRAINING NODES --
    # each rank gathers / multiplies only its contiguous share of the rows, one all-reduce of Phi^T [Phi y] joins them, and every
    # rank then runs the identical 4-scalar host optimisation.  Other model kinds stay replicated.
    comm = None
    if (getattr(data, "sharded", False) and feature == "evecs_Lc" and n_inducing_points is None
            and isinstance(kernel, ManifoldKernel) and dim * len(tr) >= 2 * data.device_array("evecs_Lc").shape[1]
            and solver in ("auto", "lowrank")):
        from .distributed import Comm
        comm = Comm()
        if comm.world > 1:
            a, b, _ = _rank_chunk(len(tr), comm)
            tr = tr[a:b]
            a2, b2, _ = _rank_chunk(len(te), comm)
            te = te[a2:b2]
        else:
            comm = None
    if feature == "evecs_Lc":
        in_train = _node_rows_device(data, train_nodes[tr])             # (n_tr * D, k)
        in_test = _node_rows_device(data, train_nodes[te])
    else:
        in_train = _rows_of(data, "evecs_L", train_nodes[tr])           # (n_tr, k_L)
        in_test = _rows_of(data, "evecs_L", train_nodes[te])
    idx_tr = torch.from_numpy(train_nodes[tr]).to(vd.device)
    idx_te = torch.from_numpy(train_nodes[te]).to(vd.device)
    out_train = vd.index_select(0, idx_tr).reshape(len(tr) * dim, -1)
    out_test = vd.index_select(0, idx_te).reshape(len(te) * dim, -1)

    P.set_default_positive_minimum(positivity_constraint)

    if n_inducing_points is None:
        GP = manifold_GPR((in_train, out_train), kernel, noise_variance=noise_variance, solver=solver, comm=comm)
    else:
        # inducing points: furthest-point sampling in FEATURE space (main.py:60-61), on the device (K1, D = k)
        from .fps import furthest_point_sampling_device
        ind, _ = furthest_point_sampling_device(in_train, N=int(n_inducing_points))
        inducing_variable = in_train.index_select(0, ind.to(torch.int64))
        GP = manifold_SGPR((in_train, out_train), kernel, inducing_variable, noise_variance=noise_variance)

    if kernel_variance is not None:
        kernel.variance.assign(kernel_variance)          # AttributeError for ManifoldKernel, as in the reference
        P.set_trainable(kernel.variance, False)
    if kernel_lengthscale is not None:
        kernel.lengthscales.assign(kernel_lengthscale)
        P.set_trainable(kernel.lengthscales, False)

    with _nvtx.stage('fit:optimise'):
        GP = optimize_model_with_scipy(GP, epochs)

    # test
    out_pred, _ = GP.predict_f(in_test)
    if comm is not None:
        # mean over ALL held-out rows: sum of this rank's row norms, one small all-reduce
        part = torch.linalg.vector_norm(out_test - out_pred.tensor, dim=1).sum().reshape(1)
        cnt = torch.tensor([float(out_test.shape[0])], dtype=torch.float64, device=part.device)
        both = torch.cat([part, cnt])
        comm.allreduce_(both)
        l2_error = float(both[0].item() / max(both[1].item(), 1.0))
    else:
        l2_error = np.linalg.norm(out_test.cpu().numpy() - out_pred.numpy(), axis=1).mean()
    print("Relative l2 error is {}".format(l2_error))
    GP.l2_error = float(l2_error)
    return GP
