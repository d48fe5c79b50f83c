"""CPU restatement of the reference's GP half.  TEST INFRASTRUCTURE ONLY (see oracle/rvgp_oracle.py).

PARITY STATUS: **parity unpinned**.  The arithmetic of this half lives in third-party packages that are
absent from /root/reference and cannot be installed here (no network): gpflow==2.6.5, tensorflow==2.12.0,
tensorflow-probability==0.20.0 (reference setup.py:11-13).  The reference ships no test, golden vector or
fixture for this boundary.  What follows restates GPflow 2.6.5's published algorithm, anchored on the
reference's own call sites:

  RVGP/kernels.py:25-67    ManifoldKernel (spectral density S, K = (X*S) X2^T, K_diag)
  RVGP/main.py:11-84       train_gp: data prep, 80/20 split, manifold_GPR, L-BFGS-B, held-out error
  RVGP/main.py:98-116      manifold_GPR.__init__ (drops noise_variance -> GPflow default 1.0) and transform
  gpflow/models/gpr.py     GPR.log_marginal_likelihood, GPR.predict_f
  gpflow/conditionals/util.py  base_conditional
  gpflow/logdensities.py   multivariate_normal
  gpflow/utilities/bijectors.py  positive(): softplus (+ Shift(lower) when lower != 0)
  gpflow/likelihoods       Gaussian(variance=1.0, lower bound 1e-6)
  gpflow/optimizers/scipy.py     scipy.optimize.minimize(jac=True, method="l-bfgs-b")

Independent third-party cross-pin: tests/test_gp_oracle_vs_sklearn.py compares the GPR equations below with scikit-learn's
GaussianProcessRegressor (available here) to 1e-9 or better -- not GPflow, hence still 'unpinned'.
Cross-checks that stand in for pins (tests/test_gp_oracle.py): analytic gradients vs central differences,
K_diag == diag(K), dense Cholesky GPR == rank-k (Woodbury) GPR, the survey's independent probe value of the
initial LML on config C1 (-2512.1283414532).
"""
import numpy as np
import scipy.linalg
import scipy.optimize

LOG2PI = np.log(2.0 * np.pi)


def softplus(u):
    return np.logaddexp(0.0, u)


def softplus_inv(y):
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))


def sigmoid(u):
    return 1.0 / (1.0 + np.exp(-u))


def eval_S(evals, nu, kappa, sigma_f, num_verticies, typ="matern", grads=False):
    """kernels.py:41-53.  With grads=True also returns dS/dnu, dS/dkappa, dS/dsigma_f (k-vectors)."""
    lam = np.asarray(evals, dtype=np.float64)
    if typ == "matern":
        a = 2.0 * nu / kappa ** 2
        S0 = np.power(lam + a, -nu)
        dl_nu = -np.log(lam + a) - nu * (2.0 / kappa ** 2) / (lam + a)
        dl_kappa = 4.0 * nu ** 2 / (kappa ** 3 * (lam + a))
    elif typ == "se":
        S0 = np.exp(-0.5 * lam * kappa ** 2)
        dl_nu = np.zeros_like(lam)
        dl_kappa = -lam * kappa
    else:
        raise NotImplementedError(typ)
    Z = S0.sum()
    S = S0 * (num_verticies / Z) * sigma_f
    if not grads:
        return S
    p = S0 / Z
    dS_nu = S * (dl_nu - (p * dl_nu).sum())
    dS_kappa = S * (dl_kappa - (p * dl_kappa).sum())
    dS_sf = S / sigma_f
    return S, dS_nu, dS_kappa, dS_sf


def K(X, S, X2=None):
    """kernels.py:55-61."""
    X2 = X if X2 is None else X2
    return (X * S) @ X2.T


def K_diag(X, S):
    """kernels.py:63-67 (the reference forms the full matrix; the value is the same)."""
    return ((X * S) * X).sum(1)


def gpr_lml_dense(X, Y, S, noise, grads=False):
    """GPR.log_marginal_likelihood with zero mean (gpflow/models/gpr.py, logdensities.multivariate_normal).
    grads=True: also dLML/dS (k,) and dLML/dnoise."""
    M = X.shape[0]
    Ky = K(X, S) + noise * np.eye(M)
    L = np.linalg.cholesky(Ky)
    alpha = scipy.linalg.solve_triangular(L, Y, lower=True)
    lml = float(-0.5 * np.sum(alpha ** 2) - 0.5 * M * LOG2PI * Y.shape[1] - Y.shape[1] * np.sum(np.log(np.diag(L))))
    if not grads:
        return lml
    a = scipy.linalg.solve_triangular(L, alpha, lower=True, trans="T")     # Ky^-1 y
    Z = scipy.linalg.solve_triangular(L, X, lower=True)                     # L^-1 Phi
    wdiag = (Z * Z).sum(0)                                                  # phi_j^T Ky^-1 phi_j
    u = X.T @ a                                                             # Phi^T Ky^-1 y
    dS = 0.5 * (u[:, 0] ** 2) - 0.5 * wdiag
    tr_inv = (M - (S * wdiag).sum()) / noise
    dnoise = 0.5 * float((a * a).sum()) - 0.5 * tr_inv
    return lml, dS, dnoise


def gpr_lml_lowrank(G, b, yy, M, S, noise, grads=False):
    """Same quantity through the rank-k identity (K = Phi S Phi^T has rank <= k):
    G = Phi^T Phi, b = Phi^T y, yy = y^T y.  Single output column."""
    k = G.shape[0]
    rs = np.sqrt(S)
    B = np.eye(k) + (rs[:, None] * G * rs[None, :]) / noise
    Lb = np.linalg.cholesky(B)
    bt = rs * b
    z = scipy.linalg.cho_solve((Lb, True), bt)
    quad = (yy - (bt @ z) / noise) / noise
    lml = float(-0.5 * quad - 0.5 * M * LOG2PI - 0.5 * M * np.log(noise) - np.sum(np.log(np.diag(Lb))))
    if not grads:
        return lml
    c = rs * z
    Gc = G @ c
    u = (b - Gc / noise) / noise
    Q = scipy.linalg.solve_triangular(Lb, rs[:, None] * G, lower=True)
    wdiag = (np.diag(G) - (Q * Q).sum(0) / noise) / noise
    dS = 0.5 * u ** 2 - 0.5 * wdiag
    aa = (yy - 2.0 * (b @ c) / noise + (c @ Gc) / noise ** 2) / noise ** 2
    tr_inv = (M - (S * wdiag).sum()) / noise
    dnoise = 0.5 * aa - 0.5 * tr_inv
    return lml, dS, dnoise


def gpr_predict_dense(X, Y, S, noise, Xnew):
    """GPR.predict_f(full_cov=False) via base_conditional (gpflow/conditionals/util.py)."""
    M = X.shape[0]
    Kmm = K(X, S) + noise * np.eye(M)
    Kmn = K(X, S, Xnew)
    Knn = K_diag(Xnew, S)
    Lm = np.linalg.cholesky(Kmm)
    A = scipy.linalg.solve_triangular(Lm, Kmn, lower=True)
    fvar = Knn - (A * A).sum(0)
    A = scipy.linalg.solve_triangular(Lm, A, lower=True, trans="T")
    fmean = A.T @ Y
    return fmean, np.tile(fvar[:, None], [1, Y.shape[1]])


def gpr_predict_lowrank(G, b, S, noise, Xnew):
    k = G.shape[0]
    rs = np.sqrt(S)
    B = np.eye(k) + (rs[:, None] * G * rs[None, :]) / noise
    Lb = np.linalg.cholesky(B)
    wbar = rs * scipy.linalg.cho_solve((Lb, True), rs * b) / noise
    Q = scipy.linalg.solve_triangular(Lb, np.diag(rs), lower=True)          # L_b^-1 S^1/2
    T = Xnew @ Q.T
    return (Xnew @ wbar)[:, None], (T * T).sum(1)[:, None]


def train_test_split_indices(n_samples, test_size=0.2, seed=0):
    """sklearn.model_selection.train_test_split(test_size, random_state=seed) index restatement
    (main.py:40-45).  Returns (train_idx, test_idx)."""
    n_test = int(np.ceil(test_size * n_samples))
    n_train = n_samples - n_test
    perm = np.random.RandomState(seed).permutation(n_samples)
    return perm[n_test:n_test + n_train], perm[:n_test]


class OracleGPR:
    """manifold_GPR restated (main.py:98-116)."""

    def __init__(self, X, Y, evals, num_verticies, nu=1.5, kappa=5.0, sigma_f=1.0, typ="matern",
                 kernel_lower=0.0, noise=1.0, noise_lower=1e-6, solver="dense"):
        self.X, self.Y, self.evals, self.nv, self.typ = X, Y, np.asarray(evals), float(num_verticies), typ
        self.kernel_lower, self.noise_lower, self.solver = kernel_lower, noise_lower, solver
        self.names = (["nu"] if typ == "matern" else []) + ["kappa", "sigma_f", "noise"]
        init = dict(nu=nu, kappa=kappa, sigma_f=sigma_f, noise=noise)
        self.u = np.array([softplus_inv(init[n] - self._lower(n)) for n in self.names], dtype=np.float64)
        if solver == "lowrank":
            self.G, self.b, self.yy = X.T @ X, X.T @ Y[:, 0], float(Y[:, 0] @ Y[:, 0])
        self.n_eval = 0

    def _lower(self, name):
        return self.noise_lower if name == "noise" else self.kernel_lower

    def params(self, u=None):
        u = self.u if u is None else u
        p = {n: self._lower(n) + softplus(ui) for n, ui in zip(self.names, u)}
        p.setdefault("nu", 0.0)
        return p

    def loss_and_grad(self, u):
        """training_loss = -LML and its gradient w.r.t. the unconstrained variables."""
        self.n_eval += 1
        try:
            with np.errstate(all="ignore"):
                return self._loss_and_grad(u)
        except (np.linalg.LinAlgError, ValueError, FloatingPointError):
            return 1e50, np.zeros_like(u)     # non-finite / non-SPD trial point: make the line search back off

    def _loss_and_grad(self, u):
        p = self.params(u)
        S, dnu, dka, dsf = eval_S(self.evals, p["nu"], p["kappa"], p["sigma_f"], self.nv, self.typ, grads=True)
        if self.solver == "lowrank":
            lml, dS, dnoise = gpr_lml_lowrank(self.G, self.b, self.yy, self.X.shape[0], S, p["noise"], grads=True)
        else:
            lml, dS, dnoise = gpr_lml_dense(self.X, self.Y, S, p["noise"], grads=True)
        dtheta = dict(nu=dS @ dnu, kappa=dS @ dka, sigma_f=dS @ dsf, noise=dnoise)
        g = np.array([dtheta[n] * sigmoid(ui) for n, ui in zip(self.names, u)])
        if not (np.isfinite(lml) and np.all(np.isfinite(g))):
            raise FloatingPointError("non-finite LML")
        return -lml, -g

    def fit(self, epochs=1000, disp=False):
        res = scipy.optimize.minimize(self.loss_and_grad, self.u, jac=True, method="L-BFGS-B",
                                      options={"maxiter": epochs})
        self.u = res.x
        self.opt_result = res
        return self

    def predict_f(self, Xnew):
        p = self.params()
        S = eval_S(self.evals, p["nu"], p["kappa"], p["sigma_f"], self.nv, self.typ)
        if self.solver == "lowrank":
            return gpr_predict_lowrank(self.G, self.b, S, p["noise"], Xnew)
        return gpr_predict_dense(self.X, self.Y, S, p["noise"], Xnew)


def prepare_training(evecs_Lc, vectors, n, train_ind, test_size=0.2, seed=0):
    """main.py:24-50 data prep: rows of the (n, D*k) eigenvector matrix for the training nodes, 80/20 split on
    node blocks, reshape to (M, k) / (M, 1)."""
    train_ind = np.arange(n) if train_ind is None else np.asarray(train_ind)
    output = vectors[train_ind]
    inp = evecs_Lc.reshape(n, -1)[train_ind]
    dim = output.shape[1]
    tr, te = train_test_split_indices(len(inp), test_size, seed)
    k = evecs_Lc.shape[1]
    return (inp[tr].reshape(len(tr) * dim, k), output[tr].reshape(len(tr) * dim, 1),
            inp[te].reshape(len(te) * dim, k), output[te].reshape(len(te) * dim, 1))


def train_gp(evecs_Lc, evals_Lc, vectors, n, train_ind=None, epochs=1000, seed=0, kernel_lower=0.0,
             solver="dense", disp=False):
    Xtr, Ytr, Xte, Yte = prepare_training(evecs_Lc, vectors, n, train_ind, seed=seed)
    gp = OracleGPR(Xtr, Ytr, evals_Lc, evecs_Lc.shape[0], nu=1.5, kappa=5.0, sigma_f=1.0, kernel_lower=kernel_lower,
                   solver=solver)
    gp.fit(epochs, disp=disp)
    pred, _ = gp.predict_f(Xte)
    gp.l2_error = float(np.linalg.norm(Yte - pred, axis=1).mean())      # main.py:80-82
    return gp


def transform(gp, evecs_Lc, n, test_ind):
    """main.py:102-116 with integer node indices."""
    k = evecs_Lc.shape[1]
    tx = evecs_Lc.reshape(n, -1)[np.asarray(test_ind)].reshape(-1, k)
    m, v = gp.predict_f(tx)
    return m.reshape(len(test_ind), -1), v.reshape(len(test_ind), -1)


# =====================================================================================================================
# the SGPR path (main.py:59-67,119-137) and kernel='rbf' (main.py:33-37).  Same status as
# above: **parity unpinned** (GPflow 2.6.5 restated from its published algorithm: gpflow/kernels/stationaries.py
# SquaredExponential, gpflow/models/sgpr.py SGPR_deprecated.elbo / predict_f, default_jitter() = 1e-6).
# Gradients are analytic (reverse mode by hand) and checked against central differences in tests/test_gp_oracle.py.
# =====================================================================================================================
DEFAULT_JITTER = 1e-6          # gpflow.config.default_jitter()


class SpectralKernel:
    """ManifoldKernel (kernels.py:25-67) as a generic kernel object: parameters nu (matern only), kappa, sigma_f."""

    def __init__(self, evals, num_verticies, typ="matern", nu=1.5, kappa=5.0, sigma_f=1.0):
        self.evals, self.nv, self.typ = np.asarray(evals, dtype=np.float64), float(num_verticies), typ
        self.names = (["nu"] if typ == "matern" else []) + ["kappa", "sigma_f"]
        self.init = dict(nu=nu, kappa=kappa, sigma_f=sigma_f)

    def _S(self, p, grads=False):
        return eval_S(self.evals, p.get("nu", 0.0), p["kappa"], p["sigma_f"], self.nv, self.typ, grads=grads)

    def K(self, p, X, X2=None):
        return K(X, self._S(p), X2)

    def K_diag(self, p, X):
        return K_diag(X, self._S(p))

    def _chain(self, p, dS):
        S, dnu, dka, dsf = self._S(p, grads=True)
        out = {"kappa": float(dS @ dka), "sigma_f": float(dS @ dsf)}
        if self.typ == "matern":
            out["nu"] = float(dS @ dnu)
        return out

    def adjoint(self, p, XA, XB, Gbar):
        """d/dparams and d/dXA of sum(Gbar * K(XA, XB))."""
        T = Gbar @ XB
        return self._chain(p, (XA * T).sum(0)), T * self._S(p)

    def diag_adjoint(self, p, X, gbar):
        return self._chain(p, (gbar[:, None] * X * X).sum(0))


class RBFKernel:
    """gpflow.kernels.RBF() = SquaredExponential(variance=1, lengthscales=1) (main.py:34):
    K = variance * exp(-0.5 * |x/l - x'/l|^2), K_diag = variance."""

    names = ["variance", "lengthscales"]

    def __init__(self, variance=1.0, lengthscales=1.0):
        self.init = dict(variance=variance, lengthscales=lengthscales)

    @staticmethod
    def _r2(p, X, X2):
        Xs = X / p["lengthscales"]
        X2s = Xs if X2 is None else X2 / p["lengthscales"]
        return (Xs * Xs).sum(1)[:, None] + (X2s * X2s).sum(1)[None, :] - 2.0 * (Xs @ X2s.T)

    def K(self, p, X, X2=None):
        return p["variance"] * np.exp(-0.5 * self._r2(p, X, X2))

    def K_diag(self, p, X):
        return np.full(X.shape[0], p["variance"])

    def adjoint(self, p, XA, XB, Gbar):
        r2 = self._r2(p, XA, XB)                       # scaled: |xa - xb|^2 / l^2
        Kab = p["variance"] * np.exp(-0.5 * r2)
        Hm = Gbar * Kab
        grads = {"variance": float(Hm.sum() / p["variance"]), "lengthscales": float((Hm * r2).sum() / p["lengthscales"])}
        dXA = (Hm @ XB - Hm.sum(1)[:, None] * XA) / p["lengthscales"] ** 2
        return grads, dXA

    def diag_adjoint(self, p, X, gbar):
        return {"variance": float(np.sum(gbar)), "lengthscales": 0.0}


def gpr_general_lml(kern, p, X, Y, noise, grads=False):
    """GPR.log_marginal_likelihood for any kernel object and R output columns (gpflow/models/gpr.py)."""
    M, R = Y.shape
    Ky = kern.K(p, X) + noise * np.eye(M)
    L = np.linalg.cholesky(Ky)
    alpha = scipy.linalg.solve_triangular(L, Y, lower=True)
    lml = float(-0.5 * np.sum(alpha ** 2) - 0.5 * M * R * LOG2PI - R * np.sum(np.log(np.diag(L))))
    if not grads:
        return lml
    a = scipy.linalg.solve_triangular(L, alpha, lower=True, trans="T")
    Kinv = scipy.linalg.cho_solve((L, True), np.eye(M))
    W = a @ a.T - R * Kinv
    g, _ = kern.adjoint(p, X, X, 0.5 * W)
    return lml, g, 0.5 * float(np.trace(W))


def gpr_general_predict(kern, p, X, Y, noise, Xnew):
    M = X.shape[0]
    Lm = np.linalg.cholesky(kern.K(p, X) + noise * np.eye(M))
    A = scipy.linalg.solve_triangular(Lm, kern.K(p, X, Xnew), lower=True)
    fvar = kern.K_diag(p, Xnew) - (A * A).sum(0)
    A = scipy.linalg.solve_triangular(Lm, A, lower=True, trans="T")
    return A.T @ Y, np.tile(fvar[:, None], [1, Y.shape[1]])


def _sgpr_common(kern, p, X, Y, Z, noise):
    Mu = Z.shape[0]
    Kuf = kern.K(p, Z, X)
    Kuu = kern.K(p, Z) + DEFAULT_JITTER * np.eye(Mu)
    L = np.linalg.cholesky(Kuu)
    sigma = np.sqrt(noise)
    A = scipy.linalg.solve_triangular(L, Kuf, lower=True) / sigma
    B = A @ A.T + np.eye(Mu)
    LB = np.linalg.cholesky(B)
    c = scipy.linalg.solve_triangular(LB, A @ Y, lower=True) / sigma
    return L, A, B, LB, c, sigma


def sgpr_elbo(kern, p, X, Y, Z, noise, grads=False):
    """gpflow/models/sgpr.py SGPR.elbo (Titsias' collapsed bound), zero mean.  grads=True: also
    (dict of kernel-parameter gradients, d/dnoise, d/dZ)."""
    N, R = Y.shape
    Mu = Z.shape[0]
    L, A, B, LB, c, sigma = _sgpr_common(kern, p, X, Y, Z, noise)
    Kdiag = kern.K_diag(p, X)
    AAT_tr = float(np.sum(A * A))
    bound = (-0.5 * N * R * LOG2PI - R * np.sum(np.log(np.diag(LB))) - 0.5 * N * R * np.log(noise)
             - 0.5 * np.sum(Y * Y) / noise + 0.5 * np.sum(c * c) - 0.5 * R * np.sum(Kdiag) / noise + 0.5 * R * AAT_tr)
    bound = float(bound)
    if not grads:
        return bound
    # reverse mode by hand: with t = LB^-T c, gamma = L^-T t, beta = (Y - Kfu gamma) / noise,
    #   dF/dKuf = L^-T [ t beta^T + (R / sigma) (I - B^-1) A ],  dF/dKuu = L^-T [ -1/2 t t^T - R/2 (B - 2I + B^-1) ] L^-1,
    #   dF/dKdiag = -R / (2 noise)
    t = scipy.linalg.solve_triangular(LB, c, lower=True, trans="T")
    Binv = scipy.linalg.cho_solve((LB, True), np.eye(Mu))
    beta = (Y - sigma * (A.T @ t)) / noise
    Q = t @ beta.T + (R / sigma) * ((np.eye(Mu) - Binv) @ A)
    Guf = scipy.linalg.solve_triangular(L, Q, lower=True, trans="T")
    Mid = -0.5 * (t @ t.T) - 0.5 * R * (B - 2.0 * np.eye(Mu) + Binv)
    W1 = scipy.linalg.solve_triangular(L, Mid, lower=True, trans="T")
    Guu = scipy.linalg.solve_triangular(L, W1.T, lower=True, trans="T").T
    g1, dZ1 = kern.adjoint(p, Z, X, Guf)
    g2, dZ2 = kern.adjoint(p, Z, Z, Guu)
    g3 = kern.diag_adjoint(p, X, np.full(N, -0.5 * R / noise))
    g = {n: g1[n] + g2[n] + g3[n] for n in g1}
    dnoise = (0.5 * (np.sum(beta * beta) - R * (N - (Mu - np.trace(Binv))) / noise)
              + 0.5 * R * np.sum(Kdiag) / noise ** 2 - 0.5 * R * AAT_tr / noise)
    return bound, g, float(dnoise), dZ1 + 2.0 * dZ2


def sgpr_predict(kern, p, X, Y, Z, noise, Xnew):
    """SGPR.predict_f(full_cov=False)."""
    L, A, B, LB, c, sigma = _sgpr_common(kern, p, X, Y, Z, noise)
    Kus = kern.K(p, Z, Xnew)
    tmp1 = scipy.linalg.solve_triangular(L, Kus, lower=True)
    tmp2 = scipy.linalg.solve_triangular(LB, tmp1, lower=True)
    mean = tmp2.T @ c
    var = kern.K_diag(p, Xnew) + (tmp2 * tmp2).sum(0) - (tmp1 * tmp1).sum(0)
    return mean, np.tile(var[:, None], [1, Y.shape[1]])


class OracleModel:
    """manifold_GPR / manifold_SGPR (main.py:98-137) for any kernel object; Z = None -> GPR.  Trainable: the kernel
    parameters not listed in ``fixed``, the likelihood variance, and (SGPR) the inducing points Z (GPflow default)."""

    def __init__(self, kern, X, Y, Z=None, kernel_lower=0.0, noise=1.0, noise_lower=1e-6, fixed=()):
        self.kern, self.X, self.Y = kern, X, Y
        self.Z = None if Z is None else np.array(Z, dtype=np.float64)
        self.kernel_lower, self.noise_lower = kernel_lower, noise_lower
        self.fixed = {n: kern.init[n] for n in fixed}
        self.names = [n for n in kern.names if n not in self.fixed] + ["noise"]
        init = dict(kern.init, noise=noise)
        self.u = np.array([softplus_inv(init[n] - self._lower(n)) for n in self.names], dtype=np.float64)
        self.n_eval = 0

    def _lower(self, name):
        return self.noise_lower if name == "noise" else self.kernel_lower

    def params(self, u=None):
        u = self.u if u is None else u
        p = {n: self._lower(n) + softplus(ui) for n, ui in zip(self.names, u)}
        p.update(self.fixed)
        return p

    def pack(self):
        return self.u.copy() if self.Z is None else np.concatenate([self.u, self.Z.ravel()])

    def objective(self, v, grads=True):
        nu = len(self.names)
        u = v[:nu]
        p = self.params(u)
        if self.Z is None:
            out = gpr_general_lml(self.kern, p, self.X, self.Y, p["noise"], grads=grads)
            if not grads:
                return out
            f, g, dnoise = out
            dZ = None
        else:
            Z = v[nu:].reshape(self.Z.shape)
            out = sgpr_elbo(self.kern, p, self.X, self.Y, Z, p["noise"], grads=grads)
            if not grads:
                return out
            f, g, dnoise, dZ = out
        g = dict(g, noise=dnoise)
        gu = np.array([g[n] * sigmoid(ui) for n, ui in zip(self.names, u)])
        return f, (gu if dZ is None else np.concatenate([gu, dZ.ravel()]))

    def loss_and_grad(self, v):
        self.n_eval += 1
        try:
            with np.errstate(all="ignore"):
                f, g = self.objective(v)
            if not (np.isfinite(f) and np.all(np.isfinite(g))):
                raise FloatingPointError
        except (np.linalg.LinAlgError, ValueError, FloatingPointError):
            return 1e50, np.zeros_like(v)
        return -f, -g

    def fit(self, epochs=1000):
        res = scipy.optimize.minimize(self.loss_and_grad, self.pack(), jac=True, method="L-BFGS-B",
                                      options={"maxiter": epochs})
        nu = len(self.names)
        self.u = res.x[:nu]
        if self.Z is not None:
            self.Z = res.x[nu:].reshape(self.Z.shape)
        self.opt_result = res
        return self

    def predict_f(self, Xnew):
        p = self.params()
        if self.Z is None:
            return gpr_general_predict(self.kern, p, self.X, self.Y, p["noise"], Xnew)
        return sgpr_predict(self.kern, p, self.X, self.Y, self.Z, p["noise"], Xnew)


def fps_features(x, N, start_idx=0):
    """furthest_point_sampling(in_train, N=n_inducing_points) (main.py:60; geometry.py:126-162) with sklearn's
    pairwise_distances, exactly as the reference calls it."""
    from sklearn.metrics import pairwise_distances
    D = pairwise_distances(x)
    perm = np.zeros(N, dtype=np.int32)
    perm[0] = start_idx
    ds = D[start_idx, :]
    for i in range(1, N):
        idx = np.argmax(ds)
        perm[i] = idx
        ds = np.minimum(ds, D[idx, :])
    return perm


def train_gp_general(evecs_Lc, evals_Lc, evecs_L, vectors, n, train_ind=None, n_inducing_points=None, kernel=None,
                     kernel_lengthscale=None, kernel_variance=None, epochs=1000, seed=0, kernel_lower=0.0, test_size=0.2):
    """main.py:11-84 for every branch: kernel None | 'rbf', GPR | SGPR."""
    train_ind = np.arange(n) if train_ind is None else np.asarray(train_ind)
    output = vectors[train_ind]
    if kernel == "rbf":
        kern = RBFKernel()
        inp = evecs_L.reshape(n, -1)[train_ind]
        dim = 1
    else:
        kern = SpectralKernel(evals_Lc, evecs_Lc.shape[0], "matern", nu=1.5, kappa=5.0, sigma_f=1.0)
        inp = evecs_Lc.reshape(n, -1)[train_ind]
        dim = output.shape[1]
    tr, te = train_test_split_indices(len(inp), test_size, seed)
    Xtr = inp[tr].reshape(len(tr) * dim, -1)
    Xte = inp[te].reshape(len(te) * dim, -1)
    Ytr = output[tr].reshape(len(tr) * dim, -1)
    Yte = output[te].reshape(len(te) * dim, -1)
    fixed = []
    if kernel_variance is not None:
        kern.init["variance"] = kernel_variance          # KeyError-free only for RBF, like the AttributeError upstream
        fixed.append("variance")
    if kernel_lengthscale is not None:
        kern.init["lengthscales"] = kernel_lengthscale
        fixed.append("lengthscales")
    Z = None
    if n_inducing_points is not None:
        Z = Xtr[fps_features(Xtr, n_inducing_points)]
    gp = OracleModel(kern, Xtr, Ytr, Z=Z, kernel_lower=kernel_lower, fixed=fixed)
    gp.fit(epochs)
    pred, _ = gp.predict_f(Xte)
    gp.l2_error = float(np.linalg.norm(Yte - pred, axis=1).mean())
    return gp
