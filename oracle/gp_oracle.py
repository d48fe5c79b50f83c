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
