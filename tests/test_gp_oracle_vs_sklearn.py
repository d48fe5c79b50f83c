"""The GP half of the oracle is 'parity unpinned' against GPflow (not installable here).  This test pins its GPR equations
to an INDEPENDENT third-party implementation that is available: scikit-learn's GaussianProcessRegressor (Rasmussen & Williams
Alg. 2.1, the same equations GPflow's GPR evaluates): log marginal likelihood and predictive mean / variance, for the RBF
kernel (gpflow.kernels.RBF <-> ConstantKernel * RBF) and for the spectral kernel K = (X*S) X2^T (kernels.py:55-61 <-> DotProduct
on the features X * sqrt(S)).  scikit-learn's predictive variance with a WhiteKernel includes the noise; GPflow's predict_f does
not (main.py:111), hence the '+ noise'."""
import warnings

import numpy as np

from oracle import gp_oracle as GO


def _data(seed=0, N=80, k=6, R=3):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(N, k)) * 0.6, rng.normal(size=(N, R)), rng.normal(size=(17, k)) * 0.6, rng


def test_rbf_gpr_equations_match_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel
    X, Y, Xs, _ = _data()
    for var, ls, noise in ((1.3, 0.9, 0.2), (0.4, 2.0, 0.01), (1.0, 1.0, 1.0)):
        ker = ConstantKernel(var, "fixed") * RBF(ls, "fixed") + WhiteKernel(noise, "fixed")
        g = GaussianProcessRegressor(kernel=ker, optimizer=None, alpha=0.0).fit(X, Y)
        p = dict(variance=var, lengthscales=ls)
        lml = GO.gpr_general_lml(GO.RBFKernel(), p, X, Y, noise)
        assert abs(lml - g.log_marginal_likelihood_value_) <= 1e-10 * abs(lml)
        m, s = g.predict(Xs, return_std=True)
        rm, rv = GO.gpr_general_predict(GO.RBFKernel(), p, X, Y, noise, Xs)
        np.testing.assert_allclose(rm, m, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(rv + noise, s ** 2, rtol=1e-9, atol=1e-11)


def test_spectral_gpr_equations_match_sklearn():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import DotProduct, WhiteKernel
    X, Y, Xs, rng = _data(1)
    evals = np.sort(rng.uniform(0, 2, X.shape[1]))
    for nu, kappa, sf, noise, typ in ((1.5, 5.0, 1.0, 0.2, "matern"), (2.5, 2.0, 0.7, 0.05, "matern"), (0.0, 1.3, 1.2, 0.3, "se")):
        S = GO.eval_S(evals, nu, kappa, sf, 300.0, typ)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                  # log(sigma_0 = 0) in sklearn's theta bookkeeping
            ker = DotProduct(sigma_0=0.0, sigma_0_bounds="fixed") + WhiteKernel(noise, "fixed")
            g = GaussianProcessRegressor(kernel=ker, optimizer=None, alpha=0.0).fit(X * np.sqrt(S), Y[:, :1])
            m, s = g.predict(Xs * np.sqrt(S), return_std=True)
        lml = GO.gpr_lml_dense(X, Y[:, :1], S, noise)
        assert abs(lml - g.log_marginal_likelihood_value_) <= 1e-9 * abs(lml)
        G, b = X.T @ X, X.T @ Y[:, 0]
        assert abs(GO.gpr_lml_lowrank(G, b, float(Y[:, 0] @ Y[:, 0]), len(X), S, noise) - g.log_marginal_likelihood_value_) <= 1e-9 * abs(lml)
        rm, rv = GO.gpr_predict_dense(X, Y[:, :1], S, noise, Xs)
        np.testing.assert_allclose(rm[:, 0], m, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(rv[:, 0] + noise, s ** 2, rtol=1e-8, atol=1e-10)
