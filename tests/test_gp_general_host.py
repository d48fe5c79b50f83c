"""CPU tests of the HOST LOGIC of the SGPR / kernel='rbf' paths: rvgp_b200/gp_general.py,
kernels.py and main.py run against tests/fake_cabi.py (a NumPy emulation of the C-ABI calls they make -- test
infrastructure, see its header) and are compared with the NumPy restatement of GPflow in oracle/gp_oracle.py.
The same comparisons run through the real CUDA library in tests/test_gpu_gp_general.py.
Tolerances: objective 1e-10 rel, gradients 1e-7 rel, predictions 1e-8."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as GO
from tests import fake_cabi


def _problem(seed=0, N=70, k=8, R=3, Mu=11):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, k)) * 0.6
    Y = rng.normal(size=(N, R))
    Z = X[rng.choice(N, Mu, replace=False)] + 0.05 * rng.normal(size=(Mu, k))
    evals = np.sort(rng.uniform(0.0, 2.0, k))
    Xs = rng.normal(size=(23, k)) * 0.6
    return X, Y, Z, evals, Xs


class _FakeData:
    def __init__(self, evals, nrows):
        self.evals_Lc = evals
        self.evecs_Lc = np.zeros((int(nrows), len(evals)))


def _kernels(evals, nv):
    from rvgp_b200.kernels import ManifoldKernel, RBF
    d = _FakeData(evals, nv)
    return [
        (ManifoldKernel(d, nu=1.7, kappa=3.0, sigma_f=0.8, typ="matern"), GO.SpectralKernel(evals, nv, "matern"),
         dict(nu=1.7, kappa=3.0, sigma_f=0.8)),
        (ManifoldKernel(d, kappa=1.3, sigma_f=1.2, typ="se"), GO.SpectralKernel(evals, nv, "se"), dict(kappa=1.3, sigma_f=1.2)),
        (RBF(variance=1.3, lengthscales=0.9), GO.RBFKernel(), dict(variance=1.3, lengthscales=0.9)),
    ]


def _assert_grads(kg, rg, scale=None):
    for n in rg:
        tol = 1e-7 * max(abs(rg[n]), scale or 0.0, 1e-12)
        assert abs(kg[n] - rg[n]) <= tol, (n, kg[n], rg[n])


def test_dense_gpr_any_kernel_matches_oracle(monkeypatch):
    fake_cabi.install(monkeypatch)
    from rvgp_b200.gp_general import DenseGPR
    X, Y, Z, evals, Xs = _problem()
    for kern, okern, p in _kernels(evals, 210.0):
        gp = DenseGPR(torch.from_numpy(X), torch.from_numpy(Y), kern)
        for noise in (0.3, 0.02):
            lml, kg, dn = gp.lml_and_grads(noise)
            rl, rg, rdn = GO.gpr_general_lml(okern, p, X, Y, noise, grads=True)
            assert abs(lml - rl) <= 1e-10 * abs(rl)
            _assert_grads(kg, rg, scale=max(abs(v) for v in rg.values()) * 1e-3)
            assert abs(dn - rdn) <= 1e-7 * abs(rdn)
            assert abs(gp.lml_and_grads(noise, grads=False) - rl) <= 1e-10 * abs(rl)
            m, v = gp.predict(noise, torch.from_numpy(Xs), chunk=10)          # several chunks
            rm, rv = GO.gpr_general_predict(okern, p, X, Y, noise, Xs)
            np.testing.assert_allclose(m.numpy(), rm, rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(v.numpy(), rv, rtol=1e-8, atol=1e-10)


def test_sgpr_bound_gradients_and_predict_match_oracle(monkeypatch):
    fake_cabi.install(monkeypatch)
    from rvgp_b200.gp_general import DeviceSGPR
    X, Y, Z, evals, Xs = _problem(1)
    for kern, okern, p in _kernels(evals, 210.0):
        sg = DeviceSGPR(torch.from_numpy(X), torch.from_numpy(Y), torch.from_numpy(Z), kern)
        for noise in (0.3, 0.05):
            f, kg, dn, dZ = sg.elbo_and_grads(noise)
            rf, rg, rdn, rdZ = GO.sgpr_elbo(okern, p, X, Y, Z, noise, grads=True)
            assert abs(f - rf) <= 1e-10 * abs(rf)
            _assert_grads(kg, rg, scale=max(abs(v) for v in rg.values()) * 1e-3)
            assert abs(dn - rdn) <= 1e-7 * abs(rdn)
            np.testing.assert_allclose(dZ.numpy(), rdZ, rtol=1e-5, atol=1e-6 * np.abs(rdZ).max())   # Kuu is jitter-conditioned (1e-6)
            assert abs(sg.elbo_and_grads(noise, grads=False) - rf) <= 1e-10 * abs(rf)
            m, v = sg.predict(noise, torch.from_numpy(Xs), chunk=7)
            rm, rv = GO.sgpr_predict(okern, p, X, Y, Z, noise, Xs)
            np.testing.assert_allclose(m.numpy(), rm, rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(v.numpy(), rv, rtol=1e-8, atol=1e-10)


def test_oracle_gradients_vs_central_differences():
    """The oracle's hand-written adjoints (the only reference the GPU path has) against finite differences."""
    X, Y, Z, evals, _ = _problem(2, N=40, k=6, R=2, Mu=7)
    rng = np.random.default_rng(5)
    for okern in (GO.SpectralKernel(evals, 90.0), GO.SpectralKernel(evals, 90.0, "se", kappa=1.3), GO.RBFKernel(1.3, 0.9)):
        for Zz in (None, Z):
            m = GO.OracleModel(okern, X, Y, Z=Zz, noise=0.3)
            v = m.pack() + 0.1 * rng.normal(size=m.pack().shape)
            f, g = m.objective(v)
            num = np.zeros_like(v)
            for i in range(len(v)):
                e = np.zeros_like(v)
                e[i] = 1e-6
                num[i] = (m.objective(v + e, False) - m.objective(v - e, False)) / 2e-6
            assert np.abs(g - num).max() <= 1e-6 * np.abs(num).max()


def test_oracle_sgpr_bound_equals_closed_form_and_tightens():
    X, Y, Z, evals, _ = _problem(3)
    okern = GO.SpectralKernel(evals, 210.0)
    p = dict(nu=1.5, kappa=5.0, sigma_f=1.0)
    N, R = Y.shape
    Kuf = okern.K(p, Z, X)
    Kuu = okern.K(p, Z) + GO.DEFAULT_JITTER * np.eye(len(Z))
    Q = Kuf.T @ np.linalg.solve(Kuu, Kuf)
    Sig = Q + 0.3 * np.eye(N)
    ld = np.linalg.slogdet(Sig)[1]
    cf = sum(-0.5 * Y[:, r] @ np.linalg.solve(Sig, Y[:, r]) - 0.5 * ld - 0.5 * N * GO.LOG2PI for r in range(R))
    cf -= R / (2 * 0.3) * (okern.K_diag(p, X).sum() - np.trace(Q))
    assert abs(GO.sgpr_elbo(okern, p, X, Y, Z, 0.3) - cf) < 1e-8 * abs(cf)
    # the collapsed bound never exceeds the exact log marginal likelihood
    assert GO.sgpr_elbo(okern, p, X, Y, Z, 0.3) <= GO.gpr_general_lml(okern, p, X, Y, 0.3) + 1e-9


class _Dual:
    def __init__(self, a):
        self.a = torch.from_numpy(np.ascontiguousarray(a))

    def get_dev(self, device):
        return self.a


class _HostData:
    """The attributes train_gp touches, backed by CPU tensors (fake C-ABI)."""

    def __init__(self, n, D, k, seed=0):
        rng = np.random.default_rng(seed)
        self.n, self.device = n, None
        self.evals_Lc = np.sort(rng.uniform(0.01, 2.0, k))
        self.evecs_Lc = rng.normal(size=(n * D, k)) / np.sqrt(n)
        self.evals_L = np.sort(rng.uniform(0.0, 2.0, k))
        self.evecs_L = rng.normal(size=(n, k))
        w = rng.normal(size=(k,)) * np.exp(-np.arange(k) / 3.0)
        self.vectors = (self.evecs_Lc @ w).reshape(n, D) * np.sqrt(n) + 0.01 * rng.normal(size=(n, D))
        self._duals = {"evecs_Lc": _Dual(self.evecs_Lc), "evecs_L": _Dual(self.evecs_L), "vectors": _Dual(self.vectors)}
        self._duals["evecs_Lc"].dev = self._duals["evecs_Lc"].a
        self._duals["evecs_Lc"].host = self.evecs_Lc

    def device_array(self, name):
        return self._duals[name].a


@pytest.mark.parametrize("kernel,n_ind", [(None, None), ("rbf", None), (None, 12), ("rbf", 9)])
def test_train_gp_every_branch_matches_oracle_recipe(monkeypatch, kernel, n_ind, capsys):
    """main.py:11-84 through the drop-in ``train_gp`` for GPR / SGPR x spectral / rbf, few L-BFGS-B iterations:
    same split, same inducing points, same objective trajectory as the oracle running the reference's recipe."""
    fake_cabi.install(monkeypatch)
    from rvgp_b200 import main as M, params as P
    P.set_default_positive_minimum(0.0)
    d = _HostData(n=60, D=3, k=7)
    train_ind = np.arange(0, 60, 1)
    epochs = 6
    # the default branch's rank-k / fused K16 kernels are covered on the GPU; here it runs through the general dense path
    extra = dict(solver="general") if (kernel is None and n_ind is None) else {}
    gp = M.train_gp(d, train_ind=train_ind, kernel=kernel, n_inducing_points=n_ind, epochs=epochs, **extra)
    P.set_default_positive_minimum(0.0)
    og = GO.train_gp_general(d.evecs_Lc, d.evals_Lc, d.evecs_L, d.vectors, d.n, train_ind=train_ind,
                             n_inducing_points=n_ind, kernel=kernel, epochs=epochs)
    assert type(gp).__name__ == ("manifold_GPR" if n_ind is None else "manifold_SGPR")
    assert abs(gp.opt_result.fun - og.opt_result.fun) <= 1e-6 * abs(og.opt_result.fun)
    assert abs(gp.l2_error - og.l2_error) <= 1e-5 * max(og.l2_error, 1e-3)
    out = capsys.readouterr().out
    assert "Relative l2 error is" in out
    assert ("Using RBF kernel" in out) == (kernel == "rbf")
    # predictions at the optimum, through transform() with positional encodings (main.py:108-109)
    feats = (d.evecs_L if kernel == "rbf" else d.evecs_Lc.reshape(d.n, -1)[:5].reshape(-1, 7))[:15]
    mean, var = gp.transform(d, feats)
    rm, rv = og.predict_f(feats)
    np.testing.assert_allclose(mean, rm.reshape(len(feats), -1), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(var, rv.reshape(len(feats), -1), rtol=1e-4, atol=1e-7)
    P.set_default_positive_minimum(0.0)


def test_kernel_variance_and_lengthscale_are_fixed_for_rbf_and_raise_for_manifold(monkeypatch):
    fake_cabi.install(monkeypatch)
    from rvgp_b200 import main as M, params as P
    P.set_default_positive_minimum(0.0)
    d = _HostData(n=40, D=3, k=5, seed=3)
    gp = M.train_gp(d, kernel="rbf", kernel_variance=2.0, kernel_lengthscale=1.5, epochs=3)
    assert abs(gp.kernel.variance.value - 2.0) < 1e-12 and abs(gp.kernel.lengthscales.value - 1.5) < 1e-12
    assert [p.name for p in gp.trainable_parameters] == ["variance"]          # only the likelihood variance is left
    with pytest.raises(AttributeError):                                         # ManifoldKernel has no .variance (main.py:70)
        M.train_gp(d, kernel_variance=2.0, epochs=1, solver="general")
    P.set_default_positive_minimum(0.0)
