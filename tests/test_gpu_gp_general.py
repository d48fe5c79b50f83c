"""GPU parity tests for SURVEY.md 8f rows 1 and 3: the SGPR path (main.py:59-67,119-137) and kernel='rbf'
(main.py:33-37) through the CUDA library, against the NumPy restatement of GPflow 2.6.5 in oracle/gp_oracle.py
(parity unpinned -- GPflow is not installable here; the oracle's adjoints are checked against finite differences in
tests/test_gp_general_host.py).  Tolerances: objective 1e-10 rel, gradients 1e-7 rel (dZ 1e-5: Kuu is conditioned by the
1e-6 jitter), predictive mean / variance 1e-6 rel (north_star)."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as GO
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.device("cuda", 0))


def _c1():
    g = load_golden("sphere_n2000_k50")
    n = 2000
    np.random.seed(0)
    train_ind = np.random.choice(np.arange(n), size=n // 2)
    return g, n, train_ind


class _D:
    def __init__(self, g):
        self.evals_Lc, self.evecs_Lc = g["evals_Lc"], g["evecs_Lc"]


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (37, 2500, 9), (300, 257, 50), (70, 4097, 3)])
def test_rbf_gram_and_adjoint_kernels(m, n, k):
    """K17 kernels against the oracle's RBFKernel: Gram, parameter adjoints, dF/dXA; several column chunks / ragged tails."""
    from rvgp_b200.kernels import RBF
    rng = np.random.default_rng(m * 7 + n)
    XA, XB = rng.normal(size=(m, k)) * 0.7, rng.normal(size=(n, k)) * 0.7
    Gbar = rng.normal(size=(m, n))
    kern, okern, p = RBF(variance=1.7, lengthscales=1.3), GO.RBFKernel(), dict(variance=1.7, lengthscales=1.3)
    Kd = kern.K(XA, XB)
    np.testing.assert_allclose(Kd.cpu().numpy(), okern.K(p, XA, XB), rtol=1e-12, atol=1e-14)
    Ks = kern.K(XA)
    np.testing.assert_allclose(Ks.cpu().numpy(), okern.K(p, XA), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(kern.K_diag(XA).cpu().numpy(), okern.K_diag(p, XA), rtol=0, atol=0)
    rg, rdX = okern.adjoint(p, XA, XB, Gbar)
    for want in (False, True):
        g, dX = kern._adjoint(_t(XA), _t(XB), _t(Gbar), want_dX=want)
        scale = np.abs(Gbar * okern.K(p, XA, XB)).sum()
        assert abs(g["variance"] - rg["variance"]) <= 1e-11 * scale
        assert abs(g["lengthscales"] - rg["lengthscales"]) <= 1e-10 * scale * 10
        if want:
            np.testing.assert_allclose(dX.cpu().numpy(), rdX, rtol=1e-9, atol=1e-11 * scale)


@pytest.mark.parametrize("D,N", [(65, 20), (200, 33), (500, 64)])
def test_fps_in_feature_space_matches_reference_recipe(D, N):
    """furthest_point_sampling on (M, k) feature rows (main.py:60): the warp-per-point path for D > 64."""
    from rvgp_b200.geometry import furthest_point_sampling
    rng = np.random.default_rng(D)
    x = rng.normal(size=(700, D)) / np.sqrt(D)
    perm, lam = furthest_point_sampling(x, N=N)
    ref = GO.fps_features(x, N)
    assert np.array_equal(perm, ref)
    from sklearn.metrics import pairwise_distances
    Dm = pairwise_distances(x)
    ds = Dm[0].copy()
    for i in range(1, N):
        assert abs(lam[i] - ds[ref[i]]) <= 1e-12 * ds[ref[i]]
        ds = np.minimum(ds, Dm[ref[i]])


def _kernel_pairs(g):
    from rvgp_b200.kernels import ManifoldKernel, RBF
    nv = g["evecs_Lc"].shape[0]
    return [
        (lambda: ManifoldKernel(_D(g), nu=1.5, kappa=5.0, sigma_f=1.0), GO.SpectralKernel(g["evals_Lc"], nv),
         dict(nu=1.5, kappa=5.0, sigma_f=1.0)),
        (lambda: RBF(variance=1.2, lengthscales=0.8), GO.RBFKernel(), dict(variance=1.2, lengthscales=0.8)),
    ]


def test_dense_gpr_multi_output_matches_oracle():
    """gpflow GPR with R = 3 output columns sharing one kernel (the channel-wise layout of kernel='rbf')."""
    from rvgp_b200.gp_general import DenseGPR
    g, n, train_ind = _c1()
    rows = train_ind[:700]
    for feats in ("evecs_Lc", "evecs_L"):
        X = (g["evecs_Lc"].reshape(n, -1)[rows][:, :50] if feats == "evecs_Lc" else g["evecs_L"][rows])
        Y = g["smoothed_field"][rows]
        for mk, okern, p in _kernel_pairs(g):
            gp = DenseGPR(_t(X), _t(Y), mk())
            for noise in (1.0, 0.05):
                lml, kg, dn = gp.lml_and_grads(noise)
                rl, rg, rdn = GO.gpr_general_lml(okern, p, X, Y, noise, grads=True)
                assert abs(lml - rl) <= 1e-10 * abs(rl)
                gs = max(abs(v) for v in rg.values())
                for nm in rg:
                    assert abs(kg[nm] - rg[nm]) <= 1e-7 * max(abs(rg[nm]), 1e-3 * gs), (nm, kg[nm], rg[nm])
                assert abs(dn - rdn) <= 1e-7 * abs(rdn)
                Xs = X[:123] + 0.01
                m, v = gp.predict(noise, _t(Xs), chunk=50)
                rm, rv = GO.gpr_general_predict(okern, p, X, Y, noise, Xs)
                np.testing.assert_allclose(m.cpu().numpy(), rm, rtol=1e-6, atol=1e-8)
                np.testing.assert_allclose(v.cpu().numpy(), rv, rtol=1e-6, atol=1e-9)


def test_general_dense_path_agrees_with_spectral_fast_path():
    """Same kernel, same data: gp_general.DenseGPR == gp.DeviceGPR (dense and rank-k) to 1e-10."""
    from rvgp_b200.gp import DeviceGPR
    from rvgp_b200.gp_general import DenseGPR
    from rvgp_b200.kernels import ManifoldKernel
    g, n, train_ind = _c1()
    Xtr, Ytr, Xte, Yte = GO.prepare_training(g["evecs_Lc"], g["smoothed_field"], n, train_ind[:300])
    kern = ManifoldKernel(_D(g), nu=1.5, kappa=5.0, sigma_f=1.0)
    S = kern.eval_S("matern")
    gen = DenseGPR(_t(Xtr), _t(Ytr), kern)
    l0, kg, dn = gen.lml_and_grads(0.3)
    for solver in ("dense", "lowrank"):
        l1, dS, dn1 = DeviceGPR(_t(Xtr), _t(Ytr), solver=solver).lml_and_grads(S, 0.3)
        assert abs(l0 - l1) <= 1e-10 * abs(l1)
        ch = kern._chain(dS)
        for nm in ch:
            assert abs(kg[nm] - ch[nm]) <= 1e-7 * max(abs(ch[nm]), 1e-6)
        assert abs(dn - dn1) <= 1e-7 * abs(dn1)


@pytest.mark.parametrize("Mu", [16, 40])
def test_sgpr_bound_gradients_predict_match_oracle(Mu):
    from rvgp_b200.gp_general import DeviceSGPR
    g, n, train_ind = _c1()
    rows = train_ind[:400]
    for feats in ("evecs_Lc", "evecs_L"):
        if feats == "evecs_Lc":
            X = g["evecs_Lc"].reshape(n, -1)[rows].reshape(-1, 50)
            Y = g["smoothed_field"][rows].reshape(-1, 1)
        else:
            X, Y = g["evecs_L"][rows], g["smoothed_field"][rows]
        Z = X[GO.fps_features(X, Mu)].copy()
        for mk, okern, p in _kernel_pairs(g):
            sg = DeviceSGPR(_t(X), _t(Y), _t(Z), mk())
            for noise in (1.0, 0.05):
                f, kg, dn, dZ = sg.elbo_and_grads(noise)
                rf, rg, rdn, rdZ = GO.sgpr_elbo(okern, p, X, Y, Z, noise, grads=True)
                assert abs(f - rf) <= 1e-9 * abs(rf)
                gs = max(abs(v) for v in rg.values())
                for nm in rg:
                    assert abs(kg[nm] - rg[nm]) <= 1e-6 * max(abs(rg[nm]), 1e-3 * gs), (nm, kg[nm], rg[nm])
                assert abs(dn - rdn) <= 1e-6 * abs(rdn)
                np.testing.assert_allclose(dZ.cpu().numpy(), rdZ, rtol=1e-4, atol=1e-5 * np.abs(rdZ).max())
                Xs = X[:97] * 1.01
                m, v = sg.predict(noise, _t(Xs), chunk=40)
                rm, rv = GO.sgpr_predict(okern, p, X, Y, Z, noise, Xs)
                np.testing.assert_allclose(m.cpu().numpy(), rm, rtol=1e-6, atol=1e-8)
                np.testing.assert_allclose(v.cpu().numpy(), rv, rtol=1e-6, atol=1e-8)


def _fresh_data_object():
    import RVGP
    from rvgp_b200 import params as P
    from tests.workloads import make_cloud
    P.set_default_positive_minimum(0.0)
    X = make_cloud("sphere", 2000, 0)
    d = RVGP.create_data_object(X, n_eigenpairs=50)
    d.random_vector_field(seed=1)
    d.smooth_vector_field(t=100)
    return RVGP, d


@pytest.mark.parametrize("kernel,n_ind", [("rbf", None), (None, 30), ("rbf", 25)])
def test_fit_rbf_and_sgpr_end_to_end(kernel, n_ind):
    """RVGP.fit(kernel='rbf') / RVGP.fit(n_inducing_points=...) through the drop-in names on config C1, against the
    oracle running the reference's recipe (main.py:11-84) on the SAME eigenbasis; 15 L-BFGS-B iterations."""
    from rvgp_b200 import params as P
    RVGP, d = _fresh_data_object()
    np.random.seed(0)
    train_ind = np.random.choice(np.arange(d.n), size=600)
    epochs = 15
    gp = RVGP.fit(d, train_ind=train_ind, kernel=kernel, n_inducing_points=n_ind, epochs=epochs)
    assert type(gp).__name__ == ("manifold_SGPR" if n_ind else "manifold_GPR")
    P.set_default_positive_minimum(0.0)
    og = GO.train_gp_general(np.asarray(d.evecs_Lc), np.asarray(d.evals_Lc), np.asarray(d.evecs_L), np.asarray(d.vectors),
                             d.n, train_ind=train_ind, n_inducing_points=n_ind, kernel=kernel, epochs=epochs)
    P.set_default_positive_minimum(0.0)
    # same optimiser, same objective and gradients -> the same trajectory up to round-off amplification
    assert abs(gp.opt_result.fun - og.opt_result.fun) <= 1e-3 * abs(og.opt_result.fun)
    assert abs(gp.l2_error - og.l2_error) <= 2e-2 * max(og.l2_error, 1e-2)
    # predictions of BOTH models at the GPU model's optimum: copy the GPU hyper-parameters into the oracle
    og.u = np.array([GO.softplus_inv(dict({p.name: p.value for p in gp.kernel.trainable_parameters},
                                          noise=gp.likelihood.variance.value)[nm] - og._lower(nm)) for nm in og.names])
    if n_ind:
        og.Z = gp.inducing_variable.numpy()
        assert og.Z.shape == (n_ind, 50)
    feats = (np.asarray(d.evecs_L)[:40] if kernel == "rbf" else np.asarray(d.evecs_Lc).reshape(d.n, -1)[:14].reshape(-1, 50))
    mean, var = gp.transform(d, feats)
    rm, rv = og.predict_f(feats)
    np.testing.assert_allclose(mean, rm.reshape(len(feats), -1), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(var, rv.reshape(len(feats), -1), rtol=1e-6, atol=1e-8)
    if kernel is None:
        # integer node indices go through evecs_Lc like the reference (main.py:104-106)
        m2, v2 = gp.transform(d, [0, 5, 9])
        assert m2.shape == (3, 3) and v2.shape == (3, 3)
