"""GPU parity tests for the GP half (K13 Gram, K14 Cholesky/TRSM, K15 rank-k path, fit/transform) against the
NumPy restatement of GPflow 2.6.5 (oracle/gp_oracle.py; parity unpinned -- GPflow is not installable here).
Tolerances: LML 1e-10 rel, gradients 1e-7, predictive mean / variance 1e-6 rel (north_star)."""
import numpy as np
import pytest
import torch

from oracle import gp_oracle as GO
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(_dev())


@pytest.mark.parametrize("n", [1, 5, 64, 65, 200, 777])
def test_potrf_trsm(n):
    from rvgp_b200.gp import _Chol
    from rvgp_b200._cabi import get_handle
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n + 3))
    A = A @ A.T + 0.5 * np.eye(n)
    Ad = _t(A)
    ch = _Chol(get_handle(0), Ad, n)
    ch.check()
    L = np.tril(Ad.cpu().numpy())
    Lref = np.linalg.cholesky(A)
    np.testing.assert_allclose(L, Lref, rtol=1e-11, atol=1e-12)
    B = rng.normal(size=(n, 7))
    for trans in (0, 1):
        Bd = _t(B)
        ch.solve(Bd, trans)
        ref = np.linalg.solve(Lref.T if trans else Lref, B)
        np.testing.assert_allclose(Bd.cpu().numpy(), ref, rtol=1e-9, atol=1e-11)
    assert abs(float(ch.logdiag_sum().item()) - np.log(np.diag(Lref)).sum()) < 1e-10 * max(1, n)


def test_potrf_not_spd_raises():
    from rvgp_b200.gp import _Chol
    from rvgp_b200._cabi import get_handle, RvgpError
    A = np.eye(100); A[70, 70] = -1.0
    ch = _Chol(get_handle(0), _t(A), 100)
    with pytest.raises(RvgpError):
        ch.check()


def test_kernel_class_K_and_Kdiag_dlpack():
    from rvgp_b200.kernels import ManifoldKernel
    g = load_golden("torus_n600_k20")

    class D:
        evals_Lc, evecs_Lc = g["evals_Lc"], g["evecs_Lc"]
    for typ in ("matern", "se"):
        kern = ManifoldKernel(D, nu=1.5, kappa=5.0, sigma_f=1.3, typ=typ)
        S = GO.eval_S(g["evals_Lc"], 1.5, 5.0, 1.3, g["evecs_Lc"].shape[0], typ)
        np.testing.assert_allclose(kern.eval_S(typ), S, rtol=1e-14)
        X, X2 = g["evecs_Lc"][:301], g["evecs_Lc"][500:777]
        Kd = kern.K(X, X2)                                     # numpy in
        assert Kd.is_cuda and hasattr(Kd, "__dlpack__")
        np.testing.assert_allclose(Kd.cpu().numpy(), GO.K(X, S, X2), rtol=1e-12, atol=1e-12)
        Kd2 = kern.K(torch.utils.dlpack.from_dlpack(_t(X).__dlpack__()))      # CUDA DLPack in
        np.testing.assert_allclose(Kd2.cpu().numpy(), GO.K(X, S), rtol=1e-12, atol=1e-12)
        kd = kern.K_diag(_t(X))
        np.testing.assert_allclose(kd.cpu().numpy(), GO.K_diag(X, S), rtol=1e-13)
        np.testing.assert_allclose(kd.cpu().numpy(), np.diag(Kd2.cpu().numpy()), rtol=1e-12)


@pytest.mark.parametrize("solver", ["dense", "lowrank"])
def test_lml_gradients_and_predict_match_oracle(solver):
    from rvgp_b200.gp import DeviceGPR
    g = load_golden("sphere_n2000_k50")
    n = 2000
    np.random.seed(0)
    train_ind = np.random.choice(np.arange(n), size=n // 2)
    Xtr, Ytr, Xte, Yte = GO.prepare_training(g["evecs_Lc"], g["smoothed_field"], n, train_ind)
    gp = DeviceGPR(_t(Xtr), _t(Ytr), solver=solver)
    nv = g["evecs_Lc"].shape[0]
    for (nu, kappa, sf, noise) in [(1.5, 5.0, 1.0, 1.0), (2.3, 3.1, 0.7, 0.05)]:
        S = GO.eval_S(g["evals_Lc"], nu, kappa, sf, nv)
        lml, dS, dn = gp.lml_and_grads(S, noise)
        rl, rdS, rdn = GO.gpr_lml_dense(Xtr, Ytr, S, noise, grads=True)
        assert abs(lml - rl) <= 1e-10 * abs(rl)
        np.testing.assert_allclose(dS, rdS, rtol=1e-7, atol=1e-9 * np.abs(rdS).max())
        assert abs(dn - rdn) <= 1e-7 * abs(rdn)
        m, v = gp.predict(S, noise, _t(Xte))
        rm, rv = GO.gpr_predict_dense(Xtr, Ytr, S, noise, Xte)
        np.testing.assert_allclose(m.cpu().numpy(), rm, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(v.cpu().numpy(), rv[:, :1], rtol=1e-6, atol=1e-9)
    if solver == "dense":
        assert abs(gp.lml_and_grads(GO.eval_S(g["evals_Lc"], 1.5, 5.0, 1.0, nv), 1.0, grads=False) - (-2512.1283414532)) < 1e-6


def test_readme_quickstart_end_to_end():
    """README.md:83-107 on config C1 through the drop-in names; compared with the oracle running the same recipe."""
    import RVGP
    from RVGP.geometry import furthest_point_sampling
    from rvgp_b200 import params as P
    from tests.workloads import make_cloud
    X = make_cloud("sphere", 2000, 0)
    sample_ind, _ = furthest_point_sampling(X, stop_crit=0.0)
    X = X[sample_ind]
    g = load_golden("sphere_n2000_k50")
    P.set_default_positive_minimum(0.0)                       # fresh-process state of gpflow.config
    d = RVGP.create_data_object(X, n_eigenpairs=50)
    d.random_vector_field(seed=1)
    d.smooth_vector_field(t=100)
    np.testing.assert_allclose(d.vectors, g["smoothed_field"], atol=1e-9)
    np.random.seed(0)
    train_ind = np.random.choice(np.arange(len(X)), size=int(0.5 * len(X)))
    gp = RVGP.fit(d, train_ind=train_ind, noise_variance=0.001)
    test_ind = [i for i in range(len(X)) if i not in train_ind]
    mean, var = gp.transform(d, test_ind)
    assert mean.shape == (len(test_ind), 3) and var.shape == (len(test_ind), 3)
    # oracle: same data, same recipe, on the REFERENCE's eigenbasis (golden)
    og = GO.train_gp(g["evecs_Lc"], g["evals_Lc"], g["smoothed_field"], 2000, train_ind, epochs=1000, kernel_lower=0.0,
                     solver="lowrank")
    om, ov = GO.transform(og, g["evecs_Lc"], 2000, test_ind)
    p, q = gp.kernel, og.params()
    assert abs(gp.opt_result.fun - og.opt_result.fun) < 1e-5 * abs(og.opt_result.fun)
    np.testing.assert_allclose(mean, om, rtol=1e-4, atol=1e-5)          # after ~100s of L-BFGS-B steps
    np.testing.assert_allclose(var, ov, rtol=1e-3, atol=1e-7)
    err = np.linalg.norm(mean - d.vectors[test_ind], axis=1).mean()
    assert err < 0.1
    # at FIXED hyper-parameters (the oracle's optimum) predictions agree to 1e-6 (north_star tolerance)
    gp.kernel.nu.assign(q["nu"]); gp.kernel.kappa.assign(q["kappa"]); gp.kernel.sigma_f.assign(q["sigma_f"])
    gp.likelihood.variance.assign(q["noise"])
    mean2, var2 = gp.transform(d, np.asarray(test_ind))                  # numpy int array accepted (App. B.1)
    np.testing.assert_allclose(mean2, om, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(var2, ov, rtol=1e-6, atol=1e-9)
    mask = np.zeros(len(X), dtype=bool); mask[test_ind] = True
    mean3, _ = gp.transform(d, mask)                                     # boolean mask accepted
    assert np.array_equal(mean3, mean2)
    # second fit of the process: kernel lower bound is now 1e-2 (main.py:30 vs :52)
    gp2 = RVGP.fit(d, train_ind=train_ind, epochs=5)
    assert gp2.kernel.kappa.transform.lower == 1e-2 and gp.kernel.kappa.transform.lower == 0.0


@pytest.mark.parametrize("k", [3, 10, 20, 50, 64])
def test_small_k_fused_kernel_matches_oracle(k):
    """K16: fused small-k evaluation == NumPy rank-k restatement == dense GPflow-style restatement."""
    from rvgp_b200.gp import DeviceGPR
    g = load_golden("sphere_n2000_k50" if k <= 50 else "sphere_n2000_k50")
    rng = np.random.default_rng(k)
    kk = min(k, 50)
    Phi = g["evecs_Lc"][rng.choice(6000, 153, replace=False)][:, :kk]
    if k > kk:
        Phi = np.concatenate([Phi, rng.normal(size=(153, k - kk))], 1)
    lam = np.sort(np.concatenate([g["evals_Lc"][:kk], 0.7 + rng.uniform(size=k - kk)]))
    Y = rng.normal(size=(153, 1))
    gp = DeviceGPR(_t(Phi), _t(Y), solver="lowrank")
    assert gp._small
    for (nu, kappa, sf, noise) in [(1.5, 5.0, 1.0, 1.0), (2.0, 2.0, 0.5, 0.01)]:
        S = GO.eval_S(lam, nu, kappa, sf, 6000)
        lml, dS, dn = gp.lml_and_grads(S, noise)
        rl, rdS, rdn = GO.gpr_lml_dense(Phi, Y, S, noise, grads=True)
        assert abs(lml - rl) <= 1e-9 * abs(rl)
        np.testing.assert_allclose(dS, rdS, rtol=1e-6, atol=1e-8 * np.abs(rdS).max())
        assert abs(dn - rdn) <= 1e-6 * abs(rdn)
        Xn = rng.normal(size=(37, k))
        m, v = gp.predict(S, noise, _t(Xn))
        rm, rv = GO.gpr_predict_dense(Phi, Y, S, noise, Xn)
        np.testing.assert_allclose(m.cpu().numpy(), rm, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(v.cpu().numpy(), rv[:, :1], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("k", [65, 100, 129, 200, 500])
def test_rank_k_evaluation_kernels_match_oracle(k):
    """K15b (64 < k): build + blocked Cholesky + the fused forward-solve / back-solve kernels, captured and replayed as a CUDA
    graph, against the dense GPflow-style restatement.  k not a multiple of the 64-row panel, repeated evaluations with
    changing parameters (graph replay), a rank-deficient Gram part (M < k for k = 500) and the not-SPD flag."""
    from rvgp_b200.gp import DeviceGPR
    from rvgp_b200._cabi import RvgpError
    g = load_golden("sphere_n2000_k50")
    rng = np.random.default_rng(k)
    M = 420
    Phi = np.concatenate([g["evecs_Lc"][rng.choice(6000, M, replace=False)][:, :50], rng.normal(size=(M, k - 50)) / np.sqrt(6000)], 1)
    lam = np.sort(np.concatenate([g["evals_Lc"][:50], 0.7 + 3 * rng.uniform(size=k - 50)]))
    Y = rng.normal(size=(M, 1))
    gp = DeviceGPR(_t(Phi), _t(Y), solver="lowrank")
    assert not getattr(gp, "_small", False)
    for rep, (nu, kappa, sf, noise) in enumerate([(1.5, 5.0, 1.0, 1.0), (2.0, 2.0, 0.5, 0.01), (1.5, 5.0, 1.0, 1.0), (2.5, 9.0, 2.0, 0.3)]):
        S = GO.eval_S(lam, nu, kappa, sf, 6000)
        lml, dS, dn = gp.lml_and_grads(S, noise)
        rl, rdS, rdn = GO.gpr_lml_dense(Phi, Y, S, noise, grads=True)
        assert abs(lml - rl) <= 1e-9 * abs(rl), (rep, lml, rl)
        np.testing.assert_allclose(dS, rdS, rtol=1e-6, atol=1e-8 * np.abs(rdS).max())
        assert abs(dn - rdn) <= 1e-6 * abs(rdn)
    Xn = rng.normal(size=(37, k)) / np.sqrt(6000)
    m, v = gp.predict(S, noise, _t(Xn))
    rm, rv = GO.gpr_predict_dense(Phi, Y, S, noise, Xn)
    np.testing.assert_allclose(m.cpu().numpy(), rm, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(v.cpu().numpy(), rv[:, :1], rtol=1e-6, atol=1e-9)
    Sbad = S.copy(); Sbad[k // 2] = -1e12                       # B loses positive definiteness -> flag -> RvgpError
    with pytest.raises((RvgpError, FloatingPointError, ValueError)):
        with np.errstate(all="raise"):
            gp.lml_and_grads(Sbad, noise)


def test_eeg_shaped_many_fits_on_fixed_eigenbasis():
    """Config C3 shape: one data object, one GP fit per frame on a fixed eigenbasis (eeg_utils.py:26-35)."""
    import time
    import RVGP
    from tests.workloads import make_cloud
    from rvgp_b200 import params as P
    X = make_cloud("scalp", 3000, 0)
    d = RVGP.create_data_object(X, n_eigenpairs=10, verbose=False)
    chans = RVGP.geometry.furthest_point_sampling(X, N=64)[0]
    rng = np.random.default_rng(1)
    Phi = d.evecs_Lc.reshape(d.n, 3, 10)
    rest = np.setdiff1d(np.arange(d.n), chans)[:200]
    t0 = time.perf_counter()
    errs = []
    for frame in range(12):
        coef = rng.normal(size=10)
        field = Phi @ coef                                  # (n, 3) smooth field = combination of eigenvectors
        d.vectors = field
        gp = RVGP.fit(d, train_ind=chans, epochs=100, noise_variance=0.001)
        pred, var = gp.transform(d, rest.reshape(-1, 1))    # (n,1) int array as in eeg_utils.py:112
        assert gp._gpr._small and pred.shape == (200, 3)
        errs.append(np.linalg.norm(pred - field[rest]) / np.linalg.norm(field[rest]))
    dt = (time.perf_counter() - t0) / 12
    assert np.median(errs) < 0.05, errs
    print("per-frame fit+transform: %.1f ms, median rel err %.2e" % (dt * 1e3, np.median(errs)))


def test_transform_with_float_positional_encodings():
    """main.py:108-109 / README.md:110: float test_ind = positional encodings (rows of the spectral embedding)."""
    import RVGP
    g = load_golden("torus_n600_k20")
    d = RVGP.create_data_object(g["X"], n_eigenpairs=20, verbose=False)
    d.vectors = g["smoothed_field"]
    gp = RVGP.fit(d, train_ind=np.arange(0, 600, 2), epochs=50)
    nodes = np.arange(1, 600, 2)[:40]
    m_int, v_int = gp.transform(d, [int(i) for i in nodes])
    enc = d.evecs_Lc.reshape(d.n, -1)[nodes].reshape(-1, 20)           # (40*3, k) float rows
    m_f, v_f = gp.transform(d, enc)
    assert m_f.shape == (120, 1)
    np.testing.assert_allclose(m_f.reshape(40, 3), m_int, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(v_f.reshape(40, 3), v_int, rtol=1e-12, atol=1e-14)


def test_interpolate_timepoint_subgraph_diffusion_matches_dense_expm():
    """eeg_utils.interpolate_timepoint(project=True, t>0) (examples/eeg_example/eeg_utils.py:96-106): the training-node signal is
    diffused on the SUB-GRAPH (principal sub-matrices of Lc and L) before the fit.  The device path (sparse slices + Chebyshev
    exp(-tA) action) is checked against the reference's dense scipy.linalg.expm recipe via the oracle."""
    import contextlib
    import io
    import scipy.sparse as sp
    import RVGP
    from rvgp_b200 import eeg_utils
    from rvgp_b200 import params as P
    from oracle import rvgp_oracle as O
    from tests.workloads import make_cloud
    X = make_cloud("sphere", 600, 3)
    d = RVGP.create_data_object(X, n_eigenpairs=20, verbose=False)
    d.random_vector_field(seed=2)
    field = d.vectors.copy()
    rng = np.random.RandomState(1)
    train_idx = np.sort(rng.choice(600, 200, replace=False))
    test_idx = np.setdiff1d(np.arange(600), train_idx)
    t = 0.3          # the sub-graph of a third of the nodes is strongly diagonally dominant: exp(-t A) decays like exp(-7 t)
    # reference recipe on the host (dense expm, restated in the oracle)
    g = d.gauges[train_idx]
    v = O.express_in_local_frame(field[train_idx], g)
    Lc_idx = np.sort(np.hstack([train_idx * 2, train_idx * 2 + 1]))
    Lc_ = sp.bsr_matrix(d.Lc.tocsr()[Lc_idx][:, Lc_idx], blocksize=(2, 2))
    L_ = d.L[train_idx][:, train_idx]
    v = O.vector_diffusion(v, t, L_, Lc_, dense=True)
    want = field.copy()
    want[train_idx] = O.express_in_local_frame(np.asarray(v), g, reverse=True)
    P.set_default_positive_minimum(0.0)
    with contextlib.redirect_stdout(io.StringIO()):
        pred = eeg_utils.interpolate_timepoint(d, train_idx, test_idx, project=True, t=t)
    np.testing.assert_allclose(d.vectors, want, atol=1e-8)          # the smoothed field that was fitted
    assert pred.shape == (len(test_idx), 3) and np.all(np.isfinite(pred))


def test_affinity_graph_matches_reference_recipe():
    """manifold_graph(typ='affinity') (geometry.py:114-118): dense Gaussian-kernel weights on the device == the reference's
    sklearn / numpy recipe, and its Laplacian == networkx's."""
    import networkx as nx
    import scipy.sparse as sp
    from sklearn.metrics import pairwise_distances
    from rvgp_b200 import geometry as geo
    from tests.workloads import make_cloud
    X = make_cloud("sphere", 300, 5)
    G = geo.manifold_graph(X, typ="affinity")
    A_ref = np.exp(-pairwise_distances(X) ** 2 / (2 * 0.1 ** 2))
    A = G.weights.reshape(300, 300).cpu().numpy()
    assert np.abs(A - A_ref).max() < 1e-12 and np.all(np.diag(A) == 1.0)
    Gn = G.to_networkx()
    Gr = nx.from_numpy_array(A_ref)
    assert Gn.number_of_edges() == Gr.number_of_edges() and np.allclose(Gn.nodes[7]["pos"], X[7])
    L = geo.compute_laplacian(G)
    Lr = sp.csr_matrix(nx.laplacian_matrix(Gr), dtype=np.float64)
    assert abs(L - Lr).max() < 1e-11
    with pytest.raises(NotImplementedError):
        import ptu_dijkstra
        ptu_dijkstra.tangent_frames(X, Gn, 2, 15)                     # weighted graph: refused, not silently unit-weighted
    with pytest.raises(ValueError):
        geo.manifold_graph(X, typ="nope")
