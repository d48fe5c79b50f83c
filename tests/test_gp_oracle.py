"""CPU checks of the GP restatement (oracle/gp_oracle.py).  GPflow/TensorFlow are absent, so this half is
'parity unpinned'; these are the cross-checks that stand in for pins."""
import numpy as np
import pytest

from oracle import gp_oracle as GO
from tests.conftest import load_golden


def _c1_setup():
    g = load_golden("sphere_n2000_k50")
    n = 2000
    # README quick-start: field = the golden smoothed random field, train on a with-replacement sample
    np.random.seed(0)
    train_ind = np.random.choice(np.arange(n), size=n // 2)
    return g, n, train_ind


def test_initial_lml_matches_survey_probe():
    g, n, train_ind = _c1_setup()
    assert len(np.unique(train_ind)) == 790
    Xtr, Ytr, Xte, Yte = GO.prepare_training(g["evecs_Lc"], g["smoothed_field"], n, train_ind)
    assert Xtr.shape == (2400, 50) and Xte.shape == (600, 50)
    S = GO.eval_S(g["evals_Lc"], 1.5, 5.0, 1.0, g["evecs_Lc"].shape[0])
    lml = GO.gpr_lml_dense(Xtr, Ytr, S, 1.0)
    assert abs(lml - (-2512.1283414532)) < 1e-6          # SURVEY.md section 8c probe value
    G, b, yy = Xtr.T @ Xtr, Xtr.T @ Ytr[:, 0], float(Ytr[:, 0] @ Ytr[:, 0])
    assert abs(GO.gpr_lml_lowrank(G, b, yy, 2400, S, 1.0) - lml) < 1e-8


def test_gradients_vs_finite_differences_and_solver_agreement():
    g = load_golden("torus_n600_k20")
    n = 600
    rng = np.random.default_rng(0)
    train_ind = rng.choice(n, 200, replace=False)
    Xtr, Ytr, _, _ = GO.prepare_training(g["evecs_Lc"], g["smoothed_field"], n, train_ind)
    for solver in ("dense", "lowrank"):
        gp = GO.OracleGPR(Xtr, Ytr, g["evals_Lc"], g["evecs_Lc"].shape[0], solver=solver)
        u = gp.u + rng.normal(scale=0.3, size=gp.u.shape)
        f, gr = gp.loss_and_grad(u)
        for i in range(len(u)):
            e = np.zeros_like(u); e[i] = 1e-6
            fd = (gp.loss_and_grad(u + e)[0] - gp.loss_and_grad(u - e)[0]) / 2e-6
            assert abs(fd - gr[i]) <= 1e-5 * max(1.0, abs(gr[i])), (solver, i, fd, gr[i])
    a = GO.OracleGPR(Xtr, Ytr, g["evals_Lc"], g["evecs_Lc"].shape[0], solver="dense")
    b = GO.OracleGPR(Xtr, Ytr, g["evals_Lc"], g["evecs_Lc"].shape[0], solver="lowrank")
    fa, ga = a.loss_and_grad(a.u)
    fb, gb = b.loss_and_grad(b.u)
    assert abs(fa - fb) < 1e-8 * abs(fa) and np.allclose(ga, gb, rtol=1e-7, atol=1e-9)


def test_kdiag_and_predict_agreement():
    g = load_golden("torus_n600_k20")
    X = g["evecs_Lc"][:300]
    S = GO.eval_S(g["evals_Lc"], 1.5, 5.0, 1.0, g["evecs_Lc"].shape[0])
    np.testing.assert_allclose(GO.K_diag(X, S), np.diag(GO.K(X, S)), rtol=1e-12)
    ev = np.linalg.eigvalsh(GO.K(X, S))
    assert ev.min() > -1e-8 * ev.max() and (ev > 1e-9 * ev.max()).sum() <= 20       # PSD, rank <= k
    Y = np.random.default_rng(1).normal(size=(300, 1))
    Xn = g["evecs_Lc"][300:500]
    for noise in (1.0, 1e-2):
        m1, v1 = GO.gpr_predict_dense(X, Y, S, noise, Xn)
        m2, v2 = GO.gpr_predict_lowrank(X.T @ X, X.T @ Y[:, 0], S, noise, Xn)
        np.testing.assert_allclose(m1, m2, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(v1, v2, rtol=1e-6, atol=1e-9)


def test_split_indices_match_sklearn():
    from sklearn.model_selection import train_test_split
    for n in (7, 64, 1000):
        for seed in (0, 3):
            tr, te = train_test_split(np.arange(n), test_size=0.2, random_state=seed)
            a, b = GO.train_test_split_indices(n, 0.2, seed)
            assert np.array_equal(tr, a) and np.array_equal(te, b)


def test_fit_improves_and_transform_shapes():
    g, n, train_ind = _c1_setup()
    gp = GO.train_gp(g["evecs_Lc"], g["evals_Lc"], g["smoothed_field"], n, train_ind, epochs=200, solver="lowrank")
    assert gp.opt_result.fun < 2512.0 and gp.l2_error < 0.2
    test_ind = [i for i in range(n) if i not in set(train_ind.tolist())][:50]
    m, v = GO.transform(gp, g["evecs_Lc"], n, test_ind)
    assert m.shape == (50, 3) and v.shape == (50, 3) and (v > 0).all()
    err = np.linalg.norm(m - g["smoothed_field"][test_ind], axis=1).mean()
    assert err < 0.2
