"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import RVGP
from tests.conftest import load_golden
from RVGP.geometry import furthest_point_sampling
cases = (("torus_n600_k20", 10),) if "--one" in sys.argv else (("torus_n600_k20", 10), ("flat3torus_R6_n900_k24", 10), ("sheet_R20_n500_k16", 14))
for case, nb in cases:
    gd = load_golden(case)
    d = RVGP.create_data_object(gd["X"], n_neighbors=nb, n_eigenpairs=len(gd["evals_Lc"]), verbose=False)
    d.random_vector_field(seed=1); d.smooth_vector_field(t=10)
    gp = RVGP.fit(d, train_ind=np.arange(0, d.n, 2), epochs=10)
    m, v = gp.transform(d, np.arange(1, d.n, 2))
    print(case, "ok", float(abs(d.evals_Lc - gd["evals_Lc"]).max()))
p, l = furthest_point_sampling(gd["X"], spacing=0.2); print("fps", len(p))
from rvgp_b200.gp import DeviceGPR
X = torch.randn((300, 20), dtype=torch.float64, device="cuda"); Y = torch.randn((300, 1), dtype=torch.float64, device="cuda")
print(DeviceGPR(X, Y, solver="dense").lml_and_grads(np.ones(20), 0.5)[0])
if "--new" in sys.argv or "--all" in sys.argv:
    # kernels added late in round 1: K17 RBF Gram / adjoint, K18 SGPR + feature-space FPS (warp path, D > 64), K19 div/curl
    # features, K20 orientation + paired eigensolver (sphere: orientable -> paired mode)
    from rvgp_b200 import params as P
    P.set_default_positive_minimum(0.0)
    gs = load_golden("sphere_n2000_k50")
    d = RVGP.create_data_object(gs["X"][:900], n_eigenpairs=20, verbose=False)
    print("paired", d.stats["paired"], "converged", d.stats["eig_Lc"]["converged"])
    d.random_vector_field(seed=1); d.smooth_vector_field(t=10)
    tr = np.arange(0, d.n, 2)
    for kern, nind in (("rbf", None), (None, 12), ("rbf", 9)):
        gp = RVGP.fit(d, train_ind=tr, kernel=kern, n_inducing_points=nind, epochs=3)
        print(kern, nind, type(gp).__name__, gp.l2_error)
    x = np.random.default_rng(0).normal(size=(300, 130))
    print("fps D=130", furthest_point_sampling(x, N=10)[0][:5])
    from rvgp_b200.eeg_utils import compute_vectorfield_features
    ge = load_golden("eeg_features")
    dv, cl = compute_vectorfield_features(ge["positions"], ge["vectors"], k=5)
    print("features", float(abs(dv - ge["div_k5"]).max()))
