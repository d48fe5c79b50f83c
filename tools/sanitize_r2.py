"""Small runs for compute-sanitizer, sized so that racecheck finishes (round 1's did not).
  k17_20 : K17 RBF Gram / adjoint, K18 SGPR + feature-space FPS (warp path), K19 div/curl features, K20 orientation + paired mode
  r2     : round-2 kernels -- filtered block Lanczos solver (real + paired) with the MMA pattern kernel for the scalar Laplacian,
           affinity graph, single-call rank-k GP evaluation (K15b, eager and CUDA-graph replay), geodesic source ranges
usage: python tools/sanitize_r2.py [k17_20] [r2]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import RVGP
from rvgp_b200 import params as P
from tests.conftest import load_golden
from tests.workloads import make_cloud

what = sys.argv[1:] or ["k17_20", "r2"]
P.set_default_positive_minimum(0.0)
if "k17_20" in what:
    gs = load_golden("sphere_n2000_k50")
    d = RVGP.create_data_object(gs["X"][:400], n_eigenpairs=12, verbose=False)
    print("paired", d.stats["paired"], "converged", d.stats["eig_Lc"]["converged"])
    d.random_vector_field(seed=1)
    tr = np.arange(0, d.n, 2)
    for kern, nind in (("rbf", None), (None, 8), ("rbf", 6)):
        gp = RVGP.fit(d, train_ind=tr, kernel=kern, n_inducing_points=nind, epochs=1)
        print(kern, nind, type(gp).__name__, gp.l2_error)
    from RVGP.geometry import furthest_point_sampling
    x = np.random.default_rng(0).normal(size=(120, 130))
    print("fps D=130", furthest_point_sampling(x, N=6)[0][:5])
    from rvgp_b200.eeg_utils import compute_vectorfield_features
    ge = load_golden("eeg_features")
    dv, cl = compute_vectorfield_features(ge["positions"], ge["vectors"], k=5)
    print("features", float(abs(dv - ge["div_k5"]).max()))
if "r2" in what:
    os.environ["RVGP_EIGSOLVER"] = "krylov"
    X = make_cloud("torus", 1500, 0)
    d = RVGP.create_data_object(X, n_eigenpairs=40, verbose=False)
    print("krylov", d.stats["eig_L"].get("solver"), d.stats["eig_L"]["spmm_kernel"], d.stats["eig_Lc"].get("solver"),
          d.stats["eig_Lc"]["spmm_kernel"], d.stats["eig_L"]["converged"], d.stats["eig_Lc"]["converged"])
    from rvgp_b200 import geometry as geo
    G = geo.manifold_graph(X[:200], typ="affinity")
    print("affinity", float(G.weights.sum()))
    from rvgp_b200.gp import DeviceGPR
    Xg = torch.randn((400, 80), dtype=torch.float64, device="cuda")
    Yg = torch.randn((400, 1), dtype=torch.float64, device="cuda")
    gpr = DeviceGPR(Xg, Yg, solver="lowrank")
    for i in range(3):                                      # eager, capture, replay
        print("K15b", gpr.lml_and_grads(np.linspace(1.0, 2.0, 80), 0.5 + 0.1 * i)[0])
    from rvgp_b200._cabi import get_handle, I64
    gr = d._graph
    seq, cnt = geo.geodesic_neighbourhoods_device(gr.indptr, gr.indices, 15)
    print("geodesic", int(cnt.min()))
print("SANITIZE_SCRIPT_DONE")
