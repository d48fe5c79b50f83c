"""Small runs for compute-sanitizer, sized so that racecheck finishes (round 1's did not).
  k17_20 : K17 RBF Gram / adjoint, K18 SGPR + feature-space FPS (warp path), K19 div/curl features, K20 orientation + paired mode
  r2     : round-2 kernels -- filtered block Lanczos solver (real + paired) with the MMA pattern kernel for the scalar Laplacian,
           affinity graph, single-call rank-k GP evaluation (K15b, eager and CUDA-graph replay), geodesic source ranges
  r2b    : late round-2 kernels -- cp.async-pipelined and small-tile DMMA GEMMs (ragged shapes, split-K, accumulate), the
           register-resident 64x64 Cholesky + inverse, the fused forward / back-solve kernels of the rank-k GP evaluation
usage: python tools/sanitize_r2.py [k17_20] [r2] [r2b]"""
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
if "r2b" in what:
    from rvgp_b200._cabi import get_handle, I64
    from rvgp_b200.eigensolver import _dgemm
    from rvgp_b200.gp import DeviceGPR, _Chol
    h = get_handle(0)
    rng = np.random.default_rng(0)
    dev = torch.device("cuda", 0)
    for (m, n, k) in ((130, 66, 4098), (704, 64, 3000), (3000, 130, 66), (37, 53, 100), (244, 244, 256)):
        A = rng.normal(size=(m, k)); B = rng.normal(size=(k, n))
        Bd = torch.from_numpy(B).to(dev)
        for akm in (0, 1):
            Ad = torch.from_numpy(np.ascontiguousarray(A if akm else A.T)).to(dev)
            for split in (1, 3):
                C = torch.zeros((m, n), dtype=torch.float64, device=dev)
                ws = torch.empty(split * m * n, dtype=torch.float64, device=dev)
                _dgemm(h, m, n, k, Ad, Ad.stride(0), akm, Bd, Bd.stride(0), 0, C, C.stride(0), split_k=split, ws=ws)
                err = float(np.abs(C.cpu().numpy() - A @ B).max())
                assert err < 1e-9, (m, n, k, akm, split, err)
            h.call("rvgp_dgemm_acc_f64", int(m), int(n), I64(k), -1.0, Ad, I64(Ad.stride(0)), int(akm), Bd, I64(n), 0, 1.0, C, I64(n))
            assert float(C.abs().max()) < 1e-9
    print("dgemm pipelined / small-tile ok")
    for nn_ in (64, 200):
        M = rng.normal(size=(nn_, nn_ + 3)); M = M @ M.T + 0.5 * np.eye(nn_)
        Md = torch.from_numpy(M).to(dev)
        ch = _Chol(h, Md, nn_); ch.check()
        assert np.abs(np.tril(Md.cpu().numpy()) - np.linalg.cholesky(M)).max() < 1e-10
    print("potrf ok")
    for kk in (130, 200):
        Xg = torch.randn((400, kk), dtype=torch.float64, device="cuda")
        Yg = torch.randn((400, 1), dtype=torch.float64, device="cuda")
        gpr = DeviceGPR(Xg, Yg, solver="lowrank")
        for i in range(3):                                  # eager, capture, replay
            print("K15b fused", kk, gpr.lml_and_grads(np.linspace(1.0, 2.0, kk), 0.5 + 0.1 * i)[0])
print("SANITIZE_SCRIPT_DONE")
