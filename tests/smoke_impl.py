"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np
import scipy.sparse as sp
import torch


def run_smoke():
    from tests.conftest import load_golden
    from rvgp_b200.eigensolver import BsrMatrix, smallest_eigenpairs
    from oracle import rvgp_oracle as O
    dev = torch.device("cuda", 0)
    g = load_golden("torus_n600_k20")
    n, d = g["X"].shape[0], int(g["dim_man"])
    # oracle (checker): Lc from the golden gauges, ARPACK spectrum
    R = O.connections(g["gauges"], g["indptr"], g["indices"])
    Lc = O.connection_laplacian(g["indptr"], g["indices"], R)
    ev_ref, _ = O.spectrum(Lc, 20)
    A = BsrMatrix(n, d, torch.from_numpy(g["indptr"]).to(dev), torch.from_numpy(g["indices"]).to(dev),
                  torch.from_numpy(np.ascontiguousarray(Lc.data)).to(dev))
    X = np.random.default_rng(0).normal(size=(n * d, 32))
    Y = A.spmm(torch.from_numpy(X).to(dev), torch.empty((n * d, 32), dtype=torch.float64, device=dev))
    assert np.abs(Y.cpu().numpy() - Lc @ X).max() < 1e-12
    ev, U = smallest_eigenpairs(A, 20, upper_bound=2.0 * (np.diff(g["indptr"]).max() - 1))
    np.testing.assert_allclose(ev.cpu().numpy(), ev_ref, rtol=1e-8, atol=1e-9)
    # round 2: the scalar Laplacian on the FP64-MMA kernel (pattern mode) and the filtered block Lanczos solver, same oracle
    gs = load_golden("sphere_n2000_k50")
    ns = gs["X"].shape[0]
    Ls = O.laplacian(gs["indptr"], gs["indices"])
    AL = BsrMatrix(ns, 1, torch.from_numpy(gs["indptr"]).to(dev), torch.from_numpy(gs["indices"]).to(dev), None)
    assert AL.enable_mma_pattern() is not None
    Xs = np.random.default_rng(1).normal(size=(ns, 64))
    Ys = AL.spmm_pattern(torch.from_numpy(Xs).to(dev), torch.empty((ns, 64), dtype=torch.float64, device=dev))
    assert np.abs(Ys.cpu().numpy() - Ls @ Xs).max() < 1e-12
    from rvgp_b200.krylov import krylov_eigenpairs
    evL_ref, _ = O.spectrum(Ls, 60)
    his = 2.0 * (np.diff(gs["indptr"]).max() - 1)
    evk, _ = krylov_eigenpairs(AL, 40, his, cut=1.05 * evL_ref[59], lam_k=evL_ref[39], block=16)
    np.testing.assert_allclose(evk.cpu().numpy(), evL_ref[:40], rtol=1e-8, atol=1e-9)
    print("smoke ok: SpMM (gather + MMA pattern) and both eigensolvers match the oracle; evals[:4] =", ev[:4].cpu().numpy())
