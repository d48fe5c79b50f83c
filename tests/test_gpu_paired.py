"""GPU tests of K20 (gauge orientation) and the paired / complex-Hermitian eigensolver mode: same eigenpairs as the real
block solver and as the reference's ARPACK golden output (eigenvalues 1e-8 rel, eigen-subspace angles < 1e-6), fall-back on
surfaces where no consistent orientation exists."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import load_golden, subspace_angle_max, eigen_clusters

pytestmark = pytest.mark.gpu


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(torch.device("cuda", 0))
    return t if dtype is None else t.to(dtype)


def _host_orientation(indptr, indices, vals):
    """BFS restatement of the orientation problem (SciPy): signs or None."""
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    n = len(indptr) - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    sgn = np.sign(vals[:, 0, 0] * vals[:, 1, 1] - vals[:, 0, 1] * vals[:, 1, 0])
    A1 = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(n, n))
    order, pred = cg.breadth_first_order(A1, 0, directed=False)
    key = rows.astype(np.int64) * n + indices
    srt = np.argsort(key)
    ks = key[srt]
    s = np.ones(n)
    for v in order[1:]:
        e = srt[np.searchsorted(ks, pred[v] * n + v)]
        s[v] = s[pred[v]] * sgn[e]
    off = rows != indices
    return s if not np.any((s[rows] * s[indices] * sgn)[off] < 0) else None


@pytest.mark.parametrize("case", ["sphere_n2000_k50", "torus_n600_k20", "sheet_R20_n500_k16"])
def test_orientation_matches_host_bfs(case):
    from rvgp_b200 import geometry as geo
    g = load_golden(case)
    ip, ix, vals = g["Lc_indptr"], g["Lc_indices"], g["Lc_data"]
    s = geo.orient_gauges_device(_t(ip, torch.int32), _t(ix, torch.int32), _t(vals))
    ref = _host_orientation(ip, ix, vals)
    assert (s is None) == (ref is None)
    if s is not None:
        sh = s.cpu().numpy().astype(np.float64)
        assert set(np.unique(sh)) <= {-1.0, 1.0} and sh[0] == 1.0
        assert np.array_equal(sh, ref)                      # connected graph, seed node 0 -> unique solution
        # flipped blocks are scaled rotations [[a, -b], [b, a]]
        rows = np.repeat(np.arange(len(ip) - 1), np.diff(ip))
        v2 = vals.copy()
        v2[:, 0, 1] *= sh[ix]; v2[:, 1, 0] *= sh[rows]; v2[:, 1, 1] *= sh[rows] * sh[ix]
        assert np.abs(v2[:, 0, 0] - v2[:, 1, 1]).max() < 1e-12 and np.abs(v2[:, 0, 1] + v2[:, 1, 0]).max() < 1e-12


def test_rot90_and_flip_kernels():
    from rvgp_b200._cabi import get_handle, I64
    h = get_handle(0)
    rng = np.random.default_rng(0)
    V = rng.normal(size=(2 * 37, 9))
    Vd = _t(V)
    out = torch.empty_like(Vd)
    h.call("rvgp_rot90_nodes_f64", I64(37), 9, Vd, I64(9), out, I64(9))
    ref = np.empty_like(V); ref[0::2] = -V[1::2]; ref[1::2] = V[0::2]
    assert np.array_equal(out.cpu().numpy(), ref)
    s = rng.choice([-1, 1], size=37).astype(np.int32)
    h.call("rvgp_flip_odd_rows_f64", I64(37), 9, _t(s), Vd, I64(9))
    ref2 = V.copy(); ref2[1::2] *= s[:, None]
    assert np.array_equal(Vd.cpu().numpy(), ref2)


def test_pair_panel_and_combine_kernels_give_the_complex_products():
    """The two halves of the batched complex block algebra (krylov.FieldOps): with z = x + i y stored as the real (2n)-vector
    (x_0, y_0, x_1, y_1, ...) and i z = J z,  V^T [W | J W] = [Re V^H W | -Im V^H W]  and  T_re + J T_im = V C  for
    T = V [Re C | Im C]; strided inputs, both beta branches of the combine."""
    from rvgp_b200._cabi import get_handle, I64
    from rvgp_b200.krylov import FieldOps
    h = get_handle(0)
    rng = np.random.default_rng(3)
    n, b, cur = 41, 6, 10
    W = rng.normal(size=(2 * n, b + 3)); V = rng.normal(size=(2 * n, cur))
    Wd, Vd = _t(W), _t(V)
    P = torch.zeros((2 * n, 2 * b + 1), dtype=torch.float64, device=Wd.device)
    h.call("rvgp_pair_panel_f64", I64(n), b, Wd, I64(b + 3), P, I64(2 * b + 1))
    Ph = P.cpu().numpy()
    JW = np.empty((2 * n, b)); JW[0::2] = -W[1::2, :b]; JW[1::2] = W[0::2, :b]
    assert np.array_equal(Ph[:, :b], W[:, :b]) and np.array_equal(Ph[:, b:2 * b], JW) and not Ph[:, 2 * b].any()
    cplx = lambda A: A[0::2] + 1j * A[1::2]
    Vc, Wc = cplx(V), cplx(W[:, :b])
    C = Vc.conj().T @ Wc
    G2 = V.T @ Ph[:, :2 * b]
    np.testing.assert_allclose(G2[:, :b], C.real, atol=1e-13)
    np.testing.assert_allclose(G2[:, b:], -C.imag, atol=1e-13)
    T = _t(V @ np.concatenate([C.real, C.imag], 1))
    for beta in (0.0, 1.0):
        out = _t(W[:, :b].copy())
        h.call("rvgp_pair_combine_f64", I64(n), b, T, I64(2 * b), -1.0, beta, out, I64(b))
        np.testing.assert_allclose(cplx(out.cpu().numpy()), beta * Wc - Vc @ C, atol=1e-12)
    # and through the block operations themselves: one Gram-Schmidt pass leaves W orthogonal to V in the complex inner product
    Q = np.linalg.qr(Vc)[0]
    Vq = np.empty((2 * n, cur)); Vq[0::2] = Q.real; Vq[1::2] = Q.imag
    ops = FieldOps(h, 2 * n, 16, 8, Wd.device, None, True)
    Wb = _t(W[:, :b].copy())
    Ch = ops.project_out(_t(Vq), Wb)
    np.testing.assert_allclose(Ch, Q.conj().T @ Wc, atol=1e-12)
    assert np.abs(Q.conj().T @ cplx(Wb.cpu().numpy())).max() < 1e-12
    G = ops.gram_self(Wb)
    Wn = cplx(Wb.cpu().numpy())
    np.testing.assert_allclose(G, Wn.conj().T @ Wn, atol=1e-12)


def test_paired_solver_equals_real_solver_and_golden():
    """Direct call on the golden sphere Lc: re-orient, solve in paired mode, un-flip, compare with ARPACK's golden output."""
    from rvgp_b200 import geometry as geo
    from rvgp_b200.eigensolver import BsrMatrix, smallest_eigenpairs, smallest_eigenpairs_paired
    from rvgp_b200._cabi import get_handle, I64
    g = load_golden("sphere_n2000_k50")
    ip, ix, vals = _t(g["Lc_indptr"], torch.int32), _t(g["Lc_indices"], torch.int32), _t(g["Lc_data"])
    n = 2000
    s = geo.orient_gauges_device(ip, ix, vals)
    assert s is not None
    sd = s.to(torch.float64)
    rows = torch.repeat_interleave(torch.arange(n, device=ip.device), (ip[1:] - ip[:-1]).long())
    v2 = vals.clone()
    v2[:, 0, 1] *= sd[ix.long()]; v2[:, 1, 0] *= sd[rows]; v2[:, 1, 1] *= sd[rows] * sd[ix.long()]
    hi = 2.0 * float((ip[1:] - ip[:-1]).max() - 1)
    for k in (50, 49):                                     # even / odd number of requested pairs
        for mma in (False, True):
            A2 = BsrMatrix(n, 2, ip, ix, v2)
            if mma:
                assert A2.enable_mma() is not None
            st = {}
            ev, U = smallest_eigenpairs_paired(A2, k, upper_bound=hi, stats=st)
            assert st["converged"] and st["paired"] and U.shape == (2 * n, k)
            get_handle(0).call("rvgp_flip_odd_rows_f64", I64(n), int(k), s, U, I64(U.stride(0)))
            e = ev.cpu().numpy()
            np.testing.assert_allclose(e, g["evals_Lc"][:k], rtol=1e-8, atol=1e-12)
            A = BsrMatrix(n, 2, ip, ix, vals)
            Res = A.matmat(U.contiguous()) - U * ev
            assert float(torch.linalg.vector_norm(Res, dim=0).max()) <= 2e-12 * hi
            Uh = U.cpu().numpy()
            assert np.abs(Uh.T @ Uh - np.eye(k)).max() < 1e-10
            # compare eigen-subspaces with the reference in its own (tangent-coordinate) representation: evecs_Lc is the
            # ambient lift gauges @ U * sqrt(N), so lift ours the same way
            G = g["gauges"]
            lift = np.einsum("bij,bjk->bik", G, Uh.reshape(n, 2, k)).reshape(n * 3, k)
            ref = g["evecs_Lc"][:, :k]
            for sl in eigen_clusters(g["evals_Lc"][:k]):
                if sl.stop == k and k % 2:                 # a pair cut at the boundary: only one vector of a 2-dim space
                    continue
                assert subspace_angle_max(lift[:, sl], ref[:, sl]) < 1e-6
    # and the real block solver gives the same eigenvalues
    ev_r, _ = smallest_eigenpairs(BsrMatrix(n, 2, ip, ix, vals), 50, upper_bound=hi)
    np.testing.assert_allclose(ev_r.cpu().numpy(), e if len(e) == 50 else g["evals_Lc"][:50], rtol=1e-8, atol=1e-12)


def test_data_object_paired_vs_unpaired_and_nonorientable(monkeypatch):
    import RVGP
    from tests.workloads import make_cloud
    X = make_cloud("torus", 6000, 0)
    monkeypatch.setenv("RVGP_PAIRED", "1")
    d1 = RVGP.create_data_object(X, n_eigenpairs=60, verbose=False)
    monkeypatch.setenv("RVGP_PAIRED", "0")
    d0 = RVGP.create_data_object(X, n_eigenpairs=60, verbose=False)
    assert d1.stats["paired"] is True and d0.stats["paired"] is False
    assert d1.stats["eig_Lc"]["m"] * 2 <= d0.stats["eig_Lc"]["m"] + 64
    np.testing.assert_allclose(d1.evals_Lc, d0.evals_Lc, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(d1.gauges, d0.gauges)                      # the re-oriented frames stay internal
    for sl in eigen_clusters(d0.evals_Lc):
        if sl.stop == 60:
            continue
        assert subspace_angle_max(d1.evecs_Lc[:, sl], d0.evecs_Lc[:, sl]) < 1e-6
    # smoothing uses the local-coordinate eigenvectors with the ORIGINAL operator: identical fields
    for d in (d0, d1):
        d.random_vector_field(seed=3)
        d.smooth_vector_field(t=50)
    np.testing.assert_allclose(d1.vectors, d0.vectors, atol=1e-9)
    # Moebius strip: no consistent orientation -> the real block solver is used, results still valid
    monkeypatch.setenv("RVGP_PAIRED", "1")
    Xm = make_cloud("moebius", 4000, 0)
    dm = RVGP.create_data_object(Xm, n_eigenpairs=30, verbose=False)
    assert dm.dim_man == 2 and dm.stats["paired"] is False and dm.stats["eig_Lc"]["converged"]


def test_vectorfield_features_match_reference_golden():
    """K19 against golden vectors produced by the UNMODIFIED reference function (tests/golden/make_golden_eeg.py)."""
    from rvgp_b200.eeg_utils import compute_vectorfield_features, compute_vectorfield_features_time
    from oracle import rvgp_oracle as O
    g = load_golden("eeg_features")
    P, V = g["positions"], g["vectors"]
    for k in (5, 3):
        div, curl = compute_vectorfield_features(P, V, k=k)
        np.testing.assert_allclose(div, g["div_k%d" % k], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(curl, g["curl_k%d" % k], rtol=1e-12, atol=1e-12)
    div, curl = compute_vectorfield_features(P, V, k=5, reference_row0=False)
    rd, rc = O.compute_vectorfield_features(P, V, k=5, reference_row0=False)
    np.testing.assert_allclose(div, rd, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(curl, rc, rtol=1e-12, atol=1e-12)
    Vt = np.stack([V, V[::-1].copy(), 2.0 * V])
    dt, ct = compute_vectorfield_features_time(range(3), P, Vt)
    np.testing.assert_allclose(dt[0], g["div_k5"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dt[2], g["div_k5"], rtol=1e-12, atol=1e-12)   # direction features ignore the magnitude
    assert ct.shape == (3, len(P), 3)
    with pytest.raises(ValueError):
        compute_vectorfield_features(P[:, :2], V[:, :2])
