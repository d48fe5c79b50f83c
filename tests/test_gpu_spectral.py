"""GPU parity tests (through the C-ABI) for K9 block-SpMM, K10 dense kernels and the eigensolver."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests.conftest import load_golden, subspace_angle_max, eigen_clusters

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _bsr_from_golden(g, which="Lc"):
    from rvgp_b200.eigensolver import BsrMatrix
    n = g["X"].shape[0]
    if which == "Lc":
        d = int(g["dim_man"])
        vals = torch.from_numpy(g["Lc_data"]).to(_dev())
        A = BsrMatrix(n, d, torch.from_numpy(g["Lc_indptr"]).to(_dev()), torch.from_numpy(g["Lc_indices"]).to(_dev()), vals)
        S = sp.bsr_matrix((g["Lc_data"], g["Lc_indices"], g["Lc_indptr"]), shape=(n * d, n * d)).tocsr()
    else:
        A = BsrMatrix(n, 1, torch.from_numpy(g["L_indptr"]).to(_dev()), torch.from_numpy(g["L_indices"]).to(_dev()), None)
        S = sp.csr_matrix((g["L_data"], g["L_indices"], g["L_indptr"]), shape=(n, n))
    return A, S


@pytest.mark.parametrize("case", ["sphere_n2000_k50", "flat3torus_R6_n900_k24", "sheet_R20_n500_k16"])
@pytest.mark.parametrize("ncols", [1, 7, 16, 32, 50, 64])
def test_spmm_matches_scipy(case, ncols):
    g = load_golden(case)
    for which in ("Lc", "L"):
        A, S = _bsr_from_golden(g, which)
        rng = np.random.default_rng(1)
        X = rng.normal(size=(A.nrows, ncols))
        W = rng.normal(size=(A.nrows, ncols))
        Xd, Wd = torch.from_numpy(X).to(_dev()), torch.from_numpy(W).to(_dev())
        Yd = torch.empty_like(Xd)
        A.spmm(Xd, Yd)
        ref = S @ X
        assert np.abs(Yd.cpu().numpy() - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
        A.spmm(Xd, Yd, alpha=0.7, beta=-1.3, gamma=0.25, W=Wd)
        ref2 = 0.7 * ref - 1.3 * X + 0.25 * W
        assert np.abs(Yd.cpu().numpy() - ref2).max() <= 1e-13 * max(1.0, np.abs(ref2).max())


def test_spmm_strided_and_empty():
    g = load_golden("torus_n600_k20")
    A, S = _bsr_from_golden(g, "Lc")
    big = torch.zeros((A.nrows, 96), dtype=torch.float64, device=_dev())
    X = np.random.default_rng(0).normal(size=(A.nrows, 20))
    big[:, 10:30] = torch.from_numpy(X).to(_dev())
    out = torch.zeros((A.nrows, 40), dtype=torch.float64, device=_dev())
    A.spmm(big[:, 10:30], out[:, 5:25])
    assert np.abs(out[:, 5:25].cpu().numpy() - S @ X).max() < 1e-12
    assert out[:, :5].abs().max().item() == 0.0 and out[:, 25:].abs().max().item() == 0.0
    with pytest.raises(ValueError):
        A.spmm(big[:, :65], torch.empty((A.nrows, 65), dtype=torch.float64, device=_dev()))


@pytest.mark.parametrize("degree", [1, 2, 3, 7, 30])
def test_cheb_filter_matches_recurrence(degree):
    from rvgp_b200._cabi import get_handle, I64
    g = load_golden("torus_n600_k20")
    A, S = _bsr_from_golden(g, "Lc")
    X = np.random.default_rng(3).normal(size=(A.nrows, 24))
    hi, cut, lo = 40.0, 1.5, 0.0
    e, c = (hi - cut) / 2, (hi + cut) / 2
    s1 = e / (lo - c); tau = 2 / s1; sg = s1
    Xp, Y = X, (S @ X - c * X) * (s1 / e)
    for i in range(2, degree + 1):
        sn = 1 / (tau - sg)
        Xp, Y, sg = Y, (S @ Y - c * Y) * (2 * sn / e) - sg * sn * Xp, sn
    V = torch.from_numpy(X).to(_dev())
    w0 = torch.empty((A.nrows, 32), dtype=torch.float64, device=_dev()); w1 = torch.empty_like(w0)
    h = get_handle(0)
    h.call("rvgp_cheb_filter_f64", A.nbrows, A.d, A.indptr, A.indices, A.vals, V, I64(V.stride(0)), w0, w1, I64(32),
           24, degree, lo, cut, hi)
    assert np.abs(V.cpu().numpy() - Y).max() <= 1e-12 * np.abs(Y).max()


@pytest.mark.parametrize("dmma", [0, 1])
@pytest.mark.parametrize("shape", [(1, 1, 5), (37, 53, 1000), (200, 129, 4097), (256, 256, 70000), (130, 7, 333)])
def test_dgemm_layouts(shape, dmma):
    from rvgp_b200.eigensolver import _dgemm
    from rvgp_b200._cabi import get_handle
    m, n, k = shape
    h = get_handle(0)
    h.set_option("dgemm_dmma", dmma)
    rng = np.random.default_rng(0)
    for akm in (0, 1):
        for bkm in (0, 1):
            for split in (1, 5):
                A = rng.normal(size=(m, k)); B = rng.normal(size=(k, n)); s = rng.uniform(0.5, 2, size=k)
                Ad = torch.from_numpy(np.ascontiguousarray(A if akm else A.T)).to(_dev())
                Bd = torch.from_numpy(np.ascontiguousarray(B.T if bkm else B)).to(_dev())
                sd = torch.from_numpy(s).to(_dev())
                C = torch.zeros((m, n + 3), dtype=torch.float64, device=_dev())
                ws = torch.empty(split * m * n, dtype=torch.float64, device=_dev())
                _dgemm(h, m, n, k, Ad, Ad.stride(0), akm, Bd, Bd.stride(0), bkm, C, C.stride(0), alpha=0.5,
                       scale_k=sd, split_k=split, ws=ws)
                ref = 0.5 * (A * s) @ B
                err = np.abs(C[:, :n].cpu().numpy() - ref).max()
                assert err <= 1e-13 * k ** 0.5 * max(1, np.abs(ref).max()), (akm, bkm, split, err)
                assert C[:, n:].abs().max().item() == 0.0
    h.set_option("dgemm_dmma", 2)                                   # the handle is shared: back to the default


@pytest.mark.parametrize("shape", [(1, 1, 1), (37, 53, 100), (200, 129, 257), (436, 192, 64), (244, 244, 256)])
def test_dgemm_small_tile_kernel(shape):
    """The 32 x 32-tile DMMA kernel that small products (the Cholesky panels of the rank-k GP evaluation) are routed to: all four
    operand layouts, ragged edges, accumulate, lower-triangle-only; against numpy and the 128 x 64-tile kernel."""
    from rvgp_b200._cabi import get_handle, I64
    m, n, k = shape
    h = get_handle(0)
    rng = np.random.default_rng(2)
    A = rng.normal(size=(m, k)); B = rng.normal(size=(k, n)); C0 = rng.normal(size=(m, n))
    ref = C0 - A @ B
    tol = 1e-13 * k ** 0.5 * max(1, np.abs(ref).max())
    try:
        for akm in (0, 1):
            for bkm in (0, 1):
                Ad = torch.from_numpy(np.ascontiguousarray(A if akm else A.T)).to(_dev())
                Bd = torch.from_numpy(np.ascontiguousarray(B.T if bkm else B)).to(_dev())
                outs = []
                for mode in (2, 1):
                    h.set_option("dgemm_dmma", mode)
                    C = torch.zeros((m, n + 1), dtype=torch.float64, device=_dev())
                    C[:, :n] = torch.from_numpy(C0).to(_dev())
                    h.call("rvgp_dgemm_acc_f64", int(m), int(n), I64(k), -1.0, Ad, I64(Ad.stride(0)), int(akm), Bd, I64(Bd.stride(0)),
                           int(bkm), 1.0, C, I64(n + 1))
                    assert C[:, n].abs().max().item() == 0.0
                    outs.append(C[:, :n].cpu().numpy())
                    assert np.abs(outs[-1] - ref).max() <= tol, (akm, bkm, mode)
                assert np.abs(outs[0] - outs[1]).max() <= tol
        if m == n:
            h.set_option("dgemm_dmma", 2)
            Ad = torch.from_numpy(A).to(_dev())
            G = torch.full((m, m), 7.0, dtype=torch.float64, device=_dev())
            h.call("rvgp_dgemm_lower_f64", int(m), int(m), I64(k), 1.0, Ad, I64(k), 1, Ad, I64(k), 1, G, I64(m), 1, None)
            Gh = G.cpu().numpy(); full = A @ A.T
            il = np.tril_indices(m)
            assert np.abs(Gh[il] - full[il]).max() <= 1e-13 * k ** 0.5 * np.abs(full).max()
            iu = np.triu_indices(m, 32)                               # tiles strictly above the diagonal are never touched
            assert (Gh[iu] == 7.0).all()
    finally:
        h.set_option("dgemm_dmma", 2)


@pytest.mark.parametrize("shape", [(2, 2, 2), (64, 64, 100001), (130, 66, 4098), (704, 64, 50000), (50001 * 2, 64, 192),
                                   (3000, 130, 66), (256, 256, 70000)])
def test_dgemm_pipelined_kernel(shape):
    """The cp.async-pipelined DMMA kernel (dgemm_dmma = 2: even extents, 16-byte aligned operands, N-major B) against numpy and
    against the register-staged kernel; ragged M / N / K tails, split-K, the accumulate entry and the lower-triangle entry."""
    from rvgp_b200.eigensolver import _dgemm
    from rvgp_b200._cabi import get_handle, I64
    m, n, k = shape
    h = get_handle(0)
    rng = np.random.default_rng(1)
    A = rng.normal(size=(m, k)); B = rng.normal(size=(k, n)); C0 = rng.normal(size=(m, n))
    ref = A @ B
    tol = 1e-13 * k ** 0.5 * max(1, np.abs(ref).max())
    Bd = torch.from_numpy(B).to(_dev())
    try:
        for akm in (0, 1):
            Ad = torch.from_numpy(np.ascontiguousarray(A if akm else A.T)).to(_dev())
            for split in (1, 3):
                outs = []
                for mode in (2, 1):
                    h.set_option("dgemm_dmma", mode)
                    C = torch.zeros((m, n + 2), dtype=torch.float64, device=_dev())
                    ws = torch.empty(split * m * n, dtype=torch.float64, device=_dev())
                    l0 = h.launches
                    _dgemm(h, m, n, k, Ad, Ad.stride(0), akm, Bd, Bd.stride(0), 0, C, C.stride(0), alpha=0.5, split_k=split, ws=ws)
                    assert h.launches - l0 == (2 if split > 1 else 1)
                    assert C[:, n:].abs().max().item() == 0.0
                    outs.append(C[:, :n].cpu().numpy())
                    assert np.abs(outs[-1] - 0.5 * ref).max() <= tol, (akm, split, mode)
                assert np.abs(outs[0] - outs[1]).max() <= tol
            # accumulate: C = -A B + C  (the eigensolver's W -= V C update)
            h.set_option("dgemm_dmma", 2)
            C = torch.from_numpy(C0).to(_dev())
            h.call("rvgp_dgemm_acc_f64", int(m), int(n), I64(k), -1.0, Ad, I64(Ad.stride(0)), int(akm), Bd, I64(n), 0, 1.0, C, I64(n))
            assert np.abs(C.cpu().numpy() - (C0 - ref)).max() <= tol
        if m == n:                                                  # SYRK-style entry: lower-triangle tiles only
            Ad = torch.from_numpy(np.ascontiguousarray(A.T)).to(_dev())
            G = torch.full((m, m), 7.0, dtype=torch.float64, device=_dev())
            ws = torch.empty(3 * m * m, dtype=torch.float64, device=_dev())
            h.call("rvgp_dgemm_lower_f64", int(m), int(m), I64(k), 1.0, Ad, I64(m), 0, Ad, I64(m), 0, G, I64(m), 3, ws)
            Gh = G.cpu().numpy(); full = A @ A.T
            il = np.tril_indices(m)
            assert np.abs(Gh[il] - full[il]).max() <= 1e-13 * k ** 0.5 * np.abs(full).max()
    finally:
        h.set_option("dgemm_dmma", 2)


def test_column_reductions_and_utils():
    from rvgp_b200._cabi import get_handle, I64, U64
    h = get_handle(0)
    N, m = 10007, 77
    rng = np.random.default_rng(0)
    A = rng.normal(size=(N, m)); B = rng.normal(size=(N, m)); th = rng.normal(size=m)
    Ad, Bd, thd = (torch.from_numpy(x).to(_dev()) for x in (A, B, th))
    ws = torch.empty(h.query("rvgp_coldot_workspace_bytes", I64(N), m) // 8, dtype=torch.float64, device=_dev())
    out = torch.empty(m, dtype=torch.float64, device=_dev())
    h.call("rvgp_coldot_f64", I64(N), m, Ad, I64(m), Bd, I64(m), out, ws)
    np.testing.assert_allclose(out.cpu().numpy(), (A * B).sum(0), rtol=1e-12, atol=1e-10)
    h.call("rvgp_resid_sq_f64", I64(N), m, Ad, I64(m), Bd, I64(m), thd, out, ws)
    np.testing.assert_allclose(out.cpu().numpy(), ((A - th * B) ** 2).sum(0), rtol=1e-12)
    h.call("rvgp_colscale_f64", I64(N), m, Ad, I64(m), thd)
    np.testing.assert_allclose(Ad.cpu().numpy(), A * th, rtol=1e-15)
    F1 = torch.empty((N, m), dtype=torch.float64, device=_dev()); F2 = torch.empty((N, 10), dtype=torch.float64, device=_dev())
    h.call("rvgp_fill_uniform_f64", I64(N), m, F1, I64(m), U64(7), I64(0), I64(0))
    h.call("rvgp_fill_uniform_f64", I64(N), 10, F2, I64(10), U64(7), I64(20), I64(0))
    F3 = torch.empty((100, m), dtype=torch.float64, device=_dev())
    h.call("rvgp_fill_uniform_f64", I64(100), m, F3, I64(m), U64(7), I64(0), I64(5000))
    assert np.array_equal(F3.cpu().numpy(), F1.cpu().numpy()[5000:5100])      # row-sharding invariant
    f1 = F1.cpu().numpy()
    assert np.array_equal(f1[:, 20:30], F2.cpu().numpy())          # counter based: depends on (seed,row,col) only
    assert abs(f1.mean()) < 0.01 and f1.min() >= -1 and f1.max() < 1 and abs(f1.std() - 3 ** -0.5) < 0.01
    perm = torch.from_numpy(rng.permutation(N // 1).astype(np.int32)).to(_dev())
    G = torch.empty_like(Bd)
    h.call("rvgp_gather_rows_f64", I64(N), m, Bd, I64(m), perm, 1, G, I64(m))
    assert np.array_equal(G.cpu().numpy(), B[perm.cpu().numpy()])


@pytest.mark.parametrize("case", ["sphere_n2000_k50", "torus_n600_k20", "flat3torus_R6_n900_k24", "sheet_R20_n500_k16"])
def test_eigensolver_matches_reference_arpack(case):
    """Eigenvalues rel 1e-8 (absolute 1e-8*|L| for the zero mode), eigen-subspace angles < 1e-6 per cluster
    against the golden ARPACK output of the unmodified reference (geometry.py:73)."""
    from rvgp_b200.eigensolver import smallest_eigenpairs
    g = load_golden(case)
    k = len(g["evals_Lc"])
    n = g["X"].shape[0]
    d = int(g["dim_man"])
    maxdeg = int(np.diff(g["indptr"]).max() - 1)
    for which, ev_ref in (("Lc", g["evals_Lc"]), ("L", g["evals_L"])):
        A, S = _bsr_from_golden(g, which)
        st = {}
        ev, U = smallest_eigenpairs(A, k, upper_bound=2.0 * maxdeg, stats=st)
        ev = ev.cpu().numpy(); U = U.cpu().numpy()
        assert st["converged"], st
        np.testing.assert_allclose(ev, ev_ref, rtol=1e-8, atol=1e-8 * 2 * maxdeg)
        assert np.abs(U.T @ U - np.eye(k)).max() < 1e-10
        assert np.abs(S @ U - U * ev).max() < 1e-8
        if which == "Lc":
            Phi = np.einsum("bij,bjk->bik", g["gauges"], (U * np.sqrt(n * d)).reshape(n, d, k)).reshape(-1, k)
            ref = g["evecs_Lc"]
        else:
            Phi, ref = U * np.sqrt(n), g["evecs_L"]
        for s in eigen_clusters(ev_ref)[:-1]:
            assert subspace_angle_max(Phi[:, s], ref[:, s]) < 1e-6, (which, s)


@pytest.mark.parametrize("ncols", [2, 16, 32, 64])
def test_spmm_rot2_storage(ncols):
    """ROT2 (a, b, flip) storage of the 2x2 connection blocks gives the same product as the plain blocks."""
    g = load_golden("sphere_n2000_k50")
    A, S = _bsr_from_golden(g, "Lc")
    assert A.compress_rot2() and A.d_code == -2
    rng = np.random.default_rng(4)
    X = rng.normal(size=(A.nrows, ncols)); W = rng.normal(size=(A.nrows, ncols))
    Xd, Wd = torch.from_numpy(X).to(_dev()), torch.from_numpy(W).to(_dev())
    Yd = torch.empty_like(Xd)
    A.spmm(Xd, Yd, alpha=0.7, beta=-1.3, gamma=0.25, W=Wd)
    ref = 0.7 * (S @ X) - 1.3 * X + 0.25 * W
    assert np.abs(Yd.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    # a matrix whose blocks are NOT scaled rotations / reflections is refused
    from rvgp_b200.eigensolver import BsrMatrix
    vals = torch.from_numpy(g["Lc_data"].copy()); vals[5, 0, 1] += 0.3
    B = BsrMatrix(A.nbrows, 2, A.indptr, A.indices, vals.to(_dev()))
    assert not B.compress_rot2() and B.d_code == 2


@pytest.mark.parametrize("case", ["sphere_n2000_k50", "flat3torus_R6_n900_k24", "torus_n600_k20"])
@pytest.mark.parametrize("ncols", [16, 32, 48, 64, 128])
def test_spmm_mma_rowmajor_matches_scipy(case, ncols):
    """K9 v4a (row-major FP64 mma.sync row-group SpMM, kept as an experiment) == SciPy for the d = 2 connection
    Laplacian and the pattern-mode L, on contiguous blocks and on column panels of a wider block vector; shapes the
    MMA path cannot take fall back."""
    g = load_golden(case)
    for which in ("Lc", "L"):
        A, S = _bsr_from_golden(g, which)
        if A.d not in (1, 2):
            assert A.enable_mma(rowmajor=True) is None
            continue
        mp = A.enable_mma(rowmajor=True)
        assert mp["ksteps"] > 0 and 0.0 < mp["fill"] <= 1.0 and mp["afrag"] is not None
        rng = np.random.default_rng(5)
        X = rng.normal(size=(A.nrows, ncols)); W = rng.normal(size=(A.nrows, ncols))
        big = torch.zeros((A.nrows, ncols + 32), dtype=torch.float64, device=_dev())
        big[:, 16:16 + ncols] = torch.from_numpy(X).to(_dev())
        Wd = torch.from_numpy(W).to(_dev())
        for Xd in (torch.from_numpy(X).to(_dev()), big[:, 16:16 + ncols]):
            assert A._mma_ok(ncols, Xd, Wd)
            Yd = torch.full((A.nrows, ncols), 7.0, dtype=torch.float64, device=_dev())
            A.spmm(Xd, Yd)
            ref = S @ X
            assert np.abs(Yd.cpu().numpy() - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
            A.spmm(Xd, Yd, alpha=0.7, beta=-1.3, gamma=0.25, W=Wd)
            ref2 = 0.7 * ref - 1.3 * X + 0.25 * W
            assert np.abs(Yd.cpu().numpy() - ref2).max() <= 1e-13 * max(1.0, np.abs(ref2).max())
        # unaligned panel / odd width: falls back to the gather kernel, same answer
        Xo = big[:, 3:3 + 10]
        assert not A._mma_ok(10, Xo)
        Yo = torch.empty((A.nrows, 10), dtype=torch.float64, device=_dev())
        A.spmm(Xo, Yo)
        assert np.abs(Yo.cpu().numpy() - S @ big[:, 3:13].cpu().numpy()).max() <= 1e-12


def _native_cases():
    return ["sphere_n2000_k50", "torus_n600_k20"]


@pytest.mark.parametrize("case", _native_cases())
@pytest.mark.parametrize("ncols", [16, 32, 64, 96, 128])
@pytest.mark.parametrize("compact", [True, False])
def test_spmm_mma_native_matches_scipy(case, ncols, compact):
    """K9 v4 (shipped): FP64-MMA SpMM on node-contiguous panels == SciPy for the d = 2 connection Laplacian, with the full
    and the compact (rotation + sign bit) fragment plans, all schedule variants, with / without W, and for a row count that
    is not a multiple of the group size."""
    from rvgp_b200.eigensolver import BsrMatrix
    from rvgp_b200._cabi import get_handle
    g = load_golden(case)
    A, S = _bsr_from_golden(g, "Lc")
    assert A.d == 2
    h = get_handle(0)
    for drop in (0, 3):       # drop trailing nodes: nbrows % 4 != 0
        if drop:
            n2 = A.nbrows - drop
            ip = A.indptr[: n2 + 1].clone()
            keep = (A.indices[: int(ip[-1])] < n2)
            # rebuild a consistent CSR on the first n2 nodes
            rows = torch.repeat_interleave(torch.arange(n2, device=_dev()), (ip[1:] - ip[:-1]).long())[keep]
            ix = A.indices[: int(ip[-1])][keep].contiguous()
            vl = A.vals[: int(ip[-1])][keep].contiguous()
            ip2 = torch.zeros(n2 + 1, dtype=torch.int32, device=_dev())
            ip2[1:] = torch.cumsum(torch.bincount(rows, minlength=n2), 0).to(torch.int32)
            B = BsrMatrix(n2, 2, ip2, ix, vl)
            Sm = sp.bsr_matrix((vl.cpu().numpy(), ix.cpu().numpy(), ip2.cpu().numpy()), shape=(2 * n2, 2 * n2)).tocsr()
        else:
            B, Sm = A, S
        B._mma_plan = None
        B.mma = B.build_mma_plan(compact=compact)
        assert B.mma["rotc"] == (1 if compact else 0)
        rng = np.random.default_rng(7)
        X = rng.normal(size=(B.nrows, ncols)); W = rng.normal(size=(B.nrows, ncols))
        Xd, Wd = torch.from_numpy(X).to(_dev()), torch.from_numpy(W).to(_dev())
        Xn, Wn = B.to_native(Xd), B.to_native(Wd)
        assert np.array_equal(B.from_native(Xn).cpu().numpy(), X)
        ref = Sm @ X
        try:
            for var in (0, 1, 2, 3, 5, 6):
                for gpw in (0, 2):
                    h.set_option("mma_variant", var); h.set_option("mma_gpw", gpw)
                    Yn = torch.full_like(Xn, 7.0)
                    B.spmm_native(Xn, Yn)
                    assert np.abs(B.from_native(Yn).cpu().numpy() - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
                    B.spmm_native(Xn, Yn, alpha=0.7, beta=-1.3, gamma=0.25, Wn=Wn, reverse=(var & 1))
                    ref2 = 0.7 * ref - 1.3 * X + 0.25 * W
                    assert np.abs(B.from_native(Yn).cpu().numpy() - ref2).max() <= 1e-13 * max(1.0, np.abs(ref2).max())
        finally:
            h.set_option("mma_variant", 1); h.set_option("mma_gpw", 0)
    with pytest.raises(ValueError):
        A.spmm_native(Xn[:, :20], Yn[:, :20])          # 10 columns: not a multiple of 16


@pytest.mark.parametrize("degree", [1, 2, 9, 40])
def test_cheb_filter_mma_matches_gather(degree):
    """The native-layout MMA filter (Lc, d = 2) and the row-major MMA filter (L) agree with the gather-kernel filter."""
    g = load_golden("sphere_n2000_k50")
    for which in ("Lc", "L"):
        A, _ = _bsr_from_golden(g, which)
        V0 = torch.from_numpy(np.random.default_rng(2).normal(size=(A.nrows, 96))).to(_dev())
        outs = []
        for on in (False, True):
            A.enable_mma(on, rowmajor=(which == "L"))
            V = V0.clone()
            w0 = torch.empty((A.nrows, 32), dtype=torch.float64, device=_dev()); w1 = torch.empty_like(w0); w2 = torch.empty_like(w0)
            if on and which == "Lc":
                assert A._mma_native_ok(32, V[:, 32:64], w0, w1, w2)
            A.cheb_filter(V[:, 32:64], w0, w1, 32, degree, 0.0, 3.0, 40.0, w2=w2)
            outs.append(V.cpu().numpy())
        assert np.abs(outs[0][:, :32] - V0[:, :32].cpu().numpy()).max() == 0.0
        assert np.abs(outs[0][:, 64:] - V0[:, 64:].cpu().numpy()).max() == 0.0
        assert np.abs(outs[0] - outs[1]).max() <= 1e-11 * np.abs(outs[0]).max()


@pytest.mark.parametrize("case", ["sphere_n2000_k50", "flat3torus_R6_n900_k24", "torus_n600_k20", "sheet_R20_n500_k16"])
@pytest.mark.parametrize("ncols", [32, 64, 96, 128])
def test_spmm_mma_pattern_matches_scipy(case, ncols):
    """K9 for the scalar Laplacian on the FP64-MMA native kernel (L (x) I_2, no matrix values streamed; spmm_mma.cu AMODE 2):
    equals SciPy's L @ X on ROW-MAJOR panels, plain and fused (alpha, beta, gamma), contiguous and strided."""
    g = load_golden(case)
    A, S = _bsr_from_golden(g, "L")
    assert A.enable_mma_pattern() is not None
    rng = np.random.default_rng(3)
    X = rng.normal(size=(A.nrows, ncols))
    W = rng.normal(size=(A.nrows, ncols))
    Xd, Wd = torch.from_numpy(X).to(_dev()), torch.from_numpy(W).to(_dev())
    Yd = torch.full_like(Xd, float("nan"))
    A.spmm_pattern(Xd, Yd)
    ref = S @ X
    assert np.abs(Yd.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()
    A.spmm_pattern(Xd, Yd, alpha=0.7, beta=-1.3, gamma=0.25, W=Wd)
    ref2 = 0.7 * ref - 1.3 * X + 0.25 * W
    assert np.abs(Yd.cpu().numpy() - ref2).max() <= 1e-13 * np.abs(ref2).max()
    A.spmm_pattern(Xd, Yd, alpha=2.0, beta=0.5)                          # beta folded into the diagonal, no W
    assert np.abs(Yd.cpu().numpy() - (2.0 * ref + 0.5 * X)).max() <= 1e-13 * np.abs(ref).max()
    big = torch.from_numpy(rng.normal(size=(A.nrows, ncols + 40))).to(_dev())      # a panel inside a wider block vector
    out = torch.full_like(big, float("nan"))
    A.spmm_pattern(big[:, 8:8 + ncols], out[:, 16:16 + ncols])
    refb = S @ big[:, 8:8 + ncols].cpu().numpy()
    assert np.abs(out[:, 16:16 + ncols].cpu().numpy() - refb).max() <= 1e-13 * np.abs(refb).max()
    assert bool(torch.isnan(out[:, :16]).all()) and bool(torch.isnan(out[:, 16 + ncols:]).all())


@pytest.mark.parametrize("degree", [1, 2, 9, 40])
def test_cheb_filter_mma_pattern_matches_gather(degree):
    g = load_golden("sphere_n2000_k50")
    A, _ = _bsr_from_golden(g, "L")
    V0 = torch.from_numpy(np.random.default_rng(2).normal(size=(A.nrows, 192))).to(_dev())
    outs = []
    for on in (False, True):
        A.enable_mma_pattern(on)
        V = V0.clone()
        w0 = torch.empty((A.nrows, 64), dtype=torch.float64, device=_dev()); w1 = torch.empty_like(w0)
        if on:
            assert A._mma_pattern_ok(64, V[:, 64:128], w0, w1)
        A.cheb_filter(V[:, 64:128], w0, w1, 64, degree, 0.0, 3.0, 40.0)
        outs.append(V.cpu().numpy())
    assert np.abs(outs[1][:, :64] - V0[:, :64].cpu().numpy()).max() == 0.0
    assert np.abs(outs[1][:, 128:] - V0[:, 128:].cpu().numpy()).max() == 0.0
    assert np.abs(outs[0] - outs[1]).max() <= 1e-11 * np.abs(outs[0]).max()


@pytest.mark.parametrize("cplx", [False, True])
def test_projected_problem_device_path_matches_host_lapack(cplx):
    """The m x m projected problems of the eigensolvers (CholeskyQR factor, Rayleigh-Ritz) run on the GPU for m >= 192
    (eigensolver._cholqr_factor / _rayleigh_ritz_small): same results as the host LAPACK path, incl. the shifted Cholesky."""
    from rvgp_b200 import eigensolver as E
    rng = np.random.default_rng(0)
    m = 256
    B = rng.normal(size=(m, 3 * m)) + (1j * rng.normal(size=(m, 3 * m)) if cplx else 0)
    G = B @ B.conj().T / (3 * m)
    A = rng.normal(size=(m, m)) + (1j * rng.normal(size=(m, m)) if cplx else 0)
    H = A + A.conj().T
    dev = _dev()
    assert E._on_device(m, dev) and not E._on_device(64, dev)
    Rinv_d, sh_d = E._cholqr_factor(G, dev)
    Rinv_h, sh_h = E._cholqr_factor(G, None)
    assert not sh_d and not sh_h and np.abs(Rinv_d - Rinv_h).max() <= 1e-11 * np.abs(Rinv_h).max()
    assert np.abs(np.triu(Rinv_d) - Rinv_d).max() == 0.0 or np.abs(np.tril(Rinv_d, -1)).max() < 1e-14
    th_d, C_d = E._rayleigh_ritz_small(G, H, dev)
    th_h, C_h = E._rayleigh_ritz_small(G, H, None)
    np.testing.assert_allclose(th_d, th_h, rtol=1e-11, atol=1e-11)
    # eigenvectors up to phase: C^H G C = I and H C = G C theta
    for C, th in ((C_d, th_d), (C_h, th_h)):
        assert np.abs(C.conj().T @ G @ C - np.eye(m)).max() < 1e-10
        assert np.abs(H @ C - (G @ C) * th).max() < 1e-9 * np.abs(H).max()
    Gs = G.copy()
    Gs[1] = Gs[0]; Gs[:, 1] = Gs[:, 0]                               # exactly singular -> shifted factor, flagged
    _, sh = E._cholqr_factor(Gs - 1e-9 * np.eye(m), dev)
    assert sh


@pytest.mark.parametrize("kind,n,k", [("flat3torus", 6000, 40), ("sphere", 12000, 64), ("moebius", 8000, 48)])
def test_krylov_solver_equals_chfsi_on_other_geometries(kind, n, k, monkeypatch):
    """The filtered block Lanczos solver (krylov.py) against the subspace-iteration solver on geometries other than the benchmark
    torus: a 3-manifold in R^6 (d = 3 blocks: REAL block mode on the gather kernel), a sphere (paired complex mode, exact
    multiplicities 2l+1) and a Moebius strip (non-orientable: no paired mode).  Same eigenvalues to 1e-9 relative, same invariant
    subspaces per cluster, true residuals under the tolerance."""
    import RVGP
    from tests.workloads import make_cloud
    from tests.conftest import eigen_clusters, subspace_angle_max
    X = make_cloud(kind, n, 0)
    out = {}
    for solver in ("chfsi", "krylov"):
        monkeypatch.setenv("RVGP_EIGSOLVER", solver)
        d = RVGP.create_data_object(X, n_eigenpairs=k, verbose=False)
        assert ("Lanczos" in str(d.stats["eig_Lc"].get("solver"))) == (solver == "krylov")
        for name in ("eig_L", "eig_Lc"):
            st = d.stats[name]
            assert st["converged"] and st["residual_max"] <= st["tol_abs"], (solver, name, st["residual_max"])
        out[solver] = (d.evals_L.copy(), d.evals_Lc.copy(), d.evecs_L.copy(), d.evecs_Lc.copy(), d.stats["paired"], d.dim_man)
    a, b = out["chfsi"], out["krylov"]
    assert a[4] == b[4] and a[5] == b[5]
    if kind == "moebius":
        assert not a[4]
    for ea, eb in ((a[0], b[0]), (a[1], b[1])):
        assert np.abs(ea - eb).max() <= 1e-9 * np.abs(ea).max()
    for ev, Ua, Ub in ((a[0], a[2], b[2]), (a[1], a[3], b[3])):
        for s in eigen_clusters(ev, rtol=1e-6)[:-1]:
            assert subspace_angle_max(Ua[:, s], Ub[:, s]) < 1e-6, (kind, s)


def test_krylov_thick_restart_on_device():
    """Thick restarts of the filtered block Lanczos solver on the GPU (a basis capacity far below what the run needs):
    same eigenvalues as the golden ARPACK output of the unmodified reference."""
    from rvgp_b200.krylov import krylov_eigenpairs
    g = load_golden("sphere_n2000_k50")
    A, S = _bsr_from_golden(g, "L")
    hi = 2.0 * (np.diff(g["L_indptr"]).max() - 1)
    ref = g["evals_L"]
    st = {}
    ev, U = krylov_eigenpairs(A, 40, hi, cut=1.3 * ref[49], lam_k=ref[39], block=16, cap_cols=144, stats=st)
    assert st["restarts"] >= 1 and st["converged"]
    np.testing.assert_allclose(ev.cpu().numpy(), ref[:40], rtol=1e-8, atol=1e-10)
    Uh = U.cpu().numpy()
    assert np.abs(S @ Uh - Uh * ev.cpu().numpy()).max() <= 1e-11 * hi
