"""CPU emulation (NumPy / SciPy) of rvgp_b200.eigensolver.smallest_eigenpairs for STUDYING the outer-iteration policy
(first-sweep degree, cond_max, buffer size, degree margin) without a GPU.  Same algorithm and the same `_next_degrees`; it
counts column-degrees (the filter cost) and outer iterations (the dense cost).  Not a product path, not a test oracle.

usage: python tools/eig_emulate.py [n] [k]"""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, scipy.linalg as sl
from rvgp_b200.eigensolver import _next_degrees, _chol_upper_shifted, _tri_inv_upper
from tests.workloads import make_cloud
from oracle import rvgp_oracle as O


def cheb(A, V, deg, lo, cut, hi, f32=False):
    if f32:                                   # FP32 panels AND FP32 matrix values: what a half-width filter kernel would do
        A = A.astype(np.float32); V = V.astype(np.float32)
    e = 0.5 * (hi - cut); c = 0.5 * (hi + cut); s1 = e / (lo - c); tau = 2 / s1; sig = s1
    Y = (A @ V - c * V) * (s1 / e); X = V
    for _ in range(2, deg + 1):
        sn = 1 / (tau - sig)
        Yn = (A @ Y - c * Y) * (2 * sn / e) - sig * sn * X
        X, Y = Y, Yn; sig = sn
    return Y.astype(np.float64)


def solve(A, k, hi_g, nex=None, panel=64, deg0=20, cond_max=1e6, margin=1.05, decade=10.0, tol=1e-12, hi=None, seed=0, V0=None,
          f32_until=0.0):
    """f32_until: filter in FP32 while the largest wanted residual is above f32_until * hi_g (0 = always FP64)."""
    N = A.shape[0]
    nex = max(16, int(math.ceil(0.2 * k))) if nex is None else nex
    m = ((k + nex + panel - 1) // panel) * panel
    hi = hi or hi_g
    tol_abs = tol * hi_g
    V = np.random.default_rng(seed).uniform(-1, 1, (N, m)) if V0 is None else V0.copy()
    deg = np.full(m, deg0); a_cut = 0.3 * hi; coldeg = 0; passes = 0; f32deg = 0; worst = np.inf
    for it in range(80):
        for p0 in range(0, m, panel):
            dg = int(deg[p0:p0 + panel].max())
            if dg > 0:
                use32 = worst > f32_until * hi_g and f32_until > 0
                V[:, p0:p0 + panel] = cheb(A, V[:, p0:p0 + panel], dg, 0.0, a_cut, hi, f32=use32); coldeg += dg * panel
                f32deg += dg * panel if use32 else 0
        for _ in range(4):
            V /= np.sqrt((V * V).sum(0)); R, sh = _chol_upper_shifted(V.T @ V); V = V @ _tri_inv_upper(R); passes += 1
            if not sh: break
        W = A @ V
        G = V.T @ V; H = V.T @ W
        R2inv = _tri_inv_upper(np.linalg.cholesky(G).T)
        Hm = R2inv.T @ H @ R2inv; th, Y = np.linalg.eigh(0.5 * (Hm + Hm.T))
        V = V @ (R2inv @ Y)
        res = np.sqrt((((A @ V) - V * th) ** 2).sum(0))
        nconv = int((res[:k] <= tol_abs).sum()); worst = float(res[:k].max())
        a_cut = float(th[-1])
        if nconv == k:
            break
        a_cut = max(a_cut, 1e-12 * hi + th[k - 1] * (1 + 1e-9))
        # the product's rule with two knobs exposed
        e = 0.5 * (hi - a_cut); c = 0.5 * (hi + a_cut)
        g = np.arccosh(np.maximum((c - th) / e, 1.0)); g0 = math.acosh(max(c / e, 1.0))
        with np.errstate(divide="ignore"):
            need = np.where(res > tol_abs, np.maximum(np.log(np.maximum(decade * res / tol_abs, 1.0)) / np.maximum(g, 1e-12), 8.0), 0.0)
        cap = math.log(cond_max) / np.maximum(g0 - g, 1e-12)
        d = np.ceil(np.minimum(need * margin + 1.0, cap)); d[res <= tol_abs] = 0
        d[k:] = np.minimum(d[k:], d[:k].max())
        deg = np.minimum(d, 6000).astype(np.int64)
    return dict(f32deg=f32deg, outer=it + 1, coldeg=coldeg, per_col=coldeg / m, m=m, passes=passes, maxres=float(res[:k].max()), evals=th[:k])


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    X = make_cloud("torus", n, 0)
    ip, ix = O.symmetrize_csr(O.knn_exact(X, 10))
    rows = np.repeat(np.arange(n), np.diff(ip))
    deg = (np.diff(ip) - 1).astype(float)
    A = sp.csr_matrix((np.where(rows == ix, deg[rows], -1.0), ix, ip), shape=(n, n))
    hi_g = 2.0 * (np.diff(ip).max() - 1)
    hi = 1.01 * float(sp.linalg.eigsh(A, 1, which="LA", return_eigenvectors=False)[0])
    base = None
    variants = (("default", {}), ("deg0=40", dict(deg0=40)), ("deg0=80", dict(deg0=80)), ("cond 1e8", dict(cond_max=1e8)),
                ("cond 1e10", dict(cond_max=1e10)), ("decade 2", dict(decade=2.0)), ("margin 1.0", dict(margin=1.0)),
                ("nex 0.4k", dict(nex=max(32, int(0.4 * k)))), ("nex 0.1k", dict(nex=max(16, int(0.1 * k)))),
                ("deg0=40 cond 1e8 decade 2", dict(deg0=40, cond_max=1e8, decade=2.0)),
                ("fp32 until 1e-5", dict(f32_until=1e-5)), ("fp32 until 1e-6", dict(f32_until=1e-6)), ("fp32 until 1e-7", dict(f32_until=1e-7)))
    only = os.environ.get("EMU_VARIANTS")
    if only:
        variants = tuple(v for v in variants if v[0] in only.split(";"))
    for name, kw in variants:
        t0 = time.perf_counter()
        r = solve(A, k, hi_g, hi=hi, **kw)
        if base is None:
            base = r
        print("%-28s outer %2d  cholqr %2d  m %4d  col-degrees %9d (%.2fx, %2.0f%% in fp32)  per column %6.0f  maxres %.1e  evals dev %.1e  %.0f s" %
              (name, r["outer"], r["passes"], r["m"], r["coldeg"], r["coldeg"] / base["coldeg"], 100.0 * r["f32deg"] / r["coldeg"], r["per_col"], r["maxres"],
               np.abs(r["evals"] - base["evals"]).max(), time.perf_counter() - t0), flush=True)
