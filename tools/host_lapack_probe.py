"""Host-side cost of the projected problems of the block eigensolver (rvgp_b200/eigensolver.py) vs BLAS thread count."""
import os, time, sys
import numpy as np, scipy.linalg
from threadpoolctl import threadpool_limits, threadpool_info
print("cpu_count", os.cpu_count(), [(d["internal_api"], d["num_threads"]) for d in threadpool_info()])
rng = np.random.default_rng(0)
def spd(m, cplx):
    A = rng.normal(size=(m, 2 * m)) + (1j * rng.normal(size=(m, 2 * m)) if cplx else 0)
    return A @ A.conj().T / m + np.eye(m)
def herm(m, cplx):
    A = rng.normal(size=(m, m)) + (1j * rng.normal(size=(m, m)) if cplx else 0)
    return A + A.conj().T
def one_outer(m, cplx):
    G, H = spd(m, cplx), herm(m, cplx)
    t0 = time.perf_counter()
    R = np.linalg.cholesky(G).conj().T
    Rinv = scipy.linalg.solve_triangular(R, np.eye(m), lower=False)
    R2 = np.linalg.cholesky(G).conj().T
    R2inv = scipy.linalg.solve_triangular(R2, np.eye(m), lower=False)
    Hm = R2inv.conj().T @ H @ R2inv
    Hm = 0.5 * (Hm + Hm.conj().T)
    t1 = time.perf_counter()
    th, Y = np.linalg.eigh(Hm)
    t2 = time.perf_counter()
    C = R2inv @ Y
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for label, m, cplx in (("real 640 (L)", 640, False), ("complex 320 (Lc paired)", 320, True)):
    for nt in (None, 1, 2, 4, 8, 16, 32):
        ctx = threadpool_limits(limits=nt, user_api="blas") if nt else threadpool_limits(limits=None)
        with ctx:
            one_outer(m, cplx)
            ts = np.array([one_outer(m, cplx) for _ in range(3)]).min(0)
        print("%-26s threads %-7s chol+inv+proj %.1f ms  eigh %.1f ms  back %.1f ms  total %.1f ms" %
              (label, nt or "default", ts[0] * 1e3, ts[1] * 1e3, ts[2] * 1e3, ts.sum() * 1e3))
