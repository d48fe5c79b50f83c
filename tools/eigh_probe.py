"""Where should the m x m projected eigenproblem of the Krylov eigensolver run?  Times host LAPACK (threads swept) against
torch.linalg.eigh on the GPU for the two shapes of config C4 (real 1280, complex 960)."""
import time
import numpy as np
import scipy.linalg as sl
import torch
from threadpoolctl import threadpool_limits
rng = np.random.default_rng(0)
for n, cplx in ((1280, False), (960, True), (1024, False), (768, True)):
    A = rng.normal(size=(n, n)) + (1j * rng.normal(size=(n, n)) if cplx else 0)
    A = A + A.conj().T
    for th in (1, 4, 8, 16):
        with threadpool_limits(limits=th, user_api="blas"):
            sl.eigh(A, driver="evd", check_finite=False)
            t = time.perf_counter(); sl.eigh(A, driver="evd", check_finite=False); t1 = time.perf_counter() - t
            t = time.perf_counter(); sl.eigvalsh(A, check_finite=False); t2 = time.perf_counter() - t
        print("n %d %s host threads %2d: eigh %.3f s  eigvalsh %.3f s" % (n, "complex" if cplx else "real", th, t1, t2), flush=True)
    Ad = torch.from_numpy(A).cuda()
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); w, v = torch.linalg.eigh(Ad); torch.cuda.synchronize(); t1 = time.perf_counter() - t
    print("n %d %s GPU torch.linalg.eigh: %.3f s" % (n, "complex" if cplx else "real", t1), flush=True)
