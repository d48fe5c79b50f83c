import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._build import build
from rvgp_b200._cabi import get_handle
A, L, _ = build("torus", 1000000)
h = get_handle(0)
dev = A.indptr.device
def bench(M, b, reps=10):
    X = torch.randn((M.nrows, b), dtype=torch.float64, device=dev); W = torch.randn_like(X); Y = torch.empty_like(X)
    kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W)
    for _ in range(3): M.spmm(X, Y, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): M.spmm(X, Y, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, Y
for b in (32, 64):
    for lpr in (0, 16, 32):
        h.set_option("spmm_lpr", lpr)
        A.d_code = 2
        t0, Y0 = bench(A, b)
        ok = A.compress_rot2()
        t1, Y1 = bench(A, b)
        print("b=%d lpr=%d plain %.4f ms  rot2(%s) %.4f ms  maxdiff %.2e  frac %.3f" % (b, lpr, t0, ok, t1, float((Y0 - Y1).abs().max()), A.spmm_bytes(b, True) / t1 / 1e6 / 6534.5))
