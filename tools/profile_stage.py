import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._build import build
from rvgp_b200._cabi import get_handle
A, L, _ = build("torus", 1000000)
h = get_handle(0)
dev = A.indptr.device
def bench(M, X, W, Y, reps=10):
    kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W)
    for _ in range(3): M.spmm(X, Y, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): M.spmm(X, Y, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for b in (32, 64):
    X = torch.randn((A.nrows, b), dtype=torch.float64, device=dev); W = torch.randn_like(X)
    Y0 = torch.empty_like(X); Y1 = torch.empty_like(X)
    by = A.spmm_bytes(b, True)
    for rot in (False,):
        A.d_code = A.d
        if rot: A.compress_rot2()
        for stage in (0,):
            pass
            for lpr in (0,):
                h.set_option("spmm_lpr", lpr)
                t = bench(A, X, W, Y1 if (stage or rot) else Y0)
                print("Lc b=%d rot2=%s remap=%d lpr=%d: %.4f ms frac %.3f  maxdiff %.1e" % (b, rot, stage, lpr, t, by / t / 1e6 / 6534.5, float((Y0 - Y1).abs().max()) if (stage or rot) else 0.0))
