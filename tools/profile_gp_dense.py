"""K13 / K14 FLOP/s: spectral Gram (X*S) X^T, blocked Cholesky, triangular solve with many right-hand sides."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from rvgp_b200._cabi import get_handle, I64
from rvgp_b200.eigensolver import _dgemm
from rvgp_b200.gp import _Chol
h = get_handle(0); dev = torch.device("cuda", 0)
out = {}
for M, k in ((8192, 200), (16384, 200), (32768, 200)):
    X = torch.randn((M, k), dtype=torch.float64, device=dev) / k ** 0.5
    S = torch.rand(k, dtype=torch.float64, device=dev) + 0.5
    K = torch.empty((M, M), dtype=torch.float64, device=dev)
    def gram():
        _dgemm(h, M, M, k, X, X.stride(0), 1, X, X.stride(0), 1, K, K.stride(0), scale_k=S)
        h.call("rvgp_add_diag_f64", K, I64(K.stride(0)), int(M), 1.0)
    gram(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); gram(); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    out["gram_M%d" % M] = dict(ms=round(t, 2), tflops=round(2.0 * M * M * k / t / 1e9, 2))
    e0.record(); ch = _Chol(h, K, M); e1.record(); torch.cuda.synchronize(); ch.check()
    t = e0.elapsed_time(e1)
    out["potrf_M%d" % M] = dict(ms=round(t, 2), tflops=round(M ** 3 / 3.0 / t / 1e9, 2))
    B = torch.randn((M, 512), dtype=torch.float64, device=dev)
    e0.record(); ch.solve(B, 0); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    out["trsm_M%d_nrhs512" % M] = dict(ms=round(t, 2), tflops=round(1.0 * M * M * 512 / t / 1e9, 2))
    del K, B, ch
print(json.dumps(out, indent=1))
