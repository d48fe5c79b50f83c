"""Isolated timing of the two skinny FP64 GEMM shapes of the Krylov eigensolver's orthogonalisation at C4 size:
Gram  C (cur x 64) = V^T W   (K = N rows, split-K)   and   apply  W (N x 64) -= V (N x cur) C.
usage: python tools/profile_krylov_gemm.py [N] [cur]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rvgp_b200._cabi import get_handle, I64
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cur = int(sys.argv[2]) if len(sys.argv) > 2 else 768
cap, b = 1216, 64
dev = torch.device("cuda", 0)
h = get_handle(0)
V = torch.randn((N, cap), dtype=torch.float64, device=dev)
W = torch.randn((N, b), dtype=torch.float64, device=dev)
C = torch.empty((cap, b), dtype=torch.float64, device=dev)
ws = torch.empty(3 * h.sm_count * 128 * 64 + cap * b, dtype=torch.float64, device=dev)
import math
def gram(split):
    h.call("rvgp_dgemm_f64", int(cur), int(b), I64(N), 1.0, V, I64(cap), 0, W, I64(b), 0, None, C, I64(b), int(split), ws)
def apply():
    h.call("rvgp_dgemm_acc_f64", int(N), int(b), I64(cur), -1.0, V, I64(cap), 1, C, I64(b), 0, 1.0, W, I64(b))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fl = 2.0 * N * cur * b
tiles = math.ceil(cur / 128)
for split in (16, 32, 49, 74, 98, 148, 222):
    if split * cur * b > ws.numel(): continue
    t = timeit(lambda: gram(split))
    print("gram  N=%d cur=%d split=%3d (CTAs %4d): %.3f ms  %.1f TFLOP/s" % (N, cur, split, split * tiles, t, fl / t / 1e9))
t = timeit(apply)
print("apply N=%d cur=%d: %.3f ms  %.1f TFLOP/s" % (N, cur, t, fl / t / 1e9))
