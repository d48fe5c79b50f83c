"""Run a few fused launches of the shipped FP64-MMA SpMM (node-contiguous panels) for ncu.
usage: ncu_mma.py torus 1000000 [variant]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.profile_spmm import build
from rvgp_b200._cabi import get_handle
kind, n = sys.argv[1], int(sys.argv[2])
A, _, _ = build(kind, n)
A.enable_mma()
if len(sys.argv) > 3:
    get_handle(0).set_option("mma_variant", int(sys.argv[3]))
b = 64
bufs = [torch.randn((A.nbrows, 2 * b), dtype=torch.float64, device="cuda") for _ in range(3)]
for i in range(6):
    A.spmm_native(bufs[(i + 1) % 3], bufs[(i + 2) % 3], alpha=0.03, beta=-0.2, gamma=0.1, Wn=bufs[i % 3])
torch.cuda.synchronize()
