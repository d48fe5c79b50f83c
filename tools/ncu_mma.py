"""Run a few fused launches of the shipped FP64-MMA SpMM (node-contiguous panels) for ncu.
usage: [RVGP_NCU_B=32] ncu_mma.py torus 1000000 [variant]   (RVGP_NCU_B: complex columns of the Lc panel, default 64)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._build import build
from rvgp_b200._cabi import get_handle
kind, n = sys.argv[1], int(sys.argv[2])
A, L, _ = build(kind, n)
A.enable_mma()
if len(sys.argv) > 3:
    get_handle(0).set_option("mma_variant", int(sys.argv[3]))
b = int(os.environ.get("RVGP_NCU_B", "64"))
bufs = [torch.randn((A.nbrows, 2 * b), dtype=torch.float64, device="cuda") for _ in range(3)]
for i in range(6):
    A.spmm_native(bufs[(i + 1) % 3], bufs[(i + 2) % 3], alpha=0.03, beta=-0.2, gamma=0.1, Wn=bufs[i % 3])
torch.cuda.synchronize()
# the scalar Laplacian on the same kernel in pattern mode (AMODE 2): row-major 64-column panels
L.enable_mma_pattern()
lb = [torch.randn((L.nbrows, 64), dtype=torch.float64, device="cuda") for _ in range(3)]
for i in range(6):
    L.spmm_pattern(lb[(i + 1) % 3], lb[(i + 2) % 3], alpha=0.03, beta=-0.2, gamma=0.1, W=lb[i % 3])
torch.cuda.synchronize()
