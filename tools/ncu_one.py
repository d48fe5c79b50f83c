"""Run a few launches of ONE SpMM configuration (for ncu). usage: ncu_one.py torus 1000000 Lc|L b lpr"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._build import build
from rvgp_b200._cabi import get_handle
kind, n, which, b, lpr = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
A, L, _ = build(kind, n)
M = A if which == "Lc" else L
h = get_handle(0); h.set_option("spmm_lpr", lpr)
X = torch.randn((M.nrows, b), dtype=torch.float64, device="cuda"); W = torch.randn_like(X); Y = torch.empty_like(X)
for _ in range(6):
    M.spmm(X, Y, alpha=0.7, beta=-0.2, gamma=0.1, W=W)
torch.cuda.synchronize()
