"""Run a few fused launches of the FP64-MMA SpMM (for ncu). usage: ncu_mma.py torus 1000000 Lc|L b"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.profile_spmm import build
kind, n, which, b = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
A, L, _ = build(kind, n)
M = A if which == "Lc" else L
M.enable_mma()
V = torch.randn((M.nrows, 640), dtype=torch.float64, device="cuda")
X = V[:, :b]; W = torch.randn((M.nrows, b), dtype=torch.float64, device="cuda"); Y = torch.empty_like(W)
for _ in range(6):
    M.spmm(X, Y, alpha=0.7, beta=-0.2, gamma=0.1, W=W)
torch.cuda.synchronize()

if which == "Lc" and os.environ.get("RVGP_NCU_NATIVE"):
    # RVGP_NCU_NATIVE = "variant,policy,rotc"
    from rvgp_b200._cabi import get_handle, I64
    var, pol, rotc = [int(v) for v in os.environ["RVGP_NCU_NATIVE"].split(",")]
    h = get_handle(0); mp = M.mma
    h.set_option("mma_variant", var); h.set_option("mma_stream_policy", pol)
    kc, af = mp["kcols"], mp["afrag"]
    if rotc:
        kc = kc.clone(); af = torch.empty(mp["ksteps"] * 16, dtype=torch.float64, device="cuda")
        bad = torch.zeros(1, dtype=torch.int32, device="cuda")
        h.call("rvgp_bsr_mma_rotc", I64(mp["ksteps"]), mp["afrag"], kc, af, bad, 1e-12)
    bufs = [torch.randn((M.nbrows, 2 * b), dtype=torch.float64, device="cuda") for _ in range(3)]
    for i in range(6):
        h.call("rvgp_bsr_spmm_mma_native_f64", M.nbrows, mp["kptr"], kc, af, int(rotc), bufs[(i + 1) % 3], I64(2 * b), bufs[i % 3], I64(2 * b),
               bufs[(i + 2) % 3], I64(2 * b), int(b), 0.03, -0.2, 0.1, 0)
    torch.cuda.synchronize()
