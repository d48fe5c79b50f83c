"""Timing of the tall-skinny Gram / apply dgemm at C4-like shapes, DFMA register tile vs DMMA (mma.sync m8n8k4)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rvgp_b200._cabi import get_handle
from rvgp_b200.eigensolver import _Dense
N, m = 500000, 640
dev = torch.device("cuda", 0)
h = get_handle(0)
V = torch.randn((N, m), dtype=torch.float64, device=dev); W = torch.empty_like(V)
G = torch.empty((m, m), dtype=torch.float64, device=dev); G2 = torch.empty_like(G)
C = torch.randn((m, m), dtype=torch.float64, device=dev)
dn = _Dense(h, N, m, dev)
out = {}
for dm in (0, 1):
    h.set_option("dgemm_dmma", dm)
    for name, fn, flops in (("gram", lambda: dn.gram(V, V, G), 2.0 * N * m * m), ("apply", lambda: dn.apply(V, C, W), 2.0 * N * m * m)):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name + ("_dmma" if dm else "_dfma")] = dict(ms=round(ms, 3), tflops=round(flops / ms / 1e9, 2))
    if dm == 0:
        G2.copy_(G)
out["gram_dmma_vs_dfma_maxdiff"] = float((G - G2).abs().max())
h.set_option("dgemm_dmma", 0)
print(json.dumps(out))
