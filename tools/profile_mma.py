"""K9 v4 (FP64-MMA row-group SpMM on node-contiguous panels) against the gather kernel at C4 size: correctness + CUDA-event
timing of a 3-buffer recurrence (like the Chebyshev filter), swept over schedule variants / cache policies.
usage: python tools/profile_mma.py [torus 1000000] [quick]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._build import build
from rvgp_b200._cabi import get_handle

kind = sys.argv[1] if len(sys.argv) > 1 else "torus"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
quick = len(sys.argv) > 3
A, L, _ = build(kind, n)
h = get_handle(0)
dev = A.indptr.device
PEAK = 6534.5
b = 64
out = {}

X = torch.randn((A.nrows, 640), dtype=torch.float64, device=dev)[:, :b]
W = torch.randn((A.nrows, b), dtype=torch.float64, device=dev)
Y0 = torch.empty_like(W)
kw = dict(alpha=0.03, beta=-0.2, gamma=0.1, W=W)
for _ in range(3):
    A.spmm(X, Y0, **kw)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    A.spmm(X, Y0, **kw)
e1.record(); torch.cuda.synchronize()
t0 = e0.elapsed_time(e1) / 20
by = A.spmm_bytes(b, True)
out["gather"] = dict(ms=round(t0, 4), frac=round(by / t0 / 1e6 / PEAK, 3))
print("Lc b=%d fused gather: %.4f ms frac %.3f" % (b, t0, by / t0 / 1e6 / PEAK), flush=True)

mp = A.enable_mma()
print("plan: ksteps %d  reuse %.2f  fill %.2f  rotc %d" % (mp["ksteps"], mp["reuse"], mp["fill"], mp["rotc"]))
Xn, Wn = A.to_native(X), A.to_native(W)
bufs = [Xn.clone(), Wn.clone(), torch.empty_like(Xn)]


def run_rec(nl, rev=0):
    for i in range(nl):
        A.spmm_native(bufs[(i + 1) % 3], bufs[(i + 2) % 3], alpha=0.03, beta=-0.2, gamma=0.1, Wn=bufs[i % 3], reverse=(rev and (i & 1)))


cfgs = [(1, 0, 1, 7)] if quick else [(v, g, 1, p) for v in (1, 0, 2, 3, 5, 6) for g in (0, 4) for p in (3, 7)]
for var, gpw, pd, pol in cfgs:
    h.set_option("mma_variant", var); h.set_option("mma_gpw", gpw); h.set_option("mma_prefetch", pd); h.set_option("mma_stream_policy", pol)
    Yn = torch.empty_like(Xn)
    A.spmm_native(Xn, Yn, alpha=0.03, beta=-0.2, gamma=0.1, Wn=Wn)
    err = float((A.from_native(Yn) - Y0).abs().max() / Y0.abs().max())
    bufs[0].copy_(Xn); bufs[1].copy_(Wn)
    run_rec(4)
    torch.cuda.synchronize()
    e0.record(); run_rec(24); e1.record(); torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1) / 24
    key = "native_v%d_g%d_pd%d_pol%d" % (var, gpw, pd, pol)
    out[key] = dict(ms=round(t1, 4), frac=round(by / t1 / 1e6 / PEAK, 3), relerr=err)
    print("%s: %.4f ms frac %.3f relerr %.1e" % (key, t1, by / t1 / 1e6 / PEAK, err), flush=True)
h.set_option("mma_variant", 1); h.set_option("mma_gpw", 0); h.set_option("mma_prefetch", 1); h.set_option("mma_stream_policy", 7)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/profile_mma.json", "w"), indent=1)
