"""K9 at C4 size (1M-point torus): CUDA-event timing of the fused Chebyshev step (Y = a A X + b X + c W, three rotating
buffers) for
  * the d = 2 connection Laplacian on the FP64-MMA native kernel, 64 columns (schedule variants / cache policies), and
  * the scalar Laplacian: gather kernel vs the MMA pattern kernel (L (x) I_2, spmm_mma.cu AMODE 2) at 64 and 128 columns.
usage: python tools/profile_k9.py [n] [what: all|L|Lc]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rvgp_b200 import geometry as geo
from rvgp_b200._cabi import get_handle
from rvgp_b200.eigensolver import BsrMatrix
from tests.workloads import make_cloud

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
what = sys.argv[2] if len(sys.argv) > 2 else "all"
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6453.4
dev = torch.device("cuda", 0)
h = get_handle(0)
Xd = torch.from_numpy(make_cloud("torus", n, 0)).to(dev)
g = geo.manifold_graph(Xd, n_neighbors=10)
seq, _ = geo.geodesic_neighbourhoods_device(g.indptr, g.indices, 15, g.max_row)
T, S = geo.tangent_frames_device(Xd, seq, 3)
gauges = geo.slice_frames_device(T, 2)
order, inv = geo.morton_order_device(Xd)
ip, ix = geo.csr_permute_device(g.indptr, g.indices, order, inv)
gp = geo.gather_rows_device(gauges.reshape(n, -1), order).reshape(n, 3, 2)
A = BsrMatrix(n, 2, ip, ix, geo.connections_device(gp, ip, ix))
L = BsrMatrix(n, 1, ip, ix, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
kw = dict(alpha=0.03, beta=-0.2, gamma=0.1)


def time_rec(step, nl=24):
    step(4)
    torch.cuda.synchronize()
    e0.record(); step(nl); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nl


if what in ("all", "Lc"):
    b = 64
    by = A.spmm_bytes(b, True)
    A.enable_mma()
    X = torch.randn((A.nrows, b), dtype=torch.float64, device=dev)
    bufs = [A.to_native(X), A.to_native(torch.randn_like(X)), A.to_native(torch.randn_like(X))]
    for var, gpw, pd, pol in [(v, 0, p, q) for v in (1, 0, 3) for p in (1, 2) for q in (7, 3, 23)]:
        h.set_option("mma_variant", var); h.set_option("mma_gpw", gpw); h.set_option("mma_prefetch", pd); h.set_option("mma_stream_policy", pol)
        t = time_rec(lambda nl: [A.spmm_native(bufs[(i + 1) % 3], bufs[(i + 2) % 3], Wn=bufs[i % 3], **kw) for i in range(nl)])
        print("Lc 64 cols native v%d gpw%d prefetch%d policy%d: %.4f ms  frac %.3f" % (var, gpw, pd, pol, t, by / t / 1e6 / PEAK), flush=True)
    h.set_option("mma_variant", 1); h.set_option("mma_gpw", 0); h.set_option("mma_prefetch", 1); h.set_option("mma_stream_policy", 7)

if what in ("all", "L"):
    for b in (64, 128):
        by = L.spmm_bytes(b, True)
        bufs = [torch.randn((n, b), dtype=torch.float64, device=dev) for _ in range(3)]
        L.enable_mma_pattern(False)
        t = time_rec(lambda nl: [L.spmm(bufs[(i + 1) % 3], bufs[(i + 2) % 3], W=bufs[i % 3], **kw) for i in range(nl)])
        print("L %d cols gather: %.4f ms  frac %.3f" % (b, t, by / t / 1e6 / PEAK), flush=True)
        ref = bufs[2].clone()
        L.spmm(bufs[0], ref, W=bufs[1], **kw)
        L.enable_mma_pattern(True)
        for v2 in ((0, 1, 2, 3, 4, 5, 6) if b == 64 else (0,)):
            for pd, pol in ((1, 7), (2, 7), (1, 3)):
                h.set_option("mma_variant_n2", v2); h.set_option("mma_prefetch", pd); h.set_option("mma_stream_policy", pol)
                out = torch.empty_like(ref)
                L.spmm_pattern(bufs[0], out, W=bufs[1], **kw)
                err = float((out - ref).abs().max() / ref.abs().max())
                t = time_rec(lambda nl: [L.spmm_pattern(bufs[(i + 1) % 3], bufs[(i + 2) % 3], W=bufs[i % 3], **kw) for i in range(nl)])
                print("L %d cols MMA pattern variant_n2=%d prefetch%d policy%d: %.4f ms  frac %.3f  relerr %.1e" %
                      (b, v2, pd, pol, t, by / t / 1e6 / PEAK, err), flush=True)
        h.set_option("mma_variant_n2", 0); h.set_option("mma_prefetch", 1); h.set_option("mma_stream_policy", 7)
