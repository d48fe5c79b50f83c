"""Per-kernel roofline table at config-C4 size (n = 1M torus): CUDA-event time, algorithmic bytes / flops (DESIGN.md
section 3), achieved rate and fraction of the measured peak (HBM 6534.5 GB/s from MEASURED_PEAKS.json, FP64 37.0 TFLOP/s
from tools/fp64_peak.cu)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.workloads import make_cloud
from rvgp_b200 import geometry as geo
from rvgp_b200.eigensolver import BsrMatrix, _Dense
from rvgp_b200._cabi import get_handle, I64

HBM, FP64 = 6534.5e9, 37.0e12
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = torch.device("cuda", 0)
h = get_handle(0)
rows = []

def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, out

def add(name, t, nbytes=None, flops=None, note=""):
    r = dict(kernel=name, ms=round(t * 1e3, 3))
    if nbytes: r.update(GBs=round(nbytes / t / 1e9, 1), frac_hbm=round(nbytes / t / HBM, 3))
    if flops: r.update(TFLOPs=round(flops / t / 1e12, 2), frac_fp64=round(flops / t / FP64, 3))
    r["note"] = note
    rows.append(r); print(r, flush=True)

X = make_cloud("torus", n, 0); D = 3; nb = 10
Xd = torch.from_numpy(X).to(dev)
t, knn = timed(lambda: geo.knn_device(Xd, nb, method="grid"))
add("K2 knn_grid (exact, D=3)", t, nbytes=8 * n * D + 4 * n * nb, note="algorithmic minimum traffic 8nD + 4 n nb; candidate pruning by cell grid")
nq = min(n, 100000)
t, _ = timed(lambda: geo.knn_device(Xd, nb, q_begin=0, q_count=nq, method="brute"), reps=1)
add("K2 knn brute (%d queries x %d candidates)" % (nq, n), t, flops=(3 * D - 1) * float(nq) * n, note="non-fused FP64 ops, 1 flop each (peak for non-FMA ops is half of 37)")
t, (ip, ix) = timed(lambda: geo.knn_to_csr_device(knn))
add("K3 knn_to_csr (radix sort + unique)", t, nbytes=(2 * n * nb + n) * 8 * 6, note="64-bit keys, ~3 read+write passes")
K = 15
maxrow = int((ip[1:] - ip[:-1]).max().item())
t, (seq, cnt) = timed(lambda: geo.geodesic_neighbourhoods_device(ip, ix, K, maxrow))
add("K4 geodesic heap emulation", t, note="latency / integer bound; %.0f ns per source" % (t / n * 1e9))
t, (T, S) = timed(lambda: geo.tangent_frames_device(Xd, seq, D))
add("K5 gauges (one-warp Jacobi SVD)", t, nbytes=8 * n * (K + 1) * D + 8 * n * D * (D + 1) + 4 * n * (K + 1))
t, ve = timed(lambda: geo.explained_variance_device(S))
add("K6 explained variance", t, nbytes=8 * n * D * 4)
gauges = geo.slice_frames_device(T, 2)
order, inv = geo.morton_order_device(Xd)
pip, pix = geo.csr_permute_device(ip, ix, order, inv)
gp = geo.gather_rows_device(gauges.reshape(n, -1), order).reshape(n, D, 2)
nnzb = pix.numel()
t, vals = timed(lambda: geo.connections_device(gp, pip, pix))
add("K7/K8 connections + Lc assembly", t, nbytes=2 * 8 * nnzb * D * 2 + 8 * nnzb * 4 + 4 * nnzb)
A = BsrMatrix(n, 2, pip, pix, vals); L = BsrMatrix(n, 1, pip, pix, None)
for M, b, nm in ((A, 64, "Lc d=2"), (L, 64, "L pattern"), (L, 128, "L pattern")):
    Xv = torch.randn((M.nrows, b), dtype=torch.float64, device=dev); W = torch.randn_like(Xv); Y = torch.empty_like(Xv)
    t, _ = timed(lambda: M.spmm(Xv, Y, alpha=0.7, beta=-0.2, gamma=0.1, W=W), reps=10, warm=3)
    add("K9 fused SpMM %s, %d cols" % (nm, b), t, nbytes=M.spmm_bytes(b, True), flops=2.0 * M.nnzb * M.d * M.d * b)
    t, _ = timed(lambda: M.spmm(Xv, Y), reps=10, warm=3)
    add("K9 plain SpMM %s, %d cols" % (nm, b), t, nbytes=M.spmm_bytes(b, False))
    del Xv, W, Y
N, m = 2 * n, 640
V = torch.randn((N, m), dtype=torch.float64, device=dev); Wb = torch.empty_like(V)
G = torch.empty((m, m), dtype=torch.float64, device=dev); C = torch.randn((m, m), dtype=torch.float64, device=dev)
dn = _Dense(h, N, m, dev)
t, _ = timed(lambda: dn.gram(V, V, G, sym=True)); add("K10 Gram V^T V (lower tiles), %dx%d" % (N, m), t, flops=1.0 * N * m * (m + 64), nbytes=8.0 * N * m)
t, _ = timed(lambda: dn.apply(V, C, Wb)); add("K10 apply V C, %dx%d" % (N, m), t, flops=2.0 * N * m * m, nbytes=16.0 * N * m)
t, _ = timed(lambda: dn.coldot(V, V)); add("K10 column norms", t, nbytes=8.0 * N * m)
del V, Wb
U = torch.randn((N, 500), dtype=torch.float64, device=dev)
t, _ = timed(lambda: geo.frame_apply_device(gp, U.reshape(n, 2, 500), 1)); add("K11 eigenvector lift (n,d,k)->(n,D,k)", t, nbytes=8.0 * n * 500 * (2 + 3) + 8 * n * 6)
del U
from rvgp_b200.fps import furthest_point_sampling_device
ns = 200000
Xs = Xd[:ns].contiguous()
t, _ = timed(lambda: furthest_point_sampling_device(Xs, N=501), reps=1)
add("K1 FPS %d points, 500 samples" % ns, t, nbytes=500 * 8.0 * ns * (D + 2), note="%.1f us per sample (one grid barrier each)" % (t / 500 * 1e6))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/kernel_rooflines.json", "w"), indent=1)
