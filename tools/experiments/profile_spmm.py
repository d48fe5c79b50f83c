"""Build the C4 (or given) connection Laplacian through the product pipeline, then time / expose the fused SpMM.
usage: python tools/profile_spmm.py torus 1000000 [reps]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.workloads import make_cloud
from rvgp_b200 import geometry as geo
from rvgp_b200.eigensolver import BsrMatrix

def build(kind, n, nb=10):
    dev = torch.device("cuda", 0)
    Xd = torch.from_numpy(make_cloud(kind, n, 0)).to(dev)
    g = geo.manifold_graph(Xd, n_neighbors=nb)
    seq, _ = geo.geodesic_neighbourhoods_device(g.indptr, g.indices, int(nb * 1.5), g.max_row)
    T, S = geo.tangent_frames_device(Xd, seq, Xd.shape[1])
    gauges = geo.slice_frames_device(T, 2)
    order, inv = geo.morton_order_device(Xd)
    ip, ix = geo.csr_permute_device(g.indptr, g.indices, order, inv)
    gp = geo.gather_rows_device(gauges.reshape(n, -1), order).reshape(n, Xd.shape[1], 2)
    vals = geo.connections_device(gp, ip, ix)
    return BsrMatrix(n, 2, ip, ix, vals), BsrMatrix(n, 1, ip, ix, None), (g.indptr, g.indices, gauges)

def main():
    kind, n = sys.argv[1], int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    A, L, raw = build(kind, n)
    dev = A.indptr.device
    out = {}
    for name, M in (("Lc", A), ("L", L)):
        for b in (8, 16, 32, 64):
            X = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
            W = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
            Y = torch.empty_like(X)
            for fused in (False, True):
                kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W) if fused else {}
                for _ in range(3):
                    M.spmm(X, Y, **kw)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    M.spmm(X, Y, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                by = M.spmm_bytes(b, fused=fused)
                out["%s_b%d_%s" % (name, b, "fused" if fused else "plain")] = dict(ms=round(ms, 4), GBs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / 6534.5, 3))
    # shared-memory-staged variant
    for name, M in (("Lc", A), ("L", L)):
        for TR in (8, 16, 32):
            plan = M.build_plan(TR)
            if plan is None:
                out["%s_TR%d" % (name, TR)] = "plan invalid"; continue
            for b in (32, 64):
                smem = M.tiled_smem_bytes(plan, b)
                if smem > 200 * 1024:
                    continue
                X = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
                W = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
                Y = torch.empty_like(X); Y2 = torch.empty_like(X)
                kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W)
                M.spmm(X, Y2, **kw)
                for _ in range(3):
                    M.spmm_tiled(plan, X, Y, **kw)
                torch.cuda.synchronize()
                err = float((Y - Y2).abs().max())
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    M.spmm_tiled(plan, X, Y, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                by = M.spmm_bytes(b, fused=True)
                out["%s_TR%d_b%d_tiled_fused" % (name, TR, b)] = dict(ms=round(ms, 4), GBs=round(by / ms / 1e6, 1), frac=round(by / ms / 1e6 / 6534.5, 3), umax=plan["umax"], usoft=plan["usoft"], umean=round(plan["umean"],1), heavy=round(plan["heavy_frac"],4), nemax=plan["nemax"], smem=smem, maxerr=err)
    print(json.dumps(out, indent=1))

if __name__ == "__main__":
    main()
