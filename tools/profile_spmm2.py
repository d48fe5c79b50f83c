"""LPR sweep of the gather SpMM. usage: python tools/profile_spmm2.py torus 1000000 [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools._build import build
from rvgp_b200._cabi import get_handle

def main():
    kind, n = sys.argv[1], int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    A, L, _ = build(kind, n)
    dev = A.indptr.device
    h = get_handle(0)
    out = {}
    for name, M in (("Lc", A), ("L", L)):
        for b in (64, 128):
            X = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
            W = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
            Y = torch.empty_like(X); Yref = torch.empty_like(X)
            kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W)
            h.set_option("spmm_v1", 0); h.set_option("spmm_lpr", 32); M.spmm(X, Yref, **kw)
            for code in (16, 32):
                v1 = 1 if code > 100 else 0
                lpr = code % 100
                h.set_option("spmm_v1", v1)
                h.set_option("spmm_lpr", lpr)
                for _ in range(3): M.spmm(X, Y, **kw)
                torch.cuda.synchronize()
                err = float((Y - Yref).abs().max())
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps): M.spmm(X, Y, **kw)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                by = M.spmm_bytes(b, fused=True)
                out["%s_b%d_lpr%d_%s" % (name, b, lpr, "v1" if v1 else "v2")] = "ms=%.4f GB/s=%.0f frac=%.3f err=%.1e" % (ms, by / ms / 1e6, by / ms / 1e6 / 6534.5, err)
    h.set_option("spmm_lpr", 0); h.set_option("spmm_v1", 0)
    for k, v in out.items(): print(k, v)

if __name__ == "__main__":
    main()
