"""K9 v4 (FP64-MMA row-group SpMM) against the gather kernel at C4 size: correctness + CUDA-event timing.
usage: python tools/profile_mma.py [torus 1000000]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.profile_spmm import build
from rvgp_b200._cabi import get_handle, I64

kind = sys.argv[1] if len(sys.argv) > 1 else "torus"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
A, L, _ = build(kind, n)
h = get_handle(0)
dev = A.indptr.device
PEAK = 6534.5


def bench(M, X, W, Y, reps=20, fused=True):
    kw = dict(alpha=0.7, beta=-0.2, gamma=0.1, W=W) if fused else {}
    for _ in range(3):
        M.spmm(X, Y, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        M.spmm(X, Y, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
for name, M in (("Lc", A),):
    for b in (64,):
        # panel of a wider block vector, like the eigensolver's V[:, p:p+b] (ld = 640)
        Vfull = torch.randn((M.nrows, 640 if M.nrows * 640 * 8 < 20e9 else b), dtype=torch.float64, device=dev)
        X = Vfull[:, :b]
        W = torch.randn((M.nrows, b), dtype=torch.float64, device=dev)
        Y0 = torch.empty_like(W); Y1 = torch.empty_like(W)
        M.enable_mma(False)
        for fused in (True, False):
            t0 = bench(M, X, W, Y0, fused=fused)
            by = M.spmm_bytes(b, fused)
            key = "%s_b%d_%s" % (name, b, "fused" if fused else "plain")
            out[key + "_gather"] = dict(ms=round(t0, 4), frac=round(by / t0 / 1e6 / PEAK, 3))
            print(key, "gather: %.4f ms frac %.3f" % (t0, by / t0 / 1e6 / PEAK), flush=True)
        mp = M.enable_mma(True)
        print("  plan: ksteps %d  reuse %.2f  fill %.2f  afrag %.0f MB" % (mp["ksteps"], mp["reuse"], mp["fill"], mp["afrag"].numel() * 8 / 1e6))
        for fused in (True, False):
            by = M.spmm_bytes(b, fused)
            key = "%s_b%d_%s" % (name, b, "fused" if fused else "plain")
            M.enable_mma(False); bench(M, X, W, Y0, reps=1, fused=fused); M.enable_mma(True)
            for gpw, pol in ((8, 0), (4, 0)):
                if not fused and (gpw, pol) != (8, 0):
                    continue
                h.set_option("mma_gpw", gpw); h.set_option("mma_stream_policy", pol)
                t1 = bench(M, X, W, Y1, fused=fused)
                err = float((Y0 - Y1).abs().max() / Y0.abs().max())
                out[key + "_mma_g%d_p%d" % (gpw, pol)] = dict(ms=round(t1, 4), frac=round(by / t1 / 1e6 / PEAK, 3), relerr=err)
                print(key, "mma gpw=%d pol=%d: %.4f ms frac %.3f  relerr %.1e" % (gpw, pol, t1, by / t1 / 1e6 / PEAK, err), flush=True)
        h.set_option("mma_gpw", 0); h.set_option("mma_stream_policy", 0)
        if name == "Lc":
            # node-contiguous layout [node][cp][q][e]
            def to_native(Z):
                return Z.reshape(M.nbrows, 2, b // 2, 2).permute(0, 2, 1, 3).contiguous().reshape(M.nbrows, 2 * b)
            def from_native(Zn):
                return Zn.reshape(M.nbrows, b // 2, 2, 2).permute(0, 2, 1, 3).contiguous().reshape(M.nrows, b)
            Xn, Wn = to_native(X), to_native(W)
            Yn = torch.empty_like(Xn)
            M.enable_mma(False); M.spmm(X, Y0, alpha=0.03, beta=-0.2, gamma=0.1, W=W); M.enable_mma(True)
            by = M.spmm_bytes(b, True)
            nk = mp["ksteps"]
            kc_r = mp["kcols"].clone(); af_c = torch.empty(nk * 16, dtype=torch.float64, device=dev)
            bad = torch.zeros(1, dtype=torch.int32, device=dev)
            h.call("rvgp_bsr_mma_rotc", I64(nk), mp["afrag"], kc_r, af_c, bad, 1e-12)
            print("rotc bad flag:", int(bad.item()))
            bufs = [Xn.clone(), Wn.clone(), Yn]
            def run_rec(rotc, rev, nl):
                # a 3-buffer recurrence like the Chebyshev filter: Y_{i+1} = f(X = Y_i, W = Y_{i-1})
                for i in range(nl):
                    Xb, Wb, Yb = bufs[(i + 1) % 3], bufs[i % 3], bufs[(i + 2) % 3]
                    h.call("rvgp_bsr_spmm_mma_native_f64", M.nbrows, mp["kptr"], kc_r if rotc else mp["kcols"], af_c if rotc else mp["afrag"],
                           int(rotc), Xb, I64(2 * b), Wb, I64(2 * b), Yb, I64(2 * b), int(b), 0.03, -0.2, 0.1, int(rev and (i & 1)))
            def check(rotc):
                h.call("rvgp_bsr_spmm_mma_native_f64", M.nbrows, mp["kptr"], kc_r if rotc else mp["kcols"], af_c if rotc else mp["afrag"],
                       int(rotc), Xn, I64(2 * b), Wn, I64(2 * b), Yn, I64(2 * b), int(b), 0.03, -0.2, 0.1, 0)
                return float((from_native(Yn) - Y0).abs().max() / Y0.abs().max())
            cfgs = []
            for var in (0, 1, 2, 3, 6, 11, 12, 13, 5):
                for gpw in (0, 4):
                    for pol in (3, 7):
                        cfgs.append((var, gpw, 1, pol, 1, 0))
            for var, gpw, pd, pol, rotc, rev in cfgs:
                h.set_option("mma_variant", var); h.set_option("mma_gpw", gpw); h.set_option("mma_prefetch", pd); h.set_option("mma_stream_policy", pol)
                err = check(rotc)
                bufs[0].copy_(Xn); bufs[1].copy_(Wn)
                run_rec(rotc, rev, 4)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); run_rec(rotc, rev, 24); e1.record(); torch.cuda.synchronize()
                t1 = e0.elapsed_time(e1) / 24
                key = "Lc_b%d_native_v%d_g%d_pd%d_pol%d_rotc%d_rev%d" % (b, var, gpw, pd, pol, rotc, rev)
                out[key] = dict(ms=round(t1, 4), frac=round(by / t1 / 1e6 / PEAK, 3), relerr=err)
                print("%s: %.4f ms frac %.3f relerr %.1e" % (key, t1, by / t1 / 1e6 / PEAK, err), flush=True)
            h.set_option("mma_variant", 0); h.set_option("mma_gpw", 0); h.set_option("mma_prefetch", 1); h.set_option("mma_stream_policy", 0)
            del Xn, Wn, Yn, bufs
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/profile_mma.json", "w"), indent=1)
