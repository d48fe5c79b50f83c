"""cProfile of RVGP.fit + transform at a bench workload (default c4): which host-side pieces of the fit are not the L-BFGS-B
evaluations.  usage: python tools/profile_fit.py [workload]"""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import RVGP

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
X, V, train_ind, test_ind, k = bench.make_inputs(wl)
dev = torch.device("cuda", 0)
Xd, Vd = torch.from_numpy(X).to(dev), torch.from_numpy(V).to(dev)
d = RVGP.create_data_object(Xd, vectors=Vd, n_eigenpairs=k, verbose=False)
torch.cuda.synchronize()
for rep in range(2):
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    with contextlib.redirect_stdout(io.StringIO()):
        gp = RVGP.fit(d, train_ind=train_ind, noise_variance=0.001)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    m, v = gp.transform(d, test_ind, as_device=True)
    torch.cuda.synchronize()
    pr.disable()
    t2 = time.perf_counter()
    print("rep %d: fit %.3f s (evaluations %s), transform %.3f s" % (rep, t1 - t0, getattr(getattr(gp, "_gpr", None), "n_eval", None), t2 - t1))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
