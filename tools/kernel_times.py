"""In-situ kernel time breakdown of one hot-path step (CUPTI through torch.profiler: every kernel of the process, ours included,
at its real in-pipeline duration -- unlike the ncu launch list, which serialises and cold-caches every launch).
usage: python tools/kernel_times.py [workload] [out.txt]"""
import collections, contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import RVGP

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
out = sys.argv[2] if len(sys.argv) > 2 else None
X, V, train_ind, test_ind, k = bench.make_inputs(wl)
dev = torch.device("cuda", 0)
Xd, Vd = torch.from_numpy(X).to(dev), torch.from_numpy(V).to(dev)


def step():
    d = RVGP.create_data_object(Xd, vectors=Vd, n_eigenpairs=k, verbose=False)
    with contextlib.redirect_stdout(io.StringIO()):
        gp = RVGP.fit(d, train_ind=train_ind, noise_variance=0.001)
    m, v = gp.transform(d, test_ind, as_device=True)
    torch.cuda.synchronize()
    return d


step()                                   # warm-up (lazy module loads, pools)
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    d = step()
wall = time.perf_counter() - t0
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("<")[0].split("(")[0].replace("void ", "").replace("rvgp::", "")
        a = agg[name]
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(a[1] for a in agg.values())
lines = ["# in-situ kernel times of one %s step (torch.profiler / CUPTI, wall %.3f s under the profiler; device busy %.3f s)" % (wl, wall, tot / 1e6),
         "# stages: %s" % {k_: round(v_, 3) for k_, v_ in d.timings.items()}]
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append("%-52s launches %6d  ms %9.2f  share %5.1f%%  avg us %9.1f" % (name[:52], cnt, us / 1e3, 100 * us / tot, us / cnt))
txt = "\n".join(lines)
print(txt)
if out:
    open(out, "w").write(txt + "\n")
