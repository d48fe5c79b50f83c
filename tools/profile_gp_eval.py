"""Where one evaluation of the rank-k GP objective spends its time (C4: k = 256, ~115 evaluations per fit):
device time of the captured evaluation graph alone, and wall time of the whole host-side evaluation (parameter upload, graph
launch, read-back, numpy gradient assembly).  usage: python tools/profile_gp_eval.py [k] [M]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rvgp_b200._cabi import get_handle, I64
from rvgp_b200.gp import DeviceGPR

k = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
X = torch.from_numpy(rng.normal(size=(M, k)) / np.sqrt(M)).to(dev)
Y = torch.from_numpy(rng.normal(size=(M, 1))).to(dev)
gp = DeviceGPR(X, Y, solver="lowrank")
S = np.exp(-np.linspace(0, 6, k)) * M
for _ in range(3):
    gp.lml_and_grads(S, 0.1)
torch.cuda.synchronize()
h = gp.h
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
e0.record()
for _ in range(reps):
    h.call("rvgp_gp_lowrank_eval_f64", int(k), gp._Gd, gp._bd, gp._par_d, gp._out_d, gp._ev_ws, I64(gp._ev_wsb))
e1.record(); torch.cuda.synchronize()
print("k=%d  device time per evaluation (graph replay, back to back): %.3f ms" % (k, e0.elapsed_time(e1) / reps))
t0 = time.perf_counter()
for i in range(reps):
    gp.lml_and_grads(S * (1 + 1e-3 * i), 0.1)
t1 = time.perf_counter()
print("k=%d  wall time per lml_and_grads: %.3f ms" % (k, (t1 - t0) / reps * 1e3))
# per-kernel device time inside the replayed graph (CUPTI)
import collections
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20):
        h.call("rvgp_gp_lowrank_eval_f64", int(k), gp._Gd, gp._bd, gp._par_d, gp._out_d, gp._ev_ws, I64(gp._ev_wsb))
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t_first, t_last = None, None
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("<")[0].split("(")[0].replace("void ", "").replace("rvgp::", "")
        agg[name][0] += 1
        agg[name][1] += ev.device_time
        s0 = ev.time_range.start
        t_first = s0 if t_first is None else min(t_first, s0)
        t_last = max(t_last or 0, ev.time_range.end)
print("span per evaluation %.1f us; busy %.1f us" % ((t_last - t_first) / 20, sum(a[1] for a in agg.values()) / 20))
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-40s per eval: launches %5.1f  us %8.1f  avg us %7.1f" % (name[:40], cnt / 20, us / 20, us / cnt))
