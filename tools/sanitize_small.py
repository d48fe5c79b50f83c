"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import RVGP
from tests.conftest import load_golden
from RVGP.geometry import furthest_point_sampling
cases = (("torus_n600_k20", 10),) if "--one" in sys.argv else (("torus_n600_k20", 10), ("flat3torus_R6_n900_k24", 10), ("sheet_R20_n500_k16", 14))
for case, nb in cases:
    gd = load_golden(case)
    d = RVGP.create_data_object(gd["X"], n_neighbors=nb, n_eigenpairs=len(gd["evals_Lc"]), verbose=False)
    d.random_vector_field(seed=1); d.smooth_vector_field(t=10)
    gp = RVGP.fit(d, train_ind=np.arange(0, d.n, 2), epochs=10)
    m, v = gp.transform(d, np.arange(1, d.n, 2))
    print(case, "ok", float(abs(d.evals_Lc - gd["evals_Lc"]).max()))
p, l = furthest_point_sampling(gd["X"], spacing=0.2); print("fps", len(p))
from rvgp_b200.gp import DeviceGPR
X = torch.randn((300, 20), dtype=torch.float64, device="cuda"); Y = torch.randn((300, 1), dtype=torch.float64, device="cuda")
print(DeviceGPR(X, Y, solver="dense").lml_and_grads(np.ones(20), 0.5)[0])
