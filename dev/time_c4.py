"""quick timing on a synthetic workload: steady-state ms/sweep (graph replay), kernel split, cold sweep"""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.engine import Engine
from icm_slam_b200.synthetic import make_synthetic
L = int(sys.argv[1]) if len(sys.argv) > 1 else 316 * 316
T = 10 * L if L != 316 * 316 else 1_000_000
d = make_synthetic(L, T=T, seed=20181 + 4)
cfg = ConfigICM.from_values(N=1, L=2 * L, cota=20.0)
e = Engine(cfg)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); e.set_stream(st.cuda_stream)
e.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
n = e.extract()
x0 = d["odometry"][:, 0]
def ev(): return torch.cuda.Event(enable_timing=True)
e.set_map(d["map_init"]); e.set_poses(d["x_init"])
a, b = ev(), ev(); torch.cuda.synchronize(); a.record(st); e.iterate(None, x0, 1); b.record(st); torch.cuda.synchronize()
cold = a.elapsed_time(b)
e.iterate(None, x0, 5)
torch.cuda.synchronize()
K = 50
a, b = ev(), ev(); a.record(st)
for _ in range(K): e.iterate(None, x0, 1)
b.record(st); torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
kt = []
for _ in range(4):
    e.iterate(None, x0, 1, timing=True, stats=True); kt.append(e.kernel_ms()); s = e.sweep_stats()
print("n %d  cold %.3f ms  steady %.4f ms/sweep (%.0f sweeps/s)  runs+assoc %.4f  solve %.4f  k_runs %.4f  dirty %d/%d  env %s" % (
    n, cold, ms, 1000 / ms, np.mean([k[0] for k in kt]), np.mean([k[1] for k in kt]), s["k_runs_ns"] * 1e-6, s["dirty_tiles"], s["n_tiles"],
    {k: v for k, v in os.environ.items() if k.startswith("ICMSLAM")}))
