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
if os.environ.get("ICMSLAM_TRACE"):
    import ctypes as C
    e.iterate(None, x0, 12)
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 256)(); cn = (C.c_uint32 * 3)()
    e.lib.icmslam_get_trace(e._h, buf, cn)
    tr = np.array(buf[:], dtype=np.uint64).reshape(32, 8).astype(np.int64)
    k_last = int(cn[0]) - 1
    names = ["runs", "assoc", "labels", "reduce", "steady", "solve", "halo", "steady_end"]
    for k in range(k_last - 5, k_last + 1):
        r = tr[k & 31]
        t0 = r[0]
        nxt = tr[(k + 1) & 31][0] if k < k_last else 0
        print("sweep %d: " % k + "  ".join("%s %+.1f" % (n, (v - t0) / 1e3) for n, v in zip(names, r) if n not in ("reduce", "halo")) + ("  next_runs %+.1f" % ((nxt - t0) / 1e3) if nxt else ""))
