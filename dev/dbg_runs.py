"""debug: per-sweep statistics of the run-record path on a synthetic workload"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.engine import Engine
from icm_slam_b200.synthetic import make_synthetic
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10 * L
d = make_synthetic(L, T=T, seed=20184)
cfg = ConfigICM.from_values(N=1, L=2 * L, cota=20.0)
e = Engine(cfg)
e.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
print("n", e.extract())
e.set_map(d["map_init"]); e.set_poses(d["x_init"])
for k in range(6):
    e.iterate(None, d["odometry"][:, 0], 1, stats=True, timing=True)
    st = e.sweep_stats()
    print(k, {a: st[a] for a in ("dirty_tiles", "n_tiles", "epoch", "stable_ids", "n_far_scans", "new_L", "lsearch", "raw_L", "n_ind", "k_runs_ns")}, e.kernel_ms())
