"""profiling driver: C4, a few sweeps without CUDA graphs (run under ncu)"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ICMSLAM_GRAPH"] = "0"
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.engine import Engine
from icm_slam_b200.synthetic import make_synthetic
L = 316 * 316
nsweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = make_synthetic(L, T=1_000_000, seed=20181 + 4)
cfg = ConfigICM.from_values(N=1, L=2 * L, cota=20.0)
e = Engine(cfg)
e.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
print("n", e.extract())
e.set_map(d["map_init"]); e.set_poses(d["x_init"])
e.iterate(None, d["odometry"][:, 0], nsweeps)
e.synchronize()
print("done", e.landmarks_actuales)
