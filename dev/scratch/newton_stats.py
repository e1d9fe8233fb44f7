import sys, numpy as np
sys.path.insert(0, "/root/repo")
import bench
from icm_slam_b200.engine import Engine
d = bench.make_data("c3")
L_true, T, _ = bench.WORKLOADS["c3"]
cfg = bench.config_for(L_true)
eng = Engine(cfg)
eng.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
n = eng.extract()
x0 = d["odometry"][:, 0].copy()
eng.set_map(d["map_init"]); eng.set_poses(d["x_init"])
xp = d["x_init"]
for k in range(12):
    eng.iterate(None, x0, 1, stats=True)
    st = eng.sweep_stats()
    x = eng.get_poses()
    print("sweep %2d  newton iters/pose %.2f  far scans %d  L %d  max|dx| %.2e %.2e" % (k + 1, st["newton_iters"] / T, st["n_far_scans"], st["new_L"], np.abs(x - xp)[:2].max(), np.abs(x - xp)[2].max()))
    xp = x
