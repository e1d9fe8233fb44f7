import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from icm_slam_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
d = bench.make_data(name)
L_true, T, _ = bench.WORKLOADS[name]
cfg = bench.config_for(L_true)
eng = Engine(cfg)
eng.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
eng.extract()
x0 = d["odometry"][:, 0].copy()
eng.set_map(d["map_init"]); eng.set_poses(d["x_init"])
xp = d["x_init"]
for k in range(12):
    eng.iterate(None, x0, 1, timing=True)
    st = eng.sweep_stats()
    x = eng.get_poses()
    print("sweep %2d  kernel %.3f ms  cert_tiles %6d  far scans %6d  L %d->%d  epoch %d  G %.3e  stable %d  max|dx| %.2e" % (
        k + 1, eng.kernel_ms()[0], st["cert_tiles"], st["n_far_scans"], st["lsearch"], st["new_L"], st["cert_epoch"], st["cert_G_pm"] * 1e-12, st["stable_ids"],
        np.abs(x - xp).max()), flush=True)
    xp = x
