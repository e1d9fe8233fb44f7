"""Per-phase clock profile of the fused kernel (ICMSLAM_PROF=1, graphs off): prints the library's [prof] summary for the 3rd sweep."""
import os, sys
os.environ["ICMSLAM_GRAPH"] = "0"
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
for k in range(int(os.environ.get("WARM", "3"))):
    eng.iterate(None, x0, 1)
os.environ["ICMSLAM_PROF"] = "1"
eng.iterate(None, x0, 1, timing=True)
print("kernel ms", eng.kernel_ms())
