"""debug: RUNS=0 vs RUNS=1 on data_IJAC2018 (golden inputs): where do the results differ?"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CONFIG_ROS, c1_inputs, golden
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.engine import Engine
g = golden("c1_ref.npz")
z, odo, u = c1_inputs()
res = {}
for runs in ("0", "1"):
    os.environ["ICMSLAM_RUNS"] = runs
    e = Engine(ConfigICM.from_values(**CONFIG_ROS))
    e.load(z, odo, u, precondition=True); e.extract()
    e.set_map(g["p0_map"]); e.set_poses(np.ascontiguousarray(g["p0_x"].copy()))
    out = []
    for k in range(6):
        e.iterate(None, odo[:, 0], 1, stats=True)
        st = e.sweep_stats()
        raw, cnt, rl = e.raw_map()
        out.append((e.get_poses().copy(), e.get_map().copy(), e.associations().copy(), raw, cnt, st))
    res[runs] = out
    e.close()
for k in range(6):
    a, b = res["0"][k], res["1"][k]
    print(k, "dirty", a[5]["dirty_tiles"], b[5]["dirty_tiles"], "/", b[5]["n_tiles"], "n_ind", a[5]["n_ind"], b[5]["n_ind"], "far", a[5]["n_far_scans"], b[5]["n_far_scans"],
          "x", np.abs(a[0] - b[0]).max(), "map", a[1].shape, b[1].shape, np.abs(a[1] - b[1]).max() if a[1].shape == b[1].shape else None,
          "labels", int((a[2] != b[2]).sum()), "raw", np.abs(a[3] - b[3]).max() if a[3].shape == b[3].shape else (a[3].shape, b[3].shape),
          "cnt", np.abs(a[4] - b[4]).max() if a[4].shape == b[4].shape else None)
    if a[3].shape == b[3].shape and np.abs(a[3] - b[3]).max() > 0:
        i = np.argwhere(np.abs(a[3] - b[3]) > 0)
        print("   raw diff at", i[:6].tolist(), "counts", a[4][i[:6, 1]], b[4][i[:6, 1]])
