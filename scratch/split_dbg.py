import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import CONFIG_ROS, c1_inputs, golden
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.engine import Engine
z, odo, u = c1_inputs()
g = golden("c1_ref.npz")
cfg = ConfigICM.from_values(**CONFIG_ROS)
res = {}
for split in (0, 1):
    os.environ["ICMSLAM_SPLIT"] = str(split)
    e = Engine(cfg, device=0)
    e.load(z, odo, u, precondition=True)
    e.extract()
    xg = np.ascontiguousarray(g["p0_x"].copy())
    m = np.ascontiguousarray(g["p0_map"].copy())
    e.landmarks_actuales = m.shape[1]
    st, Lout, mout = e.sweep(m, xg, odo[:, 0], fused=True)
    res[split] = (xg.copy(), e.associations().copy())
    e.close()
d = np.abs(res[0][0] - res[1][0]).max(axis=0)
bad = np.nonzero(d > 1e-9)[0]
print("n bad", bad.size, "labels equal", np.array_equal(res[0][1], res[1][1]))
print("bad t:", bad[:60])
print("t mod 126:", (bad % 126)[:60])
print("d:", d[bad][:20])
