#!/usr/bin/env python
"""Marginal cost of the fused kernel's phases: times the kernel with parts switched off (ICMSLAM_SKIP)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from icm_slam_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
masks = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 16, 8, 24, 28, 30, 31, 63, 127, 255]
d = bench.make_data(name)
L_true, T, _ = bench.WORKLOADS[name]
cfg = bench.config_for(L_true)
eng = Engine(cfg)
eng.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
n = eng.extract()
x0 = d["odometry"][:, 0].copy()
for m in masks:
    os.environ["ICMSLAM_SKIP"] = str(m)
    ts = []
    for k in range(5):
        eng.set_map(d["map_init"]); eng.set_poses(d["x_init"])     # same inputs for every sample
        os.environ["ICMSLAM_SKIP"] = "0"
        eng.iterate(None, x0, 1)                                   # (establishes the grid, the hints and the labels)
        os.environ["ICMSLAM_SKIP"] = str(m)
        eng.iterate(None, x0, 1, timing=True)
        ts.append(eng.kernel_ms()[0])
    print("skip=%2d  fused kernel %.3f ms (min %.3f)" % (m, float(np.median(ts[1:])), min(ts)), flush=True)
