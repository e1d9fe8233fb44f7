"""Shared helpers for the test-suite: golden fixture access and oracle configuration."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CONFIG_ROS = dict(N=30, deltat=0.1, L=1000, Q=[1, 1], R=[1, 1, 1], cte_odom=1.0, cota=300.0, dist_thr=1.0,
                  dist_thr_obs=1.0, rango_laser_max=10.0, radio=0.137)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def c1_inputs():
    g = golden("c1_inputs.npz")
    return g["observations"].astype(np.float64), g["odometry"], g["velocities"]


def c2_inputs():
    g = golden("c1_inputs.npz")
    g2 = golden("c2_inputs.npz")
    return g2["observations"].astype(np.float64), g["odometry"], g["velocities"]


def split_labels(labels, nt_calls, off):
    """Golden label streams are concatenated per `actualizar` call (non-empty scans only);
    the CSR offsets of the extraction give the same partition."""
    nt = np.diff(off)
    assert np.array_equal(nt[nt > 0], nt_calls)
    return labels
