"""ConfigICM -- the reference's configuration object (ICM_SLAM.py:60-102), same keys and attributes.

Differences, both deliberate and documented in DESIGN.md:
  * the ROS-only keys (`topic_*`, `time`, `file`, `dist_thr_obs`) are optional, so the shipped
    `config_default.yaml` -- which the reference's own ConfigICM cannot load (KeyError at
    ICM_SLAM.py:92) -- loads here;
  * extra keys select the restated sweep variants (`schedule`, `solver`, `map_view`), defaulting
    to the B200-native fast path.
"""
from __future__ import annotations

import numpy as np

DEFAULTS = dict(N=2, deltat=0.1, L=1000, Q=[1, 1], R=[1, 1, 1], cte_odom=1.0, cota=300.0, dist_thr=1.0, dist_thr_obs=1.0,
                rango_laser_max=10.0, radio=0.137)


class ConfigICM:
    def __init__(self, configFile="config_default.yaml", D=None):
        if not D:
            import yaml
            with open(configFile, "r") as arch:
                D = yaml.safe_load(arch)["D"]
        self.N = D["N"]                      # ICM passes
        self.deltat = D["deltat"]            # sampling period
        self.L = D["L"]                      # max landmarks (label capacity)
        self.Q = np.eye(2)                   # observation weights (used un-inverted, sensors.py:202)
        self.Q[0, 0] = D["Q"][0]
        self.Q[1, 1] = D["Q"][1]
        self.R = np.eye(3)                   # motion weights (sensors.py:240)
        self.R[0, 0] = D["R"][0]
        self.R[1, 1] = D["R"][1]
        self.R[2, 2] = D["R"][2]
        self.cte_odom = D["cte_odom"]
        self.cota = D["cota"]
        self.dist_thr = D["dist_thr"]
        self.dist_thr_obs = D.get("dist_thr_obs", D["dist_thr"])   # read but never used by the reference (:88)
        self.rango_laser_max = D["rango_laser_max"]
        self.radio = D["radio"]
        self.topic_laser = D.get("topic_laser", "")
        self.topic_laser_msg = D.get("topic_laser_msg", "")
        self.topic_odometry = D.get("topic_odometry", "")
        self.topic_odometry_msg = D.get("topic_odometry_msg", "")
        self.file = D.get("file", "")
        self.time = D.get("time", 0.0)
        # sweep variant (not in the reference): see DESIGN.md "modes"
        self.schedule = D.get("schedule", "redblack")
        self.solver = D.get("solver", "newton")
        self.map_view = D.get("map_view", "prev")
        self.device = D.get("device", 0)

    def set_Tf(self, Tf):
        self.Tf = Tf

    @classmethod
    def from_values(cls, **kw):
        D = dict(DEFAULTS)
        D.update(kw)
        return cls(D=D)
