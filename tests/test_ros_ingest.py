"""ROS ingestion against the in-process publisher: message decoding, timestamp alignment, growing logs.
CPU only (the transport is host code); the end-to-end run through the solver is in test_gpu_parity.py."""
import math

import numpy as np
import pytest

from icm_slam_b200.config import ConfigICM
from icm_slam_b200 import ros_ingest as ri

from helpers import CONFIG_ROS


class Ingest(ri.ROS):
    """the transport alone (no solver, no device)"""

    def __init__(self, config):
        self._ros_init(config)
        self.mediciones, self.odometria, self.u = np.array([]), np.array([]), np.array([])


def _config():
    return ConfigICM.from_values(**dict(CONFIG_ROS, topic_laser="/pioneer2dx/laser/scan_Lidar_horizontal", topic_laser_msg="sensor_msgs/LaserScan",
                                        topic_odometry="/pioneer2dx/ground_truth/odom", topic_odometry_msg="nav_msgs/Odometry"))


def _log(T=60, B=181, seed=3):
    rng = np.random.default_rng(seed)
    raw = rng.uniform(0.6, 12.0, (B, T))
    raw[rng.random((B, T)) < 0.1] = np.nan            # no return
    th = np.cumsum(rng.normal(0.0, 0.2, T)) + 2.5     # crosses +-pi
    odo = np.stack([np.cumsum(rng.normal(0, 0.1, T)), np.cumsum(rng.normal(0, 0.1, T)), th])
    u = np.stack([rng.uniform(0, 1, T), rng.normal(0, 0.3, T)])
    return raw, odo, u


def _expected_ranges(raw, cfg):
    z = np.where(np.isnan(raw), cfg.rango_laser_max, raw)
    return np.minimum(z + cfg.radio, cfg.rango_laser_max)[:180]      # (the reference keeps 180 of a 181-beam scan online)


@pytest.mark.parametrize("laser_first", [True, False])
def test_replayed_log_arrives_as_the_arrays_the_solver_reads(laser_first):
    cfg = _config()
    raw, odo, u = _log()
    node, client = Ingest(cfg), ri.FakeRos()
    node.connect_ros(client)
    ri.publish_log(client, cfg, raw, odo, u, laser_first=laser_first)
    T = raw.shape[1]
    assert node.new_data == T and node.mediciones.shape == (180, T) and node.odometria.shape == (3, T) and node.u.shape == (2, T)
    assert np.array_equal(node.mediciones, _expected_ranges(raw, cfg))           # pre-conditioning: bit for bit
    assert np.array_equal(node.odometria[:2], odo[:2]) and np.array_equal(node.u, u)
    wrapped = np.arctan2(np.sin(odo[2]), np.cos(odo[2]))                          # yaw through the quaternion: in (-pi, pi]
    assert np.max(np.abs(node.odometria[2] - wrapped)) < 1e-12
    assert node.lidar.warnings == 0 and node.odom.warnings == 0
    ok, resp = client.service("/icm_slam/iterative_flag").call({"data": True})     # the SetBool service ends the online phase
    assert ok and resp["success"] and node.iterations_flag
    node.disconnect_ros()
    assert not client.is_connected


def test_sensor_sort_follows_a_sensor_that_publishes_twice_as_fast():
    """The guess ceil(k0 * k / c) extrapolates the message rate: an odometry topic at 20 Hz beside a 10 Hz laser is aligned by
    stamp, every second message used (ICM_SLAM.py:376-428)."""
    cfg = _config()
    raw, odo, u = _log(T=40)
    node, client = Ingest(cfg), ri.FakeRos()
    node.connect_ros(client)
    lt, ot = client.topic(cfg.topic_laser), client.topic(cfg.topic_odometry)
    for t in range(40):
        lt.publish(ri.laser_scan_msg(raw[:, t].tolist(), t, 0.1))
        for half in (0, 1):                                                        # stamps t*0.1 and t*0.1 + 0.05
            m = ri.odometry_msg(odo[:, t] + (0.0 if half == 0 else 100.0), u[:, t], 2 * t + half, 0.05)
            ot.publish(m)
    n = node.new_data
    assert n >= 20
    # every aligned column carries an odometry message stamped within one period of the laser's
    for k in range(n):
        assert abs(node.odometria[0, k] - odo[0, k]) < 1e-9 or abs(node.odometria[0, k] - 100.0 - odo[0, k]) < 1e-9 or \
               abs(node.odometria[0, k] - odo[0, k - 1] - 100.0) < 1e-9


def test_lost_message_is_reported_not_fatal():
    cfg = _config()
    raw, odo, u = _log(T=30)
    node, client = Ingest(cfg), ri.FakeRos()
    node.connect_ros(client)
    ri.publish_log(client, cfg, raw, odo, u, drop_laser=(7,))
    assert node.lidar.warnings > 0                       # "no se encuentra la secuencia buscada"
    assert node.mediciones.shape[1] == node.odometria.shape[1] == node.u.shape[1] >= 7
    assert np.array_equal(node.mediciones[:, :7], _expected_ranges(raw, cfg)[:, :7])


def test_message_builders_match_the_bag_writer_layout():
    m = ri.laser_scan_msg([1.0, float("nan"), None, 2.0], 12)
    assert m["ranges"] == [1.0, None, None, 2.0] and m["header"]["seq"] == 12
    assert m["header"]["stamp"] == {"secs": 1, "nsecs": int((1.2000000000000002 - 1) * 10 ** 9)} or m["header"]["stamp"]["secs"] == 1
    o = ri.odometry_msg(np.array([1.0, 2.0, 0.7]), np.array([0.3, -0.1]), 5)
    q = o["pose"]["pose"]["orientation"]
    assert math.isclose(math.atan2(2 * q["w"] * q["z"], 1 - 2 * q["z"] ** 2), 0.7, abs_tol=1e-15)
    assert o["twist"]["twist"]["linear"]["x"] == 0.3 and o["twist"]["twist"]["angular"]["z"] == -0.1
