"""TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference modules.

This file imports the reference's own Python modules from /root/reference/scripts
(read-only, present only in the build container) or from the git-ignored copy
baseline/_ref/scripts (made by baseline/setup_ref.py; it travels to the GPU box)
behind stub `roslibpy` / `matplotlib` modules, and replays the reference's offline
path without a ROS network.  It is used by `oracle/make_golden.py` to mint the
fixtures committed under tests/golden/, by CPU tests that are skipped when the
reference is absent, and by bench.py's reference arm / cpu_baseline leg (the timed
CPU baseline).  Nothing in the product package imports it.

What is replayed (reference file:line):
  * pass 0  = sensors.py:61-104 (`ICM_ROS.inicializar_online` minus the ROS wait loop)
              driving the unmodified `inicializar_online_process` (sensors.py:106-123)
  * sweep   = the unmodified `ICM_ROS.iterations_process_offline` (sensors.py:125-168)
  * range pre-conditioning = sensors_definitions.py:22 / IJAC2018_python.txt:43
"""
from __future__ import annotations

import os
import sys
import types
from copy import deepcopy

import numpy as np

def _find_ref_dir():
    """ICM_REF_DIR, else the read-only reference of the build container, else the copy baseline/setup_ref.py ships to the GPU box."""
    env = os.environ.get("ICM_REF_DIR")
    if env:
        return env
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in ("/root/reference/scripts", os.path.join(here, "baseline", "_ref", "scripts")):
        if os.path.isfile(os.path.join(d, "sensors.py")):
            return d
    return "/root/reference/scripts"


REF_DIR = _find_ref_dir()

CONFIG_ROS = dict(  # values of scripts/config_ros.yaml
    N=30, deltat=0.1, L=1000, Q=[1, 1], R=[1, 1, 1], cte_odom=1.0, cota=300.0,
    dist_thr=1.0, dist_thr_obs=1.0, rango_laser_max=10.0, radio=0.137,
    topic_laser="/pioneer2dx/laser/scan_Lidar_horizontal", topic_laser_msg="sensor_msgs/LaserScan",
    topic_odometry="/pioneer2dx/ground_truth/odom", topic_odometry_msg="nav_msgs/Odometry",
    file="data_IJAC2018.mat", time=275.0, sensors="revisar",
)


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "sensors.py"))


def _install_stubs():
    if "roslibpy" not in sys.modules:
        sys.modules["roslibpy"] = types.ModuleType("roslibpy")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        for name in ("figure", "plot", "axis", "show", "pause", "clf"):
            setattr(plt, name, lambda *a, **k: None)
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


_mods = None


def load_reference():
    """Returns (sensors, ICM_SLAM) reference modules, unmodified."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("reference not present at %s" % REF_DIR)
        _install_stubs()
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import ICM_SLAM as ref_icm  # noqa
        import sensors as ref_sensors  # noqa
        _mods = (ref_sensors, ref_icm)
    return _mods


def make_config(**over):
    _, ref_icm = load_reference()
    D = dict(CONFIG_ROS)
    D.update(over)
    return ref_icm.ConfigICM(D=D)


def precondition(z, config):
    """IJAC2018_python.txt:43 == sensors_definitions.py:21-22."""
    z = np.array(z, dtype=np.float64)
    z[np.isnan(z)] = config.rango_laser_max
    return np.minimum(z + config.radio, z * 0.0 + config.rango_laser_max)


def make_solver(config, mediciones, odometria, u, x0=None):
    """An ICM_ROS instance wired with arrays instead of ROS sensors (ctor bypassed:
    sensors.py:16-49 only builds roslibpy listeners)."""
    ref_sensors, ref_icm = load_reference()
    s = ref_sensors.ICM_ROS.__new__(ref_sensors.ICM_ROS)
    s.config = config
    s.new_data = 0
    s.mediciones = np.ascontiguousarray(mediciones, dtype=np.float64)
    s.odometria = np.ascontiguousarray(odometria, dtype=np.float64)
    s.u = np.ascontiguousarray(u, dtype=np.float64)
    s.x0 = np.array([s.odometria[:, 0]]).T if x0 is None else np.asarray(x0, float).reshape(3, 1)
    s.iterations_flag = True
    s.debug = False
    s.mapa_obj = ref_icm.Mapa(config)
    return s


def pass0(s, log=None):
    """Replay of sensors.py:61-104 (offline: column 0 is 'the first arrived scan')."""
    ref_sensors, ref_icm = load_reference()
    cfg = s.config
    s.x0 = np.array([s.odometria[:, 0]]).T
    xt = deepcopy(s.x0)
    x = deepcopy(s.x0)
    y = np.zeros((2, cfg.L))
    s.mapa_obj = ref_icm.Mapa(cfg)
    z = ref_icm.filtrar_z(s.mediciones[:, 0], cfg)
    zt = ref_icm.tras_rot_z(xt, z)
    y, c = s.mapa_obj.actualizar(y, y, zt[:, 2:4])
    if log is not None:
        log.append(np.asarray(c).copy())
    T = s.mediciones.shape[1]
    for t in range(1, T):
        s.t = t
        if log is not None:
            _hook_actualizar(s, log)
        y, xt = s.inicializar_online_process(y, xt)
        xt = np.reshape(xt, (3, 1))
        x = np.concatenate((x, xt), axis=1)
    _unhook(s)
    raw_L = s.mapa_obj.landmarks_actuales
    raw_counts = s.mapa_obj.cant_obs_i.copy()
    raw_map = y.copy()
    yy = s.mapa_obj.filtrar(y)
    yy = yy[:, : s.mapa_obj.landmarks_actuales]
    s.mapa_viejo = deepcopy(yy)
    s.positions = deepcopy(x)
    return dict(mapa=yy.copy(), x=x.copy(), raw_L=raw_L, raw_counts=raw_counts, raw_map=raw_map)


def _hook_actualizar(s, log):
    m = s.mapa_obj
    if getattr(m, "_hooked", False):
        return
    orig = m.actualizar

    def wrapped(mapa, ref, obs):
        out, c = orig(mapa, ref, obs)
        log.append(np.asarray(c).copy())
        return out, c

    m.actualizar = wrapped
    m._hooked = True
    m._orig = orig


def _unhook(s):
    m = s.mapa_obj
    if getattr(m, "_hooked", False):
        del m.actualizar
        m._hooked = False


def sweep(s, mapa_viejo, x, record=False):
    """One unmodified `iterations_process_offline`.  With record=True also returns the
    per-scan association label arrays, the raw (pre-filtrar) map / counts / Lact and the
    NM evaluation count."""
    ref_sensors, ref_icm = load_reference()
    rec = {}
    if record:
        labels = []
        _hook_actualizar(s, labels)
        m = s.mapa_obj
        orig_f = m.filtrar

        def filtrar_wrapped(mapa):
            rec["raw_L"] = int(m.landmarks_actuales)
            rec["raw_counts"] = m.cant_obs_i[: rec["raw_L"]].copy()
            rec["raw_map"] = np.array(mapa[:, : rec["raw_L"]], copy=True)
            return orig_f(mapa)

        m.filtrar = filtrar_wrapped
        nev = [0]
        of_xn, of_x = s.fun_xn, s.fun_x
        s.fun_xn = lambda v: (nev.__setitem__(0, nev[0] + 1), of_xn(v))[1]
        s.fun_x = lambda v: (nev.__setitem__(0, nev[0] + 1), of_x(v))[1]
    x_in = x.copy()
    mapa_out, x_out = s.iterations_process_offline(mapa_viejo, x)
    if record:
        _unhook(s)
        del s.mapa_obj.filtrar
        del s.fun_xn, s.fun_x
        rec["labels"] = labels
        rec["nev"] = nev[0]
        rec["x_in"] = x_in
        rec["counts_out"] = s.mapa_obj.cant_obs_i[: s.mapa_obj.landmarks_actuales].copy()
    return mapa_out, x_out, rec


def load_ijac(path=None):
    import scipy.io as sio
    d = sio.loadmat(path or os.path.join(REF_DIR, "data_IJAC2018.mat"))
    return (np.array(d["observations"], float), np.array(d["odometry"], float),
            np.array(d["velocities"], float))


def load_palomar(path=None):
    import scipy.io as sio
    d = sio.loadmat(path or os.path.join(REF_DIR, "datos_palomar1.mat"))["datos"][0, 0]
    return (np.array(d["observaciones"], float), np.array(d["odometria"], float),
            np.array(d["control"], float))
