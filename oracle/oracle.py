"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/icm_oracle.c (the CPU parity oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libicm_oracle.so")

SCHED = {"sequential": 0, "redblack": 1}
SOLVER = {"nm": 0, "newton": 1}
VIEW = {"running": 0, "full": 1, "prev": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "icm_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-comment",
                               "-o", _SO, src, "-lm"])
    return _SO


class OrcConfig(C.Structure):
    _fields_ = [("deltat", C.c_double), ("q1", C.c_double), ("q2", C.c_double), ("r1", C.c_double),
                ("r2", C.c_double), ("r3", C.c_double), ("cte_odom", C.c_double), ("cota", C.c_double),
                ("dist_thr", C.c_double), ("rango_laser_max", C.c_double), ("radio", C.c_double),
                ("L", C.c_int32), ("pad_", C.c_int32)]


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_entrepi.restype = C.c_double
        _lib.orc_entrepi.argtypes = [C.c_double]
        _lib.orc_fun_x.restype = C.c_double
        _lib.orc_fun_xn.restype = C.c_double
        _lib.orc_solve_pose.restype = C.c_long
        _lib.orc_extract_all.restype = C.c_long
        _lib.orc_mapa_new.restype = C.c_void_p
        _lib.orc_mapa_counts.restype = _dp
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_cfg(config=None, **kw) -> OrcConfig:
    """From a ConfigICM-like object (attributes Q, R as 2x2/3x3 or lists) or keyword values."""
    g = (lambda k, dflt=None: kw[k] if k in kw else (getattr(config, k) if config is not None and hasattr(config, k) else dflt))
    Q = np.asarray(g("Q", [1, 1]), float)
    R = np.asarray(g("R", [1, 1, 1]), float)
    q = np.diag(Q) if Q.ndim == 2 else Q
    r = np.diag(R) if R.ndim == 2 else R
    return OrcConfig(float(g("deltat", 0.1)), q[0], q[1], r[0], r[1], r[2], float(g("cte_odom", 1.0)),
                     float(g("cota", 300.0)), float(g("dist_thr", 1.0)), float(g("rango_laser_max", 10.0)),
                     float(g("radio", 0.137)), int(g("L", 1000)), 0)


def beam_tables(B: int):
    """ang[i]=(i*pi)/180, cos, sin exactly as numpy forms them in ICM_SLAM.py:44,51-53."""
    nind = np.arange(B)
    ang = nind * np.pi / 180.0
    return ang, np.cos(ang), np.sin(ang)


def entrepi(a: float) -> float:
    return lib().orc_entrepi(float(a))


def precondition(z, radio, rmax):
    z = _c(z)
    out = np.empty_like(z)
    lib().orc_precondition(_d(z), C.c_long(z.size), C.c_double(radio), C.c_double(rmax), _d(out))
    return out


def filtrar_obs(obs, max_dist=10.0, cant_max=15):
    obs = _c(obs)
    B, T = obs.shape
    out = np.empty_like(obs)
    a = np.empty(T, np.int32)
    rc = lib().orc_filtrar_obs(_d(obs), B, T, C.c_long(T), C.c_double(max_dist), int(cant_max), _d(out),
                               C.c_long(T), _i(a))
    if rc != 0:
        raise RuntimeError("orc_filtrar_obs rc=%d" % rc)
    return out, a


def extract_all(med, cfg: OrcConfig):
    """filtrar_z on every column -> dict(off, beam, d, bx, by, ang)."""
    med = _c(med)
    B, T = med.shape
    ang, cb, sb = beam_tables(B)
    off = np.zeros(T + 1, np.int32)
    cap = B * T
    beam = np.empty(cap, np.int32)
    d = np.empty(cap)
    bx = np.empty(cap)
    by = np.empty(cap)
    n = lib().orc_extract_all(_d(med), B, T, C.c_long(T), _d(cb), _d(sb), C.c_double(cfg.rango_laser_max),
                              C.c_double(cfg.dist_thr), _i(off), _i(beam), _d(d), _d(bx), _d(by))
    if n < 0:
        raise RuntimeError("orc_extract_all rc=%d" % n)
    return dict(off=off, beam=beam[:n].copy(), d=d[:n].copy(), bx=bx[:n].copy(), by=by[:n].copy(), ang=ang,
                n=int(n), T=T, B=B)


def filtrar_z(z, cfg: OrcConfig):
    """One scan; returns the (n,4) array [d, ang, bx, by] like ICM_SLAM.filtrar_z, plus beam idx."""
    z = _c(z).reshape(-1)
    e = extract_all(z.reshape(-1, 1), cfg)
    return np.stack([e["d"], e["ang"][e["beam"]], e["bx"], e["by"]], axis=1), e["beam"]


def tras_rot(pose, bx, by):
    pose = _c(pose).reshape(3)
    bx, by = _c(bx), _c(by)
    wx = np.empty_like(bx)
    wy = np.empty_like(by)
    lib().orc_tras_rot(_d(pose), int(bx.size), _d(bx), _d(by), _d(wx), _d(wy))
    return wx, wy


class Mapa:
    """State of ICM_SLAM.Mapa (landmarks_actuales, cant_obs_i) held by the C oracle."""

    def __init__(self, cfg: OrcConfig):
        self.cfg = cfg
        self.L = int(cfg.L)
        self._h = C.c_void_p(lib().orc_mapa_new(self.L))

    def __del__(self):
        try:
            lib().orc_mapa_free(self._h)
        except Exception:
            pass

    @property
    def landmarks_actuales(self):
        return int(lib().orc_mapa_get_lact(self._h))

    @landmarks_actuales.setter
    def landmarks_actuales(self, v):
        lib().orc_mapa_set_lact(self._h, int(v))

    @property
    def cant_obs_i(self):
        return np.ctypeslib.as_array(lib().orc_mapa_counts(self._h), shape=(self.L,))

    def clear_obs(self):
        lib().orc_mapa_clear_obs(self._h)

    def actualizar(self, mapa, mapa_referencia, obs):
        """Branch B only.  mapa: 2xL C-contiguous float64, updated in place."""
        assert mapa.flags.c_contiguous and mapa.shape == (2, self.L)
        ref = _c(mapa_referencia)
        obs = _c(obs)
        wx, wy = _c(obs[:, 0]), _c(obs[:, 1])
        c = np.empty(obs.shape[0], np.int32)
        rc = lib().orc_actualizar(self._h, C.c_double(self.cfg.dist_thr), _d(mapa), _d(ref), int(ref.shape[1]),
                                  C.c_long(ref.shape[1]), int(obs.shape[0]), _d(wx), _d(wy), _i(c))
        if rc == -2:
            raise IndexError("label capacity L exceeded (ICM_SLAM.py:191)")
        return mapa, c

    def filtrar(self, mapa):
        mapa = _c(mapa)
        out = np.zeros((2, self.L))
        rc = lib().orc_filtrar(self._h, C.c_double(self.cfg.cota), C.c_double(self.cfg.dist_thr), _d(mapa), _d(out))
        if rc == -4:
            raise ValueError("no landmark survives cota (ICM_SLAM.py:255)")
        if rc != 0:
            raise RuntimeError("orc_filtrar rc=%d" % rc)
        return out


def calc_cambio(y, mapa_viejo):
    y, old = _c(y), _c(mapa_viejo)
    out = np.empty(3)
    lib().orc_calc_cambio(_d(y), int(y.shape[1]), C.c_long(y.shape[1]), _d(old), int(old.shape[1]),
                          C.c_long(old.shape[1]), _d(out))
    return tuple(out)


def fun_xn(cfg, x, x_ant, x_pos, u_ant, u_act, odo3, d, alpha, sx, sy):
    a = [_c(v).reshape(-1) for v in (x, x_ant, x_pos, u_ant, u_act)]
    o = _c(odo3)  # 3x3: columns t-1, t, t+1
    oc = [np.ascontiguousarray(o[:, j]) for j in range(3)]
    d, alpha, sx, sy = _c(d), _c(alpha), _c(sx), _c(sy)
    return lib().orc_fun_xn(C.byref(cfg), *[_d(v) for v in a], *[_d(v) for v in oc], int(d.size), _d(d), _d(alpha),
                            _d(sx), _d(sy))


def fun_x(cfg, x, x_ant, u_ant, odo2, d, alpha, sx, sy):
    a = [_c(v).reshape(-1) for v in (x, x_ant, u_ant)]
    o = _c(odo2)
    oc = [np.ascontiguousarray(o[:, j]) for j in range(2)]
    d, alpha, sx, sy = _c(d), _c(alpha), _c(sx), _c(sy)
    return lib().orc_fun_x(C.byref(cfg), *[_d(v) for v in a], *[_d(v) for v in oc], int(d.size), _d(d), _d(alpha),
                           _d(sx), _d(sy))


def solve_pose(cfg, solver, x_ant, x_pos, u_ant, u_act, odo3, d, alpha, sx, sy):
    """x_pos None -> causal problem (fun_x).  Returns (pose[3], n_energy_evals)."""
    has_next = x_pos is not None
    z3, z2 = np.zeros(3), np.zeros(2)
    a = [_c(x_ant).reshape(-1), _c(x_pos).reshape(-1) if has_next else z3, _c(u_ant).reshape(-1),
         _c(u_act).reshape(-1) if has_next else z2]
    o = _c(odo3)
    oc = [np.ascontiguousarray(o[:, j]) for j in range(o.shape[1])]
    if len(oc) < 3:
        oc.append(z3)
    d, alpha, sx, sy = _c(d), _c(alpha), _c(sx), _c(sy)
    out = np.empty(3)
    nev = lib().orc_solve_pose(C.byref(cfg), SOLVER[solver], int(has_next), *[_d(v) for v in a],
                               *[_d(v) for v in oc], int(d.size), _d(d), _d(alpha), _d(sx), _d(sy), _d(out))
    return out, int(nev)


def sweep(cfg: OrcConfig, mapa: Mapa, ext: dict, odo, u, x0, map_in, x, schedule="sequential", solver="nm",
          view="running", newton_tol=1e-14):
    """iterations_process_offline restated.  x (3xT float64 C-contiguous) is updated IN PLACE and
    returned, like the reference.  Returns dict(map, x, c, seen, raw_map, raw_counts, raw_L, nev, status)."""
    assert x.flags.c_contiguous and x.dtype == np.float64
    T = x.shape[1]
    odo, u = _c(odo), _c(u)
    x0 = _c(x0).reshape(3)
    map_in = _c(map_in)
    Lin = map_in.shape[1]
    n = ext["n"]
    c = np.full(max(n, 1), -9, np.int32)
    seen_x = np.zeros(max(n, 1))
    seen_y = np.zeros(max(n, 1))
    L = int(cfg.L)
    raw_map = np.zeros((2, L))
    raw_counts = np.zeros(L)
    raw_L = C.c_int32(0)
    map_out = np.zeros((2, L))
    nev = C.c_long(0)
    rc = lib().orc_sweep(C.byref(cfg), mapa._h, int(T), _i(ext["off"]), _i(ext["beam"]), _d(ext["d"]), _d(ext["bx"]),
                         _d(ext["by"]), _d(ext["ang"]), _d(odo), C.c_long(odo.shape[1]), _d(u), C.c_long(u.shape[1]),
                         _d(x0), _d(map_in), int(Lin), C.c_long(max(Lin, 1)), _d(x), C.c_long(T), SCHED[schedule],
                         SOLVER[solver], VIEW[view], C.c_double(newton_tol), _i(c), _d(seen_x), _d(seen_y),
                         _d(raw_map), _d(raw_counts), C.byref(raw_L), _d(map_out), C.byref(nev))
    if rc == 1:
        return dict(map=map_in.copy(), x=x, status=1)
    if rc == -2:
        raise IndexError("label capacity L exceeded (ICM_SLAM.py:191)")
    if rc == -3:
        raise IndexError("last scan has no observations (sensors.py:148)")
    if rc == -4:
        raise ValueError("no landmark survives cota (ICM_SLAM.py:255)")
    if rc != 0:
        raise RuntimeError("orc_sweep rc=%d" % rc)
    La = mapa.landmarks_actuales
    return dict(map=map_out[:, :La].copy(), x=x, c=c[:n], seen=np.stack([seen_x[:n], seen_y[:n]]),
                raw_map=raw_map[:, : raw_L.value].copy(), raw_counts=raw_counts[: raw_L.value].copy(),
                raw_L=int(raw_L.value), nev=int(nev.value), status=0,
                counts=mapa.cant_obs_i[:La].copy())
