"""The reference's operator surface for the offline batch path, backed by libicmslam.so.

Names, argument meaning and error behaviour follow the reference so that its drivers
(sensors.py:284-320, example.py:37-54, external_options.py:37-92) run unchanged on top of this
module:

* `ICM_SLAM(config, x0='')`  -- the solver.  The reference has no class of this name (ICM_SLAM is
  its tools *module*); the solver class is `ICM_ROS` (sensors.py:15) and, for the legacy offline
  API, `ICM_method` (ICM_SLAM_old.py:59).  Both names are aliases of this class:
  `iterations_process_offline(mapa_viejo, x)` (sensors.py:125-168) and
  `load_data / itererar` (ICM_SLAM_old.py:249, :336).
* `Mapa(config)` -- `landmarks_actuales`, `cant_obs_i`, `clear_obs()`, `filtrar(mapa)`
  (ICM_SLAM.py:104-265).
* `filtrar_z(z, config)`, `tras_rot_z`, `calc_cambio`, `entrepi`, `Rota` (ICM_SLAM.py:22-58, :455-495).

The sweep, extraction, association, map filter, pass 0 and calc_cambio run in the CUDA library through the C ABI
(include/icmslam.h); there is no CPU fallback -- without the built library or without a GPU these calls raise.  The
four geometry helpers `entrepi`, `Rota`, `tras_rot_z`, `precondicionar` are host numpy one-liners, as in the reference
(a handful of flops on one pose / one array; the sweep kernels have their own device versions).

DEFAULT MODE.  `ICM_SLAM(config)` runs the restated parallel sweep (red-black schedule, exact Newton solve, previous-map
view: config.schedule / solver / map_view = 'redblack' / 'newton' / 'prev'), which converges to the same fixed point as the
reference but is NOT the reference's iteration: a reference YAML gives the reference's own poses sweep by sweep only with
schedule='sequential', solver='nm', map_view='running' (DESIGN.md section 2).
"""
from __future__ import annotations

from copy import copy

import numpy as np

from .config import ConfigICM
from .engine import Engine

__all__ = ["ICM_SLAM", "ICM_ROS", "ICM_method", "Mapa", "ConfigICM", "filtrar_z", "tras_rot_z", "calc_cambio", "entrepi",
           "Rota", "precondicionar", "load_mat"]


# ---- geometry helpers (host-side conveniences of ICM_SLAM.py:455-488; a handful of flops) --------
def entrepi(angulo):
    """ICM_SLAM.py:455-463."""
    angulo = np.mod(angulo, 2 * np.pi)
    if angulo > np.pi:
        angulo = angulo - 2 * np.pi
    return angulo


def Rota(theta):
    """ICM_SLAM.py:482-488."""
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, s], [-s, c]])


def tras_rot_z(x, z):
    """ICM_SLAM.py:465-480: body -> world, columns 2:4 of z updated IN PLACE like the reference."""
    x = np.asarray(x, dtype=np.float64).reshape(3)
    th = x[2] - np.pi / 2.0
    c, s = np.cos(th), np.sin(th)
    R = np.array([[c, s], [-s, c]])
    z[:, 2:4] = np.matmul(z[:, 2:4], R) + x[0:2]
    return z


def precondicionar(z, config):
    """sensors_definitions.py:21-22 / IJAC2018_python.txt:43: NaN -> max, z = min(z + radio, max)."""
    z = np.array(z, dtype=np.float64)
    z[np.isnan(z)] = config.rango_laser_max
    return np.minimum(z + config.radio, config.rango_laser_max)


def load_mat(path):
    """Both `.mat` layouts the reference ships: data_IJAC2018.mat
    (`observations`, `odometry`, `velocities`; createbag.py:124-127) and datos_palomar1.mat
    (struct `datos` with `observaciones`, `odometria`, `control`).  Returns (z, odometria, u)."""
    import scipy.io as sio
    d = sio.loadmat(path)
    if "datos" in d:
        s = d["datos"][0, 0]
        return (np.array(s["observaciones"], dtype=np.float64), np.array(s["odometria"], dtype=np.float64),
                np.array(s["control"], dtype=np.float64))
    return (np.array(d["observations"], dtype=np.float64), np.array(d["odometry"], dtype=np.float64),
            np.array(d["velocities"], dtype=np.float64))


_scratch = {}


def _scratch_engine(config, L=None):
    key = (int(getattr(config, "device", 0)), float(config.dist_thr), float(config.rango_laser_max), float(config.cota),
           int(L if L is not None else config.L))
    e = _scratch.get(key)
    if e is None:
        e = Engine(config, device=key[0], L=key[4])
        _scratch[key] = e
    return e


def filtrar_z(z, config):
    """ICM_SLAM.py:22-58 for one scan (a length-B column, already pre-conditioned).  Returns the
    (n, 4) array [d, angle, d cos, d sin]; an empty scan gives shape (0,) or (0, 4) exactly as the
    reference does (np.array([]) when fewer than two beams pass the range gate)."""
    z = np.asarray(z, dtype=np.float64).reshape(-1, 1)
    e = _scratch_engine(config)
    e.load(z, np.zeros((3, 1)), np.zeros((2, 1)), precondition=False)
    e.extract()
    g = e.get_extraction()
    if e.n == 0:
        nvalid = int((_median3(z[:, 0]) < config.rango_laser_max).sum())
        return np.array([]) if nvalid <= 1 else np.zeros((0, 4))
    return np.stack([g["d"], g["beam"] * np.pi / 180.0, g["bx"], g["by"]], axis=1)


def _median3(z):
    p = np.concatenate([[0.0], z, [0.0]])
    return np.median(np.stack([p[:-2], p[1:-1], p[2:]]), axis=0)


def calc_cambio(y, mapa_viejo, config=None):
    """ICM_SLAM.py:490-495 -> [min, max, mean]."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    old = np.ascontiguousarray(mapa_viejo, dtype=np.float64)
    cfg = config if config is not None else ConfigICM.from_values(L=max(int(y.shape[1]), int(old.shape[1]), 1))
    e = _scratch_engine(cfg, L=max(int(y.shape[1]), int(old.shape[1]), int(cfg.L)))
    mn, mx, me = e.calc_cambio(y, old)
    return [mn, mx, me]


class Mapa:
    """ICM_SLAM.py:104-265.  `landmarks_actuales` and `cant_obs_i` live in a device handle: the solver's once the Mapa is
    attached to one (ICM_SLAM.load_data / mapa_obj), else a private handle created on first use."""

    def __init__(self, config):
        self.L = config.L
        self.cota = config.cota
        self.dist_thr = config.dist_thr
        self._config = config
        self._engine = None
        self._own = None
        self._lact = 0

    def _eng(self):
        if self._engine is not None:
            return self._engine
        if self._own is None:
            self._own = Engine(self._config, device=int(getattr(self._config, "device", 0)))
            self._own.landmarks_actuales = self._lact
        return self._own

    def _attach(self, engine):
        if self._own is not None:                  # carry the state over
            self._lact = self._own.landmarks_actuales
            cant = self._own.counts(self.L)
            self._own.close()
            self._own = None
            engine.set_counts(cant)
        self._engine = engine
        engine.landmarks_actuales = self._lact

    @property
    def landmarks_actuales(self):
        if self._engine is None and self._own is None:
            return self._lact
        return self._eng().landmarks_actuales

    @landmarks_actuales.setter
    def landmarks_actuales(self, v):
        self._lact = int(v)
        if self._engine is not None or self._own is not None:
            self._eng().landmarks_actuales = int(v)

    @property
    def cant_obs_i(self):
        if self._engine is None and self._own is None:
            return np.zeros(self.L)
        return self._eng().counts(min(self.L, self._eng().L))

    def clear_obs(self):
        """ICM_SLAM.py:119-126."""
        if self._engine is not None or self._own is not None:
            self._eng().set_counts(None)

    def actualizar(self, mapa, mapa_referencia, obs):
        """ICM_SLAM.py:128-201 for one scan: nearest landmark of `mapa_referencia` (its first landmarks_actuales columns) for
        every row of `obs` (n, 2), gate dist_thr, one new label for the scan's far observations, running means written into
        `mapa` (2 x L, in place), counts into cant_obs_i.  Returns (mapa, c) like the reference.  The very first call of a
        fresh Mapa (landmarks_actuales == 0) clusters the scan as scipy's fcluster does (:160-165)."""
        if not (isinstance(mapa, np.ndarray) and mapa.dtype == np.float64 and mapa.ndim == 2 and mapa.strides[1] == 8):
            raise TypeError("mapa must be a (2, L) float64 array (it is updated in place, like the reference's)")
        c = self._eng().associate(mapa, mapa_referencia, obs)
        return mapa, c

    def filtrar(self, mapa, cant_obs_i=None):
        """ICM_SLAM.py:204-265 on a 2 x L map; counts default to the handle's (the last sweep's / actualizar's).
        Returns the 2 x L buffer (caller slices [:, :landmarks_actuales]) like the reference."""
        e = self._eng()
        mapa = np.ascontiguousarray(mapa, dtype=np.float64)
        cnt = np.ascontiguousarray(self.cant_obs_i if cant_obs_i is None else cant_obs_i, dtype=np.float64)
        n = min(mapa.shape[1], cnt.shape[0], self.landmarks_actuales if cant_obs_i is None else cnt.shape[0])
        out, cout, Lout = e.filter_map(mapa[:, :n], cnt[:n])
        self._lact = Lout
        return out


class ICM_SLAM:
    """Offline ICM-SLAM solver on one B200.  See the module docstring for the name mapping."""

    def __init__(self, config, x0=""):
        if isinstance(x0, str) and x0 == "":
            self.x0 = np.zeros((3, 1))          # sensors.py:19-22
        else:
            self.x0 = np.asarray(x0, dtype=np.float64).reshape(3, 1)
        self.config = config
        self.odometria = np.array([])
        self.mediciones = np.array([])
        self.u = np.array([])
        self.iterations_flag = False
        self.debug = False
        self.mapa_obj = None
        self.mapa_viejo = None
        self.positions = None
        self._engine = Engine(config, device=int(getattr(config, "device", 0)))
        self._loaded = None
        self._attached = None

    # ---- data ---------------------------------------------------------------------------------
    def load_data(self, mapa_obj, mediciones, u, odometria, x0=""):
        """ICM_SLAM_old.py:249-264.  `mediciones` are pre-conditioned ranges (B x T), `u` 2 x T,
        `odometria` 3 x T -- the arrays ROS.principal_callback accumulates (ICM_SLAM.py:332-339)."""
        self.mediciones = mediciones
        self.u = u
        self.mapa_obj = copy(mapa_obj)
        self.odometria = odometria
        if not (isinstance(x0, str) and x0 == ""):
            self.x0 = np.asarray(x0, dtype=np.float64).reshape(3, 1)
        elif odometria is not None and np.size(odometria):
            pass   # the legacy API keeps zeros; ICM_ROS sets x0 = odometria[:,0] in inicializar_online (sensors.py:61)
        self._sync_data()

    @staticmethod
    def _fingerprint(a):
        """Identity of an input array: where it lives, its shape and strides, and its four corner values.  The reference
        REPLACES its logs when they grow (np.concatenate, sensors.py:266-270), which this catches; after in-place edits
        call `invalidate()`.  (A few microseconds per call: it runs on every sweep.)"""
        a = np.asarray(a)
        if a.size == 0:
            return (0,)
        corners = (float(a.flat[0]), float(a.flat[-1])) if a.ndim != 2 else (float(a[0, 0]), float(a[0, -1]), float(a[-1, 0]), float(a[-1, -1]))
        return (a.__array_interface__["data"][0], a.shape, a.strides) + corners

    def _data_key(self):
        return (self._fingerprint(self.mediciones), self._fingerprint(self.u), self._fingerprint(self.odometria))

    def adopt_engine(self, engine):
        """Use an Engine that already holds (mediciones, odometria, u) -- loaded and extracted through the Engine API --
        instead of uploading the dataset again."""
        if self._engine is not engine:
            self._engine.close()
            self._engine = engine
        self._loaded = self._data_key()
        if self.mapa_obj is None:
            self.mapa_obj = Mapa(self.config)
        self.mapa_obj._attach(engine)
        self._attached = self.mapa_obj

    def invalidate(self):
        """Forget the device copy of the dataset: the next call uploads and re-extracts it (after in-place edits of
        `mediciones` / `u` / `odometria`, e.g. appending into a preallocated buffer)."""
        self._loaded = None

    def _sync_data(self):
        key = self._data_key()
        if self._loaded != key:
            self._engine.load(self.mediciones, self.odometria, self.u, precondition=False)
            self._engine.extract()
            self._loaded = key
            self._attached = None
        if self.mapa_obj is None:
            self.mapa_obj = Mapa(self.config)
        if self._attached is not self.mapa_obj:       # a Mapa the caller replaced (or a fresh dataset): attach it to the engine
            self.mapa_obj._attach(self._engine)
            self._attached = self.mapa_obj

    # ---- the model functions the reference leaves to the user (sensors.py:170-282) ----------------------------------------
    # Same names, same attribute protocol (self.t, self.xt, self.x_ant, self.x_pos, self.medicion_actual, self.mapa_visto, self.u,
    # self.odometria), evaluated on the device through icmslam_pose_eval.  The sweep kernels implement exactly this model
    # (unicycle + 2-D laser, ICMSLAM_MODEL_UNICYCLE_LASER2D); overriding these methods in a subclass does NOT change the sweep.
    def g(self, xt, ut):
        """sensors.py:206-211."""
        x, _, _ = self._engine.pose_eval("g", x_ant=np.asarray(xt, dtype=np.float64).reshape(3), u_ant=np.asarray(ut, dtype=np.float64).reshape(2))
        return x.reshape((3, 1))

    def h(self, xt, zt):
        """sensors.py:175-204: observation potential of pose xt for zt = [range, beam angle] rows against self.mapa_visto."""
        zt = np.asarray(zt, dtype=np.float64)
        _, f, _ = self._engine.pose_eval("h", z=zt[:, 0:2], seen=self.mapa_visto, x=np.asarray(xt, dtype=np.float64).reshape(3))
        return f

    def _pose_args(self, with_next):
        t = int(self.t)
        z = np.asarray(self.medicion_actual, dtype=np.float64)
        kw = dict(z=z[:, 0:2], seen=self.mapa_visto, x_ant=np.asarray(self.xt, dtype=np.float64).reshape(3),
                  u_ant=np.asarray(self.u)[:, t - 1])
        if with_next:
            kw.update(x_pos=np.asarray(self.x_pos, dtype=np.float64).reshape(3), u_act=np.asarray(self.u)[:, t],
                      odo=np.asarray(self.odometria)[:, t - 1:t + 2])
        else:
            kw.update(odo=np.asarray(self.odometria)[:, t - 1:t + 1])
        return kw

    def fun_xn(self, x):
        """sensors.py:224-255."""
        return self._engine.pose_eval("energy", x=x, **self._pose_args(True))[1]

    def fun_x(self, x):
        """sensors.py:266-282."""
        return self._engine.pose_eval("energy", x=x, **self._pose_args(False))[1]

    def minimizar_xn(self, medicion_actual, mapa_visto, x, t):
        """sensors.py:213-222: Nelder-Mead on fun_xn from (x_ant + x_pos) / 2."""
        x = np.asarray(x, dtype=np.float64)
        self.x_ant = x[:, t - 1].reshape((3, 1))
        self.x_pos = x[:, t + 1].reshape((3, 1))
        self.xt = x[:, t - 1].reshape((3, 1))
        self.t = t
        self.medicion_actual = medicion_actual
        self.mapa_visto = mapa_visto
        solver = "min_nm" if getattr(self.config, "solver", "newton") == "nm" else "min_newton"
        return self._engine.pose_eval(solver, **self._pose_args(True))[0]

    def minimizar_x(self, medicion_actual, mapa_visto):
        """sensors.py:257-264: Nelder-Mead on fun_x from g(xt, u[t-1]) (self.xt and self.t as the caller left them)."""
        self.medicion_actual = medicion_actual
        self.mapa_visto = mapa_visto
        solver = "min_nm" if getattr(self.config, "solver", "newton") == "nm" else "min_newton"
        return self._engine.pose_eval(solver, **self._pose_args(False))[0]

    # ---- the sweep ------------------------------------------------------------------------------
    def iterations_process_offline(self, mapa_viejo, x):
        """sensors.py:125-168: one ICM sweep.  `x` (3 x T) is updated IN PLACE and returned;
        `mapa_viejo` (2 x L) is not modified; returns (mapa_refinado (2 x L'), x).  The sweep variant
        is config.schedule / solver / map_view; ('sequential', 'nm', 'running') is the
        reference's own semantics."""
        self._sync_data()
        cfg = self.config
        mapa_viejo = np.ascontiguousarray(mapa_viejo, dtype=np.float64)
        xin = x
        xc = x if (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous) else np.ascontiguousarray(x, dtype=np.float64)
        st, Lout, mapa = self._engine.sweep(mapa_viejo, xc, np.asarray(self.x0).reshape(3), schedule=getattr(cfg, "schedule", "redblack"),
                                            solver=getattr(cfg, "solver", "newton"), view=getattr(cfg, "map_view", "prev"))
        if xc is not xin:
            xin[...] = xc
        if st == 1:   # empty first scan: inputs returned unchanged (sensors.py:137-139)
            return mapa_viejo, xin
        return mapa, xin          # (a fresh array owned by the caller, like the reference's deepcopy, sensors.py:167)

    itererar = iterations_process_offline          # ICM_SLAM_old.py:336

    def iterar(self, x, N=None):
        """The driver loop (sensors.py:302-315) with the map resident on the device: N sweeps starting
        from self.mapa_viejo; returns (mapa_refinado, x)."""
        self._sync_data()
        cfg = self.config
        self._engine.set_map(self.mapa_viejo)
        self._engine.iterate(x, np.asarray(self.x0).reshape(3), int(cfg.N if N is None else N),
                             schedule=getattr(cfg, "schedule", "redblack"), solver=getattr(cfg, "solver", "newton"),
                             view=getattr(cfg, "map_view", "prev"))
        self.mapa_viejo = self._engine.get_map()
        return self.mapa_viejo, x

    def inicializar(self, x=None):
        """Pass 0, the causal initialisation (ICM_method.inicializar, ICM_SLAM_old.py:266-333; the same computation as
        ICM_ROS.inicializar_online + inicializar_online_process, sensors.py:51-123, replayed on the loaded log): returns
        (mapa_inicial, x) and sets `mapa_viejo`, `positions` and the Mapa state like the reference does."""
        self._sync_data()
        xs, mapa = self._engine.pass0(np.asarray(self.x0).reshape(3))
        if x is not None and getattr(x, "shape", None) == xs.shape:
            x[...] = xs
            xs = x
        self.mapa_viejo = mapa.copy()
        self.positions = xs.copy()
        self.iterations_flag = True
        return mapa, xs

    def inicializar_online(self):
        """sensors.py:51-104 without the ROS wait loop: x0 = odometria[:,0], then pass 0 over the log."""
        self.x0 = np.array([np.asarray(self.odometria)[:, 0]]).T          # sensors.py:61
        self.inicializar()

    @property
    def engine(self):
        return self._engine

    def associations(self):
        return self._engine.associations()


ICM_ROS = ICM_SLAM
ICM_method = ICM_SLAM
