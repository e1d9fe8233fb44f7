"""Row f2 of SURVEY.md section 8: the step either side of the hot path for the offline batch run -- `.mat` ingestion
(both layouts the reference ships), the `filtrar_obs.m` pre-pass on the GPU, pass 0, N device-resident sweeps with the
`calc_cambio` history (sensors.py:302-315 / example.py:37-54), and writers for the poses / landmarks / history.

    res = run_offline("datos_palomar1.mat", config)            # raw log: filtrar_obs first
    save_result("out.mat", res)                                # or .npz

Everything numeric runs in libicmslam (CUDA); this module only moves arrays and files.
"""
from __future__ import annotations

import os

import numpy as np

from .config import ConfigICM
from .icm import ICM_SLAM, Mapa, calc_cambio, load_mat, precondicionar


def prepare_scans(z, config, raw=None, cant_max=15):
    """Range scans (B x T) as a log holds them -> the pre-conditioned ranges the solver loads.  `raw` (default: decided
    from the data, any range above rango_laser_max) selects the `filtrar_obs.m` pre-pass (filtrar_obs.m:6-50: at most
    `cant_max` nearest valid beams per scan, interpolated through the dense scans), run on the GPU; then
    sensors_definitions.py:20-22 (NaN -> max, + radio, clamp)."""
    z = np.asarray(z, dtype=np.float64)
    if raw is None:
        raw = bool(np.nanmax(z) > config.rango_laser_max)
    if raw:
        from .icm import _scratch_engine
        z = _scratch_engine(config).filtrar_obs(z, float(config.rango_laser_max), int(cant_max))
    return precondicionar(z, config)


def run_offline(data, config, N=None, x0=None, raw=None, history=True):
    """The reference's offline driver in one call.  `data` is a `.mat` path or a (z, odometria, u) triple.
    Returns a dict: x (3 x T), mapa (2 x L), cambios (N x 3: min / max / mean landmark change per sweep, ICM_SLAM.py:490-495),
    mapa_inicial, x_inicial, landmarks_actuales, labels (association index of every kept beam in the last sweep)."""
    if isinstance(data, (str, os.PathLike)):
        z, odo, u = load_mat(data)
    else:
        z, odo, u = data
    odo = np.ascontiguousarray(odo, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    zc = prepare_scans(z, config, raw=raw)
    icm = ICM_SLAM(config, x0=odo[:, 0] if x0 is None else x0)
    icm.load_data(Mapa(config), zc, u, odo)
    mapa0, xs0 = icm.inicializar()
    mapa0, xs0 = np.array(mapa0), np.array(xs0)
    n_sweeps = int(config.N if N is None else N)
    eng = icm.engine
    x0v = np.asarray(icm.x0, dtype=np.float64).reshape(3)
    mode = dict(schedule=getattr(config, "schedule", "redblack"), solver=getattr(config, "solver", "newton"),
                view=getattr(config, "map_view", "prev"))
    eng.set_map(mapa0)
    eng.set_poses(xs0)
    cambios = []
    mapa_prev = mapa0
    for _ in range(n_sweeps):
        eng.iterate(None, x0v, 1, **mode)              # poses and map stay on the device
        if history:
            mapa_now = eng.get_map()
            cambios.append(calc_cambio(mapa_now, mapa_prev, config))
            mapa_prev = mapa_now
    x = eng.get_poses()
    mapa = eng.get_map()
    icm.mapa_viejo, icm.positions = mapa.copy(), x.copy()
    return dict(x=x, mapa=mapa, cambios=np.array(cambios, dtype=np.float64).reshape(-1, 3), mapa_inicial=mapa0, x_inicial=xs0,
                landmarks_actuales=int(mapa.shape[1]), labels=eng.associations(), icm=icm)


def save_result(path, res):
    """Poses (3 x T), landmarks (2 x L) and the calc_cambio history as `.mat` (scipy.io.savemat, the format the reference's
    MATLAB tooling reads) or `.npz`, by extension."""
    out = {k: np.asarray(res[k]) for k in ("x", "mapa", "cambios", "mapa_inicial", "x_inicial", "labels") if k in res}
    if str(path).endswith(".mat"):
        import scipy.io as sio
        sio.savemat(path, out)
    else:
        np.savez_compressed(path, **out)
    return path


def load_result(path):
    if str(path).endswith(".mat"):
        import scipy.io as sio
        d = sio.loadmat(path)
        return {k: np.array(v) for k, v in d.items() if not k.startswith("__")}
    return dict(np.load(path))
