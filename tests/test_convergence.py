"""Convergence-level comparison with the reference.  tests/golden/c1_converged.npz holds the UNMODIFIED
reference's result of config_ros.yaml's N = 30 sweeps on data_IJAC2018.mat after its own pass 0 (oracle/make_golden_convergence.py;
~12 minutes of the reference on one core).

The default mode (red-black schedule, exact Newton solve, previous-map view) is a different iteration from the reference's
(sequential Gauss-Seidel, Nelder-Mead stopped at xtol = 1e-3, running means): both descend the same joint energy
E(x, y) = sum_t fun_x(x_t | x_{t-1}, landmarks) (sensors.py:258-282: every motion / odometry term once, every observation once),
but ICM moves slowly along the chain and the reference's inner solver stops up to centimetres short of each conditional minimum
(SURVEY.md section 0), so after 30 sweeps the two are at different points of the descent, not at a common fixed point:
E(pass 0) = 285.00, E(reference, 30 sweeps) = 283.23, E(fast mode, 30 sweeps) = 270.60 (CPU oracle in the same mode).
What is asserted:
  * the fast mode finds the SAME 11 landmarks: each within TOL_LM = dist_thr / 4 = 0.25 m of the reference's (landmarks are
    >= 2 m apart on this log, so the correspondence is unambiguous; measured: <= 0.155 m), poses within 0.25 m / 0.1 rad
    (measured: 0.14 m / 0.04 rad);
  * its joint energy, evaluated by the oracle's restatement of the reference's own fun_x, is not above the reference's;
  * the GPU's 30 sweeps equal the CPU oracle's 30 sweeps in the same mode to 1e-6 m / 1e-8 rad (same-mode parity at depth);
  * the reference's OWN mode on the GPU (sequential / NM / running) equals the unmodified reference after all 30 sweeps to
    1e-6 m / 1e-8 rad (the north_star's tolerance, at depth).
"""
import numpy as np
import pytest

from helpers import CONFIG_ROS, c1_inputs, golden

pytestmark = pytest.mark.gpu

TOL_LM = 0.25


def _energy(orc, ocfg, ext, odo, u, x, mapa):
    off = ext["off"]
    E = 0.0
    for t in range(1, x.shape[1]):
        a, b = off[t], off[t + 1]
        d, al = ext["d"][a:b], ext["ang"][ext["beam"][a:b]]
        sx = sy = np.zeros(0)
        if b > a:
            wx, wy = orc.tras_rot(x[:, t], ext["bx"][a:b], ext["by"][a:b])
            D = np.hypot(mapa[0][:, None] - wx[None, :], mapa[1][:, None] - wy[None, :])
            c, ok = D.argmin(0), D.min(0) <= ocfg.dist_thr
            d, al, sx, sy = d[ok], al[ok], mapa[0][c[ok]], mapa[1][c[ok]]
        E += orc.fun_x(ocfg, x[:, t], x[:, t - 1], u[:, t - 1], odo[:, t - 1:t + 1], d, al, sx, sy)
    return E


def test_fast_mode_vs_reference_after_30_sweeps():
    from icm_slam_b200.config import ConfigICM
    from icm_slam_b200.engine import Engine
    from oracle import oracle as orc
    g = golden("c1_converged.npz")
    n = int(g["nsweeps"])
    z, odo, u = c1_inputs()
    ocfg = orc.make_cfg(**CONFIG_ROS)
    ext = orc.extract_all(orc.precondition(z, ocfg.radio, ocfg.rango_laser_max), ocfg)
    e = Engine(ConfigICM.from_values(**CONFIG_ROS))
    e.load(z, odo, u, precondition=True)
    e.extract()
    e.set_map(g["p0_map"])
    e.set_poses(np.ascontiguousarray(g["p0_x"].copy()))
    e.iterate(None, odo[:, 0], n)
    xg, mg = e.get_poses(), e.get_map()
    e.close()
    xr, mr = g["x_ref"], g["map_ref"]
    assert mg.shape == mr.shape == (2, 11)
    assert np.hypot(*(mg - mr)).max() <= TOL_LM
    assert np.abs(xg[:2] - xr[:2]).max() <= 0.25 and np.abs(xg[2] - xr[2]).max() <= 0.1
    E_fast, E_ref = _energy(orc, ocfg, ext, odo, u, xg, mg), _energy(orc, ocfg, ext, odo, u, xr, mr)
    assert E_fast <= E_ref, (E_fast, E_ref)
    # same-mode parity at depth: the CPU oracle's 30 sweeps
    m = orc.Mapa(ocfg)
    mp = g["p0_map"].copy()
    m.landmarks_actuales = mp.shape[1]
    xo = np.ascontiguousarray(g["p0_x"].copy())
    for _ in range(n):
        mp = orc.sweep(ocfg, m, ext, odo, u, odo[:, 0], mp, xo, "redblack", "newton", "prev")["map"]
    assert np.abs(xg[:2] - xo[:2]).max() <= 1e-6 and np.abs(xg[2] - xo[2]).max() <= 1e-8
    assert mg.shape == mp.shape and np.abs(mg - mp).max() <= 1e-6


def test_reference_mode_tracks_the_reference_through_30_sweeps():
    from icm_slam_b200.config import ConfigICM
    from icm_slam_b200.engine import Engine
    g = golden("c1_converged.npz")
    n = int(g["nsweeps"])
    z, odo, u = c1_inputs()
    e = Engine(ConfigICM.from_values(**CONFIG_ROS))
    e.load(z, odo, u, precondition=True)
    e.extract()
    e.set_map(g["p0_map"])
    e.set_poses(np.ascontiguousarray(g["p0_x"].copy()))
    e.iterate(None, odo[:, 0], n, schedule="sequential", solver="nm", view="running")
    xg, mg = e.get_poses(), e.get_map()
    e.close()
    assert mg.shape == g["map_ref"].shape
    dm, dx, dth = np.abs(mg - g["map_ref"]).max(), np.abs(xg[:2] - g["x_ref"][:2]).max(), np.abs(xg[2] - g["x_ref"][2]).max()
    print("reference mode vs reference after %d sweeps: map %.3e m, poses %.3e m / %.3e rad" % (n, dm, dx, dth))
    # (measured on B200: map 3e-15 m, poses 0 m / 7e-14 rad -- every Nelder-Mead decision of 30 x 1832 solves reproduced)
    assert dm <= 1e-6 and dx <= 1e-6 and dth <= 1e-8
