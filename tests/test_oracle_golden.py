"""Pins the CPU oracle (oracle/icm_oracle.c) against fixtures minted by the UNMODIFIED reference
(oracle/make_golden.py).  The reference has no tests of its own; these
fixtures are what "identical to the reference" means for every later GPU parity test.
"""
import numpy as np
import pytest

from helpers import CONFIG_ROS, c1_inputs, c2_inputs, golden
from oracle import oracle as orc

U = None


def units():
    global U
    if U is None:
        U = golden("units.npz")
    return U


def cfg_ros(**kw):
    d = dict(CONFIG_ROS)
    d.update(kw)
    return orc.make_cfg(**d)


# ------------------------------------------------------------------ a1: filtrar_obs.m
def test_filtrar_obs_reproduces_dataset_pair():
    """datos_palomar1.mat --filtrar_obs.m--> data_IJAC2018.mat, bit for bit."""
    raw, _, _ = c2_inputs()
    want, _, _ = c1_inputs()
    got, a = orc.filtrar_obs(raw, 10.0, 15)
    assert np.array_equal(got, want)
    assert a.max() <= 15


# ------------------------------------------------------------------ a3: filtrar_z
def test_filtrar_z_units_bit_exact():
    u = units()
    cfg = cfg_ros()
    ext = orc.extract_all(u["fz_scans"], cfg)
    assert np.array_equal(ext["off"], u["fz_off"])
    rows = u["fz_rows"]
    assert np.array_equal(ext["d"], rows[:, 0])
    assert np.array_equal(ext["ang"][ext["beam"]], rows[:, 1])
    assert np.array_equal(ext["bx"], rows[:, 2])
    assert np.array_equal(ext["by"], rows[:, 3])
    # the crafted cases really cover the empty outcomes
    nt = np.diff(ext["off"])
    assert (nt == 0).sum() >= 10 and nt.max() > 100


def test_filtrar_z_c1_counts():
    """17 892 kept beams, 100 empty scans, max 17 per scan."""
    z, _, _ = c1_inputs()
    cfg = cfg_ros()
    ext = orc.extract_all(orc.precondition(z, cfg.radio, cfg.rango_laser_max), cfg)
    nt = np.diff(ext["off"])
    assert ext["n"] == 17892 and (nt == 0).sum() == 100 and nt.max() == 17


# ------------------------------------------------------------------ a4: tras_rot_z
def test_tras_rot_units():
    u = units()
    for p, zi, zo in zip(u["tr_poses"], u["tr_in"], u["tr_out"]):
        wx, wy = orc.tras_rot(p, zi[:, 2], zi[:, 3])
        # bit-exact: numpy's matmul fuses the second product (acc = fma(a1, b1, a0*b0))
        assert np.array_equal(wx, zo[:, 2]) and np.array_equal(wy, zo[:, 3])


# ------------------------------------------------------------------ a5: Mapa.actualizar
def test_actualizar_units():
    u = units()
    cfg = cfg_ros(L=60)
    for q in range(int(u["ac_nseq"])):
        ref = u[f"ac{q}_ref"]
        m = orc.Mapa(cfg)
        m.landmarks_actuales = ref.shape[1]
        y = np.zeros((2, 60))
        for r in range(int(u[f"ac{q}_ncall"])):
            y, c = m.actualizar(y, ref, u[f"ac{q}_{r}_obs"])
            assert np.array_equal(c, u[f"ac{q}_{r}_c"]), (q, r)
            assert m.landmarks_actuales == int(u[f"ac{q}_{r}_Lact"])
            assert np.array_equal(m.cant_obs_i, u[f"ac{q}_{r}_cant"])
            assert np.max(np.abs(y - u[f"ac{q}_{r}_y"])) <= 1e-13


# ------------------------------------------------------------------ a6: Mapa.filtrar
def test_filtrar_units():
    u = units()
    cfg = cfg_ros(L=40, cota=10.0)
    nok = 0
    for q in range(int(u["fl_n"])):
        if not int(u[f"fl{q}_ok"]):
            continue
        nok += 1
        pts, cnt = u[f"fl{q}_in"], u[f"fl{q}_cnt"]
        m = orc.Mapa(cfg)
        La = pts.shape[1]
        m.landmarks_actuales = La
        m.cant_obs_i[:La] = cnt
        y = np.zeros((2, 40))
        y[:, :La] = pts
        out = m.filtrar(y)
        assert m.landmarks_actuales == int(u[f"fl{q}_Lact"]), q
        assert np.array_equal(m.cant_obs_i, u[f"fl{q}_cant"]), q
        assert np.max(np.abs(out - u[f"fl{q}_out"])) <= 1e-12, q
    assert nok >= 8


def test_calc_cambio_units():
    u = units()
    got = orc.calc_cambio(u["cc_new"], u["cc_old"])
    assert np.allclose(got, u["cc_out"], rtol=0, atol=1e-15)


# ------------------------------------------------------------------ a9-a11: energies, solvers
def _en_cfg(u, q):
    v = u["en_cfgv"][q]
    return cfg_ros(Q=[v[0], v[1]], R=[v[2], v[3], v[4]], cte_odom=v[5])


def test_energies_match_reference():
    u = units()
    for q in range(len(u["en_n"])):
        n = int(u["en_n"][q])
        cfg = _en_cfg(u, q)
        z, seen = u["en_z"][q][:n], u["en_seen"][q][:n]
        uu, odo = u["en_u"][q], u["en_odo"][q]
        fxn = orc.fun_xn(cfg, u["en_x"][q], u["en_x_ant"][q], u["en_x_pos"][q], uu[:, 0], uu[:, 1], odo, z[:, 0], z[:, 1],
                         seen[:, 0], seen[:, 1])
        fx = orc.fun_x(cfg, u["en_x"][q], u["en_x_ant"][q], uu[:, 0], odo[:, :2], z[:, 0], z[:, 1], seen[:, 0], seen[:, 1])
        assert abs(fxn - u["en_fxn"][q]) <= 1e-12 * max(1.0, abs(fxn)), q
        assert abs(fx - u["en_fx"][q]) <= 1e-12 * max(1.0, abs(fx)), q


def test_nelder_mead_matches_scipy_fmin():
    """The restated NM reproduces the reference's fmin results and evaluation counts."""
    u = units()
    for q in range(len(u["en_n"])):
        n = int(u["en_n"][q])
        cfg = _en_cfg(u, q)
        z, seen = u["en_z"][q][:n], u["en_seen"][q][:n]
        uu, odo = u["en_u"][q], u["en_odo"][q]
        p, nev = orc.solve_pose(cfg, "nm", u["en_x_ant"][q], u["en_x_pos"][q], uu[:, 0], uu[:, 1], odo, z[:, 0], z[:, 1],
                                seen[:, 0], seen[:, 1])
        assert np.max(np.abs(p - u["en_min_xn"][q])) <= 1e-9, (q, p, u["en_min_xn"][q])
        assert nev == int(u["en_nev_xn"][q]), q
        p, nev = orc.solve_pose(cfg, "nm", u["en_x_ant"][q], None, uu[:, 0], None, odo[:, :2], z[:, 0], z[:, 1],
                                seen[:, 0], seen[:, 1])
        assert np.max(np.abs(p - u["en_min_x"][q])) <= 1e-9, q
        assert nev == int(u["en_nev_x"][q]), q


def test_newton_is_the_exact_minimiser_of_the_pinned_energy():
    """The 'exact' solver's answer is a stationary point of the energy pinned above, and its
    energy is <= the reference NM's (which stops at xtol=1e-3)."""
    u = units()
    for q in range(len(u["en_n"])):
        n = int(u["en_n"][q])
        cfg = _en_cfg(u, q)
        z, seen = u["en_z"][q][:n], u["en_seen"][q][:n]
        uu, odo = u["en_u"][q], u["en_odo"][q]
        args = (u["en_x_ant"][q], u["en_x_pos"][q], uu[:, 0], uu[:, 1], odo, z[:, 0], z[:, 1], seen[:, 0], seen[:, 1])
        p, _ = orc.solve_pose(cfg, "newton", *args)
        f = lambda v: orc.fun_xn(cfg, v, *args)  # noqa: E731
        f0 = f(p)
        assert f0 <= f(u["en_min_xn"][q]) + 1e-12
        h = 1e-5
        for k in range(3):
            e = np.zeros(3)
            e[k] = h
            grad = (f(p + e) - f(p - e)) / (2 * h)
            assert abs(grad) <= 2e-7 * max(1.0, f0), (q, k, grad)
        # causal problem
        args2 = (u["en_x_ant"][q], uu[:, 0], odo[:, :2], z[:, 0], z[:, 1], seen[:, 0], seen[:, 1])
        p2, _ = orc.solve_pose(cfg, "newton", u["en_x_ant"][q], None, uu[:, 0], None, odo[:, :2], z[:, 0], z[:, 1],
                               seen[:, 0], seen[:, 1])
        f2 = lambda v: orc.fun_x(cfg, v, *args2)  # noqa: E731
        assert f2(p2) <= f2(u["en_min_x"][q]) + 1e-12
        for k in range(3):
            e = np.zeros(3)
            e[k] = h
            assert abs((f2(p2 + e) - f2(p2 - e)) / (2 * h)) <= 2e-7 * max(1.0, f2(p2)), (q, k)


# ------------------------------------------------------------------ a8: the sweep, reference semantics
def _run_reference_mode(gold, z, odo, u, map0, x0_arr, cfg, Lact0=None, nsweeps=None):
    med = orc.precondition(z, cfg.radio, cfg.rango_laser_max)
    ext = orc.extract_all(med, cfg)
    m = orc.Mapa(cfg)
    m.landmarks_actuales = map0.shape[1] if Lact0 is None else Lact0
    x = np.ascontiguousarray(x0_arr.copy())
    mapa = map0.copy()
    nsweeps = int(gold["nsweeps"]) if nsweeps is None else nsweeps
    for k in range(1, nsweeps + 1):
        p = "s%d_" % k
        assert np.array_equal(x, gold[p + "x_in"])
        r = orc.sweep(cfg, m, ext, odo, u, odo[:, 0], mapa, x, "sequential", "nm", "running")
        nt = np.diff(ext["off"])
        assert np.array_equal(nt[nt > 0], gold[p + "nt"]), "extraction partition differs"
        assert np.array_equal(r["c"], gold[p + "labels"]), "association labels differ in sweep %d" % k
        assert r["raw_L"] == int(gold[p + "raw_L"])
        assert np.array_equal(r["raw_counts"], gold[p + "raw_counts"])
        assert np.max(np.abs(r["raw_map"] - gold[p + "raw_map"])) <= 1e-11
        assert r["map"].shape == gold[p + "map_out"].shape
        assert np.max(np.abs(r["map"] - gold[p + "map_out"])) <= 1e-11
        assert np.array_equal(r["counts"], gold[p + "counts_out"])
        assert r["nev"] == int(gold[p + "nev"]), "Nelder-Mead took a different path"
        dx = np.abs(x - gold[p + "x_out"])
        assert dx[:2].max() <= 1e-9 and dx[2].max() <= 1e-9, dx.max(axis=1)
        cam = orc.calc_cambio(r["map"], mapa)
        assert np.allclose(cam, gold[p + "cambio"], rtol=0, atol=1e-11)
        mapa = r["map"]
        # carry the reference's own poses forward so a last-bit drift cannot accumulate into
        # a label flip in the next sweep (teacher forcing)
        x = np.ascontiguousarray(gold[p + "x_out"].copy())


def test_sweep_reference_mode_c1():
    """data_IJAC2018.mat: 2 sweeps of (sequential, NM, running) == ICM_ROS.iterations_process_offline."""
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    _run_reference_mode(g, z, odo, u, g["p0_map"], g["p0_x"], cfg_ros())


def test_sweep_reference_mode_c2():
    g = golden("c2_ref.npz")
    z, odo, u = c2_inputs()
    _run_reference_mode(g, z, odo, u, g["p0_map"], g["p0_x"], cfg_ros())


@pytest.mark.parametrize("name", ["synth_a.npz", "synth_b.npz"])
def test_sweep_reference_mode_synthetic(name):
    g = golden(name)
    cfg = cfg_ros(L=int(g["cfg_L"]), cota=float(g["cfg_cota"]))
    _run_reference_mode(g, g["observations"].astype(np.float64), g["odometry"], g["velocities"], g["map_init"],
                        g["x_init"], cfg)


def test_known_answers_from_survey():
    """SURVEY.md App. B values (independent record of the same reference run)."""
    import hashlib
    g = golden("c1_ref.npz")
    lab = g["s1_labels"].astype(np.int64)
    assert hashlib.sha256(lab.tobytes()).hexdigest().startswith("89fdbbfd6d2bdac1")
    assert lab.size == 17892 and lab.max() == 77 and int(g["s1_raw_L"]) == 78
    assert int(g["s1_nev"]) == 108766
    assert np.allclose(g["s1_cambio"], [0.0010211004, 0.0176403159, 0.0042196014], atol=1e-9)
    assert np.allclose(g["p0_x"][:, 916], [7.567557262574, 4.127036391777, 8.17156814594], atol=1e-11)
    assert np.allclose(g["s1_x_out"][:, 916], [7.562955973383, 4.125231421779, 8.17115687063], atol=1e-11)
    assert np.allclose(g["s2_x_out"][:, 1832], [-0.482178191892, 0.366393793252, 12.513832441742], atol=1e-11)
