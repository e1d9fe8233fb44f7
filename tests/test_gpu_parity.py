"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI,
against (a) the fixtures minted from the unmodified reference and (b) the pinned CPU oracle on
the same seeded inputs.

Bars (BASELINE.json north_star): association labels, label counts, landmark counts bit-exact;
poses / landmarks within 1e-6 m and 1e-8 rad in matching solver modes.
"""
import numpy as np
import pytest

from helpers import CONFIG_ROS, c1_inputs, c2_inputs, golden

pytestmark = pytest.mark.gpu

TOL_XY = 1e-6
TOL_TH = 1e-8


def _cfg(**kw):
    from icm_slam_b200.config import ConfigICM
    d = dict(CONFIG_ROS)
    d.update(kw)
    return ConfigICM.from_values(**d)


def _engine(cfg, z, odo, u):
    from icm_slam_b200.engine import Engine
    e = Engine(cfg)
    e.load(z, odo, u, precondition=True)
    e.extract()
    return e


def _oracle(cfgd, z, odo, u):
    from oracle import oracle as orc
    ocfg = orc.make_cfg(**cfgd)
    ext = orc.extract_all(orc.precondition(z, ocfg.radio, ocfg.rango_laser_max), ocfg)
    return orc, ocfg, ext


# ---------------------------------------------------------------------------------------------
def test_library_loads_and_reports_abi():
    from icm_slam_b200 import _lib
    assert _lib.lib().icmslam_abi_version() == 1


def test_extraction_matches_oracle_bit_exact_c1_c2():
    for z, odo, u in (c1_inputs(), c2_inputs()):
        e = _engine(_cfg(), z, odo, u)
        orc, ocfg, ext = _oracle(CONFIG_ROS, z, odo, u)
        g = e.get_extraction()
        assert np.array_equal(g["off"], ext["off"])
        assert np.array_equal(g["beam"], ext["beam"])
        assert np.array_equal(g["d"], ext["d"]) and np.array_equal(g["bx"], ext["bx"]) and np.array_equal(g["by"], ext["by"])
        e.close()
    assert ext["n"] > 17892  # c2 is the raw, denser log


def test_extraction_units_vs_reference_fixture():
    u = golden("units.npz")
    T = u["fz_scans"].shape[1]
    e = _engine(_cfg(), u["fz_scans"], np.zeros((3, T)), np.zeros((2, T)))   # fixture scans are already pre-conditioned
    e2 = None
    from icm_slam_b200.engine import Engine
    e2 = Engine(_cfg())
    e2.load(u["fz_scans"], np.zeros((3, T)), np.zeros((2, T)), precondition=False)
    e2.extract()
    g = e2.get_extraction()
    assert np.array_equal(g["off"], u["fz_off"])
    rows = u["fz_rows"]
    assert np.array_equal(g["d"], rows[:, 0]) and np.array_equal(g["bx"], rows[:, 2]) and np.array_equal(g["by"], rows[:, 3])
    assert np.array_equal(g["beam"] * np.pi / 180.0, rows[:, 1])
    e.close()
    e2.close()


def test_filtrar_obs_reproduces_dataset_pair():
    raw, odo, u = c2_inputs()
    want, _, _ = c1_inputs()
    from icm_slam_b200.engine import Engine
    e = Engine(_cfg())
    got = e.filtrar_obs(raw, 10.0, 15)
    assert np.array_equal(got, want)
    e.close()


def _teacher_forced_reference(gold, z, odo, u, cfgd, map0_key, x0_key, view):
    """Per sweep: feed the reference's own input poses/map; association labels, label counts, raw
    landmark count and filtered landmark count must equal the reference bit for bit (they do not
    depend on the pose update)."""
    cfg = _cfg(**cfgd)
    e = _engine(cfg, z, odo, u)
    mapa = gold[map0_key].copy()
    e.landmarks_actuales = mapa.shape[1]
    for k in range(1, int(gold["nsweeps"]) + 1):
        p = "s%d_" % k
        x = np.ascontiguousarray(gold[p + "x_in"].copy())
        st, Lout, mout = e.sweep(mapa, x, odo[:, 0], schedule="redblack", solver="newton", view=view, fused=False)
        assert st == 0
        assert np.array_equal(e.associations(), gold[p + "labels"]), "labels differ in sweep %d" % k
        raw, cnt, rl = e.raw_map()
        assert rl == int(gold[p + "raw_L"])
        assert np.array_equal(cnt, gold[p + "raw_counts"])
        assert np.max(np.abs(raw - gold[p + "raw_map"])) <= 1e-9
        assert Lout == gold[p + "map_out"].shape[1]
        assert np.max(np.abs(mout - gold[p + "map_out"])) <= 1e-9
        assert np.array_equal(e.counts(Lout), gold[p + "counts_out"])
        cam = e.calc_cambio(mout, mapa)
        assert np.allclose(cam, gold[p + "cambio"], rtol=0, atol=1e-9)
        mapa = gold[p + "map_out"].copy()
        assert e.landmarks_actuales == mapa.shape[1]
    e.close()


@pytest.mark.parametrize("view", ["prev", "full", "running"])
def test_associations_and_map_vs_reference_c1(view):
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    _teacher_forced_reference(g, z, odo, u, {}, "p0_map", "p0_x", view)


def test_associations_and_map_vs_reference_c2():
    g = golden("c2_ref.npz")
    z, odo, u = c2_inputs()
    _teacher_forced_reference(g, z, odo, u, {}, "p0_map", "p0_x", "prev")


@pytest.mark.parametrize("name", ["synth_a.npz", "synth_b.npz"])
def test_associations_and_map_vs_reference_synthetic(name):
    g = golden(name)
    _teacher_forced_reference(g, g["observations"].astype(np.float64), g["odometry"], g["velocities"],
                              dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"])), "map_init", "x_init", "running")


def _reference_mode_gpu(gold, z, odo, u, cfgd, map0):
    """(sequential, NM, running) on the GPU == the unmodified reference: poses to 1e-6 m / 1e-8 rad."""
    cfg = _cfg(**cfgd)
    e = _engine(cfg, z, odo, u)
    mapa = map0.copy()
    e.landmarks_actuales = mapa.shape[1]
    for k in range(1, int(gold["nsweeps"]) + 1):
        p = "s%d_" % k
        x = np.ascontiguousarray(gold[p + "x_in"].copy())
        st, Lout, mout = e.sweep(mapa, x, odo[:, 0], schedule="sequential", solver="nm", view="running", stats=True)
        assert np.array_equal(e.associations(), gold[p + "labels"])
        d = np.abs(x - gold[p + "x_out"])
        assert d[:2].max() <= TOL_XY and d[2].max() <= TOL_TH, d.max(axis=1)
        assert np.max(np.abs(mout - gold[p + "map_out"])) <= TOL_XY
        assert e.sweep_stats()["newton_iters"] == int(gold[p + "nev"])   # same number of energy evaluations
        mapa = gold[p + "map_out"].copy()
    e.close()


def test_reference_mode_poses_c1():
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    _reference_mode_gpu(g, z, odo, u, {}, g["p0_map"])


@pytest.mark.parametrize("name", ["synth_a.npz", "synth_b.npz"])
def test_reference_mode_poses_synthetic(name):
    g = golden(name)
    _reference_mode_gpu(g, g["observations"].astype(np.float64), g["odometry"], g["velocities"],
                        dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"])), g["map_init"])


MODES = [("redblack", "newton", "prev"), ("redblack", "newton", "full"), ("redblack", "newton", "running"),
         ("sequential", "newton", "running"), ("redblack", "nm", "running"), ("sequential", "newton", "prev")]


@pytest.mark.parametrize("schedule,solver,view", MODES)
def test_modes_vs_oracle_c1(schedule, solver, view):
    """3 chained sweeps on data_IJAC2018.mat in each restated mode: GPU == oracle in the same mode."""
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    orc, ocfg, ext = _oracle(CONFIG_ROS, z, odo, u)
    e = _engine(_cfg(), z, odo, u)
    mo = orc.Mapa(ocfg)
    map_o = g["p0_map"].copy()
    map_g = g["p0_map"].copy()
    mo.landmarks_actuales = map_o.shape[1]
    e.landmarks_actuales = map_g.shape[1]
    xo = np.ascontiguousarray(g["p0_x"].copy())
    xg = xo.copy()
    for k in range(3):
        r = orc.sweep(ocfg, mo, ext, odo, u, odo[:, 0], map_o, xo, schedule, solver, view)
        st, Lout, mout = e.sweep(map_g, xg, odo[:, 0], schedule=schedule, solver=solver, view=view, fused=False,
                                 newton_tol=1e-12, newton_maxit=50)
        assert np.array_equal(e.associations(), r["c"])
        assert Lout == r["map"].shape[1]
        d = np.abs(xg - xo)
        assert d[:2].max() <= TOL_XY and d[2].max() <= TOL_TH, (k, d.max(axis=1))
        assert np.max(np.abs(mout - r["map"])) <= TOL_XY
        map_o = r["map"]
        map_g = np.ascontiguousarray(mout.copy())
    e.close()


def _synthetic_case(L_true, T, seed, cota=20.0):
    from icm_slam_b200.synthetic import make_synthetic
    d = make_synthetic(L_true, T=T, seed=seed)
    cfgd = dict(CONFIG_ROS, L=2 * L_true + 64, cota=cota)
    return d, cfgd


@pytest.mark.parametrize("view", ["prev", "full", "running"])
def test_synthetic_medium_vs_oracle(view):
    """T=6000 / 625 landmarks: grid association vs the oracle's brute force, 2 sweeps."""
    d, cfgd = _synthetic_case(625, 6000, 20181 + 7)
    z, odo, u = d["observations"], d["odometry"], d["velocities"]
    orc, ocfg, ext = _oracle(cfgd, z, odo, u)
    e = _engine(_cfg(**cfgd), z, odo, u)
    assert e.n == ext["n"]
    mo = orc.Mapa(ocfg)
    map_o = d["map_init"].copy()
    map_g = d["map_init"].copy()
    mo.landmarks_actuales = map_o.shape[1]
    e.landmarks_actuales = map_g.shape[1]
    xo = np.ascontiguousarray(d["x_init"].copy())
    xg = xo.copy()
    for k in range(2):
        r = orc.sweep(ocfg, mo, ext, odo, u, odo[:, 0], map_o, xo, "redblack", "newton", view)
        st, Lout, mout = e.sweep(map_g, xg, odo[:, 0], schedule="redblack", solver="newton", view=view, fused=False,
                                 newton_tol=1e-12, newton_maxit=50)
        assert np.array_equal(e.associations(), r["c"])
        assert Lout == r["map"].shape[1]
        dd = np.abs(xg - xo)
        assert dd[:2].max() <= TOL_XY and dd[2].max() <= TOL_TH, (k, dd.max(axis=1))
        assert np.max(np.abs(mout - r["map"])) <= TOL_XY
        map_o = r["map"]
        map_g = np.ascontiguousarray(mout.copy())
    e.close()


def test_filter_map_units_vs_reference_fixture():
    u = golden("units.npz")
    from icm_slam_b200.engine import Engine
    e = Engine(_cfg(L=40, cota=10.0))
    for q in range(int(u["fl_n"])):
        if not int(u[f"fl{q}_ok"]):
            continue
        out, cnt, Lout = e.filter_map(u[f"fl{q}_in"], u[f"fl{q}_cnt"])
        assert Lout == int(u[f"fl{q}_Lact"]), q
        assert np.array_equal(cnt, u[f"fl{q}_cant"]), q
        assert np.max(np.abs(out - u[f"fl{q}_out"])) <= 1e-12, q
    e.close()


@pytest.mark.parametrize("fused", [False, True])
def test_error_codes_mirror_reference_failures(fused):
    g = golden("synth_a.npz")
    z = g["observations"].astype(np.float64).copy()
    odo, u = g["odometry"], g["velocities"]
    cfgd = dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"]))
    # empty last scan -> IndexError (sensors.py:148)
    z2 = z.copy()
    z2[:, -1] = 10.0
    e = _engine(_cfg(**cfgd), z2, odo, u)
    e.landmarks_actuales = g["map_init"].shape[1]
    with pytest.raises(IndexError):
        e.sweep(g["map_init"].copy(), np.ascontiguousarray(g["x_init"].copy()), odo[:, 0], fused=fused)
    e.close()
    # empty first scan -> inputs returned unchanged (sensors.py:137-139)
    z3 = z.copy()
    z3[:, 0] = 10.0
    e = _engine(_cfg(**cfgd), z3, odo, u)
    e.landmarks_actuales = g["map_init"].shape[1]
    x = np.ascontiguousarray(g["x_init"].copy())
    st, Lout, mout = e.sweep(g["map_init"].copy(), x, odo[:, 0], fused=fused)
    assert st == 1 and np.array_equal(x, g["x_init"]) and np.array_equal(mout, g["map_init"])
    e.close()
    # label capacity -> IndexError (ICM_SLAM.py:191)
    e = _engine(_cfg(L=g["map_init"].shape[1], cota=1.0), z, odo, u)
    e.landmarks_actuales = g["map_init"].shape[1]
    with pytest.raises(IndexError):
        e.sweep(g["map_init"].copy()[:, :5] + 100.0, np.ascontiguousarray(g["x_init"].copy()), odo[:, 0], fused=fused)
    e.close()
    # nothing reaches cota -> ValueError (ICM_SLAM.py:255)
    e = _engine(_cfg(L=int(g["cfg_L"]), cota=1e9), z, odo, u)
    e.landmarks_actuales = g["map_init"].shape[1]
    with pytest.raises(ValueError):
        e.sweep(g["map_init"].copy(), np.ascontiguousarray(g["x_init"].copy()), odo[:, 0], fused=fused)
    e.close()


# ---- the fused single-kernel path (redblack, newton, prev) -----------------------------------------
def test_fused_teacher_forced_vs_reference_c1_c2():
    """The fused kernel's association / label / count outputs against the reference's own, sweep by sweep."""
    for gname, inputs in (("c1_ref.npz", c1_inputs), ("c2_ref.npz", c2_inputs)):
        gold = golden(gname)
        z, odo, u = inputs()
        e = _engine(_cfg(), z, odo, u)
        mapa = gold["p0_map"].copy()
        e.landmarks_actuales = mapa.shape[1]
        for k in range(1, int(gold["nsweeps"]) + 1):
            p = "s%d_" % k
            x = np.ascontiguousarray(gold[p + "x_in"].copy())
            st, Lout, mout = e.sweep(mapa, x, odo[:, 0], fused=True)
            assert st == 0
            assert np.array_equal(e.associations(), gold[p + "labels"]), "labels differ in sweep %d" % k
            raw, cnt, rl = e.raw_map()
            assert rl == int(gold[p + "raw_L"])
            assert np.array_equal(cnt, gold[p + "raw_counts"])
            assert np.max(np.abs(raw - gold[p + "raw_map"])) <= 1e-9
            assert Lout == gold[p + "map_out"].shape[1]
            assert np.max(np.abs(mout - gold[p + "map_out"])) <= 1e-9
            assert np.array_equal(e.counts(Lout), gold[p + "counts_out"])
            mapa = gold[p + "map_out"].copy()
        e.close()


def _fused_vs_oracle(z, odo, u, cfgd, map0, x_init, nsweeps, chained):
    orc, ocfg, ext = _oracle(cfgd, z, odo, u)
    e = _engine(_cfg(**cfgd), z, odo, u)
    mo = orc.Mapa(ocfg)
    map_o = map0.copy()
    mo.landmarks_actuales = map_o.shape[1]
    xo = np.ascontiguousarray(x_init.copy())
    xg = xo.copy()
    if chained:
        e.set_map(map0)
        e.set_poses(xg)
    else:
        map_g = map0.copy()
        e.landmarks_actuales = map_g.shape[1]
    for k in range(nsweeps):
        r = orc.sweep(ocfg, mo, ext, odo, u, odo[:, 0], map_o, xo, "redblack", "newton", "prev")
        if chained:
            e.iterate(None, odo[:, 0], 1, fused=True)
            xg = e.get_poses()
            mout = e.get_map()
            Lout = mout.shape[1]
        else:
            st, Lout, mout = e.sweep(map_g, xg, odo[:, 0], fused=True)
            map_g = np.ascontiguousarray(mout.copy())
        assert np.array_equal(e.associations(), r["c"]), k
        assert Lout == r["map"].shape[1]
        d = np.abs(xg - xo)
        assert d[:2].max() <= TOL_XY and d[2].max() <= TOL_TH, (k, d.max(axis=1))
        assert np.max(np.abs(mout - r["map"])) <= TOL_XY
        map_o = r["map"]
    e.close()


@pytest.mark.parametrize("chained", [False, True])
def test_fused_vs_oracle_c1(chained):
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    _fused_vs_oracle(z, odo, u, dict(CONFIG_ROS), g["p0_map"], g["p0_x"], 3, chained)


def test_fused_vs_oracle_c2_dense_scans():
    """datos_palomar1 raw scans: up to 119 kept beams per scan (tiles beyond the shared-memory budget)."""
    g = golden("c2_ref.npz")
    z, odo, u = c2_inputs()
    _fused_vs_oracle(z, odo, u, dict(CONFIG_ROS), g["p0_map"], g["p0_x"], 2, False)


@pytest.mark.parametrize("T,chained", [(6000, True), (2047, False), (127, False), (3, False)])
def test_fused_vs_oracle_synthetic(T, chained):
    d, cfgd = _synthetic_case(625, T, 20181 + 7, cota=20.0 if T > 1000 else 1.0)
    _fused_vs_oracle(d["observations"], d["odometry"], d["velocities"], cfgd, d["map_init"], d["x_init"], 2, chained)


def test_fused_small_smem_budget_fallback(monkeypatch):
    """Tiles whose observations exceed the shared-memory budget read global memory directly: same results."""
    monkeypatch.setenv("ICMSLAM_OBS_CAP", "64")
    d, cfgd = _synthetic_case(625, 1500, 20181 + 9)
    _fused_vs_oracle(d["observations"], d["odometry"], d["velocities"], cfgd, d["map_init"], d["x_init"], 2, False)


def test_fused_graph_replay_many_sweeps_c1():
    """8 chained sweeps on data_IJAC2018.mat (landmarks merge in the filter: the one-block merge path and the
    grid rebuild run, and from the second sweep on the sweep replays as a CUDA graph) against the oracle."""
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    orc, ocfg, ext = _oracle(CONFIG_ROS, z, odo, u)
    e = _engine(_cfg(), z, odo, u)
    mo = orc.Mapa(ocfg)
    map_o = g["p0_map"].copy()
    mo.landmarks_actuales = map_o.shape[1]
    xo = np.ascontiguousarray(g["p0_x"].copy())
    e.set_map(g["p0_map"])
    e.set_poses(xo)
    for k in range(8):
        r = orc.sweep(ocfg, mo, ext, odo, u, odo[:, 0], map_o, xo, "redblack", "newton", "prev")
        map_o = r["map"]
    e.iterate(None, odo[:, 0], 3)
    e.iterate(None, odo[:, 0], 5)
    xg = e.get_poses()
    mout = e.get_map()
    assert np.array_equal(e.associations(), r["c"])
    assert mout.shape == map_o.shape
    d = np.abs(xg - xo)
    assert d[:2].max() <= TOL_XY and d[2].max() <= TOL_TH, d.max(axis=1)
    assert np.max(np.abs(mout - map_o)) <= TOL_XY
    e.close()


def test_fused_launch_count_and_kernel_timer():
    d, cfgd = _synthetic_case(625, 3000, 20181 + 11)
    e = _engine(_cfg(**cfgd), d["observations"], d["odometry"], d["velocities"])
    e.set_map(d["map_init"])
    e.set_poses(d["x_init"])
    n0 = e.launch_count()
    e.iterate(None, d["odometry"][:, 0], 2, timing=True)
    a, b = e.kernel_ms()
    assert a > 0.0 and b > 0.0      # k_runs + k_assoc_tiles | k_solve_tile
    assert e.launch_count() - n0 >= 2 * 14
    e.close()


# ---- time-segment partition: several segments emulated on ONE GPU must reproduce the single-engine sweep -----
def _segmented_vs_single(d, cfgd, world, nsweeps):
    import ctypes as C
    import torch
    from icm_slam_b200 import _lib
    from icm_slam_b200.multigpu import SegmentedSolver, SEG_REC
    z, odo, u = d["observations"], d["odometry"], d["velocities"]
    cfg = _cfg(**cfgd)
    single = _engine(cfg, z, odo, u)
    single.set_map(d["map_init"])
    single.set_poses(d["x_init"])
    sols = [SegmentedSolver(cfg, r, world, device=0) for r in range(world)]
    for s in sols:
        s.load(z, odo, u, precondition=True)
        s.set_map(d["map_init"])
        s.set_poses(d["x_init"])
        s._bind()
    assert sum(s.t_hi - s.t_lo for s in sols) == z.shape[1]
    opts = _lib.SweepOpts(_lib.SCHED["redblack"], _lib.SOLVER["newton"], _lib.VIEW["prev"], 0, 0.0, 1, 0)
    for k in range(nsweeps):
        single.iterate(None, odo[:, 0], 1)
        for s in sols:
            _lib.check(s.engine.lib.icmslam_seg_begin(s.engine._h, C.c_void_p(s.x0.ctypes.data), C.byref(opts)), s.engine._h)
        for s in sols:
            s.engine.synchronize()
        allrec = torch.stack([s._views["rec"].clone() for s in sols]).contiguous()          # the all-gather
        assert allrec.shape == (world, SEG_REC)
        for s in sols:
            _lib.check(s.engine.lib.icmslam_seg_exchange(s.engine._h, C.c_void_p(allrec.data_ptr()), s.rank, world), s.engine._h)
        for s in sols:
            s.engine.synchronize()
        for key in ("all",):                                               # the sum-reduction: ONE block of int64 words, as in production
            tot = torch.stack([s._views[key] for s in sols]).sum(0)
            for s in sols:
                s._views[key].copy_(tot)
        torch.cuda.synchronize()
        for s in sols:
            _lib.check(s.engine.lib.icmslam_seg_finish(s.engine._h), s.engine._h)
        xs = np.concatenate([s.owned_poses() for s in sols], axis=1)
        x1 = single.get_poses()
        assert np.array_equal(xs, x1), (k, np.abs(xs - x1).max())
        m1 = single.get_map()
        for s in sols:
            assert np.array_equal(s.get_map(), m1), k
        off1 = single.get_extraction()["off"]
        c1 = single.associations()
        for s in sols:
            g_lo, g_hi = s.segments[s.rank]
            offs = s.engine.get_extraction()["off"]
            cs = s.engine.associations()[offs[s.t_lo]:offs[s.t_hi]]
            assert np.array_equal(cs, c1[off1[g_lo]:off1[g_hi]]), k
    single.close()
    for s in sols:
        s.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_segments_reproduce_single_engine_bit_for_bit(world):
    d, cfgd = _synthetic_case(625, 4000, 20181 + 13)
    _segmented_vs_single(d, cfgd, world, 3)


def test_segments_with_new_labels_and_merges():
    """Landmarks displaced in the initial map: scans create new labels (numbered across segments) and the filter merges."""
    d, cfgd = _synthetic_case(625, 3000, 20181 + 14)
    d = dict(d)
    m = d["map_init"].copy()
    m[:, ::9] += 3.0
    m = np.concatenate([m, m[:, :40] + 0.3], axis=1)      # near-duplicates: merged by Mapa.filtrar
    d["map_init"] = m
    cfgd = dict(cfgd, L=4096)
    _segmented_vs_single(d, cfgd, 4, 3)


def test_segmented_solver_single_rank_graph_replay():
    """SegmentedSolver with world = 1 (no collectives): eager sweeps, then CUDA-graph replays of sweep pairs,
    equal bit for bit to the plain engine."""
    import torch
    from icm_slam_b200.multigpu import SegmentedSolver
    d, cfgd = _synthetic_case(625, 4000, 20181 + 15)
    z, odo, u = d["observations"], d["odometry"], d["velocities"]
    cfg = _cfg(**cfgd)
    single = _engine(cfg, z, odo, u)
    single.set_map(d["map_init"]); single.set_poses(d["x_init"])
    single.iterate(None, odo[:, 0], 9)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        sol = SegmentedSolver(cfg, 0, 1, device=0)
        sol.engine.set_stream(stream.cuda_stream)
        sol.load(z, odo, u, precondition=True)
        sol.set_map(d["map_init"]); sol.set_poses(d["x_init"])
        sol.sweep(7)          # 2 eager + capture + 2 replays + 1 eager
        sol.sweep(2)          # 1 replay
        stream.synchronize()
        assert sol._graph is not None
        assert np.array_equal(sol.owned_poses(), single.get_poses())
        assert np.array_equal(sol.get_map(), single.get_map())
    sol.close(); single.close()


def test_fused_exact_ties_and_duplicate_landmarks():
    """Exact duplicates in the previous map make every observation near them an exact distance tie (np.argmin takes
    the first index), near-duplicates exercise Mapa.filtrar's merge path, displaced landmarks create new labels:
    four chained sweeps (grid search, then hints) against the oracle."""
    d, cfgd = _synthetic_case(625, 2500, 20181 + 17)
    m = d["map_init"].copy()
    dup = m[:, 10:60].copy()                      # exact copies, appended AFTER the originals and also
    m = np.concatenate([dup[:, :25], m, dup[:, 25:], m[:, 100:130] + 0.2], axis=1)   # ... BEFORE them
    m[:, 200::11] += 2.5
    cfgd = dict(cfgd, L=4096)
    _fused_vs_oracle(d["observations"], d["odometry"], d["velocities"], cfgd, m, d["x_init"], 4, True)


def test_fused_sparse_and_empty_scans():
    """Blocks of empty scans (no kept beam): the averaging rule of sensors.py:147-151 inside the red-black order."""
    d, cfgd = _synthetic_case(625, 3000, 20181 + 18)
    z = d["observations"].copy()
    z[:, 200:203] = 10.0          # three consecutive empty scans
    z[:, 1001] = 10.0
    z[:, 1500:1502] = 10.0
    z[:, 2:5] = 10.0              # right after the pinned first pose
    _fused_vs_oracle(z, d["odometry"], d["velocities"], cfgd, d["map_init"], d["x_init"], 3, True)


# ---- pass 0 (causal initialisation) against the reference fixtures ------------------------------------------------
@pytest.mark.parametrize("gname,inputs", [("c1_ref.npz", c1_inputs), ("c2_ref.npz", c2_inputs)])
def test_pass0_vs_reference(gname, inputs):
    """inicializar_online replayed natively: labels of every scan bit-exact (incl. the fcluster step at t = 0), raw
    landmark count / counts exact, poses within 1e-6 m / 1e-8 rad of the reference's Nelder-Mead, map within 1e-6 m."""
    g = golden(gname)
    z, odo, u = inputs()
    e = _engine(_cfg(), z, odo, u)
    x, mapa = e.pass0(odo[:, 0])
    assert np.array_equal(e.associations(), g["p0_labels"])
    raw, cnt, rl = e.raw_map()
    assert rl == int(g["p0_raw_L"])
    assert np.array_equal(cnt, g["p0_raw_counts"])
    assert np.max(np.abs(raw - g["p0_raw_map"])) <= 1e-9
    d = np.abs(x - g["p0_x"])
    assert d[:2].max() <= TOL_XY and d[2].max() <= TOL_TH, d.max(axis=1)
    assert mapa.shape == g["p0_map"].shape
    assert np.max(np.abs(mapa - g["p0_map"])) <= TOL_XY
    assert e.landmarks_actuales == mapa.shape[1]
    e.close()


def test_offline_batch_path_through_the_reference_surface():
    """load_data -> inicializar -> N x itererar, the legacy driver of external_options.py:58-89, on data_IJAC2018."""
    from icm_slam_b200.icm import ICM_SLAM, Mapa, precondicionar, calc_cambio
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    cfg = _cfg(N=2)
    icm = ICM_SLAM(cfg, x0=odo[:, 0])
    icm.load_data(Mapa(cfg), precondicionar(z, cfg), u, odo)
    mapa_inicial, x = icm.inicializar(np.zeros((3, z.shape[1])))
    assert np.max(np.abs(mapa_inicial - g["p0_map"])) <= TOL_XY and np.abs(x - g["p0_x"])[:2].max() <= TOL_XY
    mapa_viejo = mapa_inicial.copy()
    for it in range(cfg.N):
        mapa_refinado, x = icm.itererar(mapa_viejo, x)
        cam = calc_cambio(mapa_refinado, mapa_viejo, cfg)
        assert len(cam) == 3 and cam[0] <= cam[2] <= cam[1]
        mapa_viejo = mapa_refinado.copy()
    assert mapa_refinado.shape == (2, 11)         # the notebook's 11 trees


# ---- run records (runs.cuh): steady-state sweeps on certified run records must equal full re-association bit for bit ----
def _chain(z, odo, u, cfgd, map0, x_init, nsweeps, env, stats=False):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        e = _engine(_cfg(**cfgd), z, odo, u)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    e.set_map(map0)
    e.set_poses(np.ascontiguousarray(x_init.copy()))
    out = []
    for k in range(nsweeps):
        e.iterate(None, odo[:, 0], 1, fused=True, stats=stats)
        st = e.sweep_stats()
        out.append((e.get_poses().copy(), e.get_map().copy(), e.associations().copy(), st["dirty_tiles"], st["n_tiles"]))
    e.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("gname", ["synth_a.npz", "synth_b.npz"])
def test_run_record_sweeps_equal_full_association_synthetic(gname):
    g = golden(gname)
    z, odo, u = g["observations"].astype(np.float64), g["odometry"], g["velocities"]
    cfgd = dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"]))
    a = _chain(z, odo, u, cfgd, g["map_init"], g["x_init"], 8, {"ICMSLAM_RUNS": "0"}, stats=True)
    b = _chain(z, odo, u, cfgd, g["map_init"], g["x_init"], 8, {"ICMSLAM_RUNS": "1"}, stats=True)
    for k, (ra, rb) in enumerate(zip(a, b)):
        assert np.array_equal(ra[2], rb[2]), k
        assert np.array_equal(ra[0], rb[0]), k          # poses: bit for bit
        assert np.array_equal(ra[1], rb[1]), k          # map: bit for bit
        assert ra[3] == ra[4]                           # RUNS=0: every tile through the association kernel
    assert b[0][3] == b[0][4], "the first sweep has no records to run on"
    assert any(r[3] < r[4] for r in b[1:]), "no tile ever ran on its run records: %r" % ([r[3:] for r in b],)


@pytest.mark.gpu
def test_run_record_sweeps_equal_full_association_c1():
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    a = _chain(z, odo, u, dict(CONFIG_ROS), g["p0_map"], g["p0_x"], 6, {"ICMSLAM_RUNS": "0"})
    b = _chain(z, odo, u, dict(CONFIG_ROS), g["p0_map"], g["p0_x"], 6, {"ICMSLAM_RUNS": "1"})
    for k, (ra, rb) in enumerate(zip(a, b)):
        assert np.array_equal(ra[2], rb[2]), k
        assert np.array_equal(ra[0], rb[0]), k
        assert np.array_equal(ra[1], rb[1]), k


# ---- C5: independent trajectories, one handle each (batch.py) == the same trajectories run one at a time ---------
def test_trajectory_batch_equals_individual_runs():
    from icm_slam_b200.batch import TrajectoryBatch
    from icm_slam_b200.synthetic import make_synthetic
    cfgd = dict(L=2 * 256 + 64, cota=20.0)
    cfg = _cfg(**cfgd)
    data = [make_synthetic(256, T=1024, seed=20181 + 100 + i) for i in range(5)]
    world = 2
    got = {}
    for rank in range(world):                  # both ranks emulated on this GPU
        b = TrajectoryBatch(cfg, rank, world, device=0)
        for i in b.owned(len(data)):
            d = data[i]
            b.add(i, d["observations"], d["odometry"], d["velocities"], d["map_init"], d["x_init"])
        b.iterate(4)
        got.update(b.results())
        b.close()
    assert sorted(got) == list(range(len(data)))
    for i, d in enumerate(data):
        e = _engine(cfg, d["observations"], d["odometry"], d["velocities"])
        e.set_map(d["map_init"])
        e.set_poses(d["x_init"])
        e.iterate(None, d["odometry"][:, 0], 4)
        assert np.array_equal(e.get_poses(), got[i][0]) and np.array_equal(e.get_map(), got[i][1])
        e.close()


# ---- row f2: .mat -> filtrar_obs -> pass 0 -> N sweeps -> result files, in one call (offline.py) -------------------
def test_run_offline_one_call_matches_stepwise(tmp_path):
    import scipy.io as sio
    from icm_slam_b200.offline import run_offline, save_result, load_result
    from icm_slam_b200.icm import ICM_SLAM, Mapa, calc_cambio, precondicionar
    raw, odo, u = c2_inputs()                                   # the raw log (datos_palomar1.mat layout)
    path = str(tmp_path / "datos.mat")
    sio.savemat(path, {"datos": {"observaciones": raw, "odometria": odo, "control": u}})
    cfg = _cfg(N=3)
    res = run_offline(path, cfg)
    assert res["x"].shape == (3, raw.shape[1]) and res["cambios"].shape == (3, 3)
    # stepwise: the filtered log is data_IJAC2018.mat (the dataset pair), then the facade's own calls
    z1, _, _ = c1_inputs()
    icm = ICM_SLAM(cfg, x0=odo[:, 0])
    icm.load_data(Mapa(cfg), precondicionar(z1, cfg), u, odo)
    mapa, x = icm.inicializar()
    assert np.array_equal(np.array(mapa), res["mapa_inicial"]) and np.array_equal(np.array(x), res["x_inicial"])
    mapa = np.array(mapa)
    x = np.ascontiguousarray(np.array(x))
    for k in range(3):
        nuevo, x = icm.iterations_process_offline(mapa, x)
        assert np.allclose(calc_cambio(nuevo, mapa, cfg), res["cambios"][k], rtol=0, atol=1e-12)
        mapa = nuevo
    assert np.max(np.abs(x - res["x"])[:2]) <= TOL_XY and np.max(np.abs(x - res["x"])[2]) <= TOL_TH
    assert mapa.shape == res["mapa"].shape and np.max(np.abs(mapa - res["mapa"])) <= TOL_XY
    for ext in ("mat", "npz"):
        back = load_result(save_result(str(tmp_path / ("out." + ext)), res))
        assert np.array_equal(back["x"], res["x"]) and np.array_equal(back["mapa"], res["mapa"])


# ---- the sweep's variants: scheduling and staging switches change nothing (bit for bit) ---------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"ICMSLAM_OVERLAP": "0"}, {"ICMSLAM_GRAPH": "0"}, {"ICMSLAM_RUNS": "0"}, {"ICMSLAM_OBS_CAP": "96"},
                                 {"ICMSLAM_BLOCKS_PER_SM": "1"}, {"ICMSLAM_STEADY": "0"}])
def test_sweep_variants_agree(env):
    g = golden("synth_b.npz")
    z, odo, u = g["observations"].astype(np.float64), g["odometry"], g["velocities"]
    cfgd = dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"]))
    base = {"ICMSLAM_OVERLAP": "1", "ICMSLAM_GRAPH": "1", "ICMSLAM_RUNS": "1", "ICMSLAM_STEADY": "1"}
    a = _chain(z, odo, u, cfgd, g["map_init"], g["x_init"], 5, dict(base))
    b = _chain(z, odo, u, cfgd, g["map_init"], g["x_init"], 5, dict(base, **env))
    for k, (ra, rb) in enumerate(zip(a, b)):
        assert np.array_equal(ra[2], rb[2]), (env, k)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), (env, k)


# ---- the per-scan step on the reference's surface: Mapa.actualizar (ICM_SLAM.py:128-201) through icmslam_associate ----------
def test_mapa_actualizar_vs_reference_units():
    from icm_slam_b200.icm import Mapa
    u = golden("units.npz")
    cfg = _cfg(L=60)
    for q in range(int(u["ac_nseq"])):
        ref = u[f"ac{q}_ref"]
        m = Mapa(cfg)
        m.landmarks_actuales = ref.shape[1]
        m.clear_obs()
        y = np.zeros((2, 60))
        for r in range(int(u[f"ac{q}_ncall"])):
            y2, c = m.actualizar(y, ref, u[f"ac{q}_{r}_obs"])
            assert y2 is y and c.dtype == np.int64
            assert np.array_equal(c, u[f"ac{q}_{r}_c"]), (q, r)
            assert m.landmarks_actuales == int(u[f"ac{q}_{r}_Lact"])
            assert np.array_equal(m.cant_obs_i, u[f"ac{q}_{r}_cant"])
            assert np.max(np.abs(y - u[f"ac{q}_{r}_y"])) <= 1e-13


def test_mapa_actualizar_first_scan_clusters_like_scipy():
    """landmarks_actuales == 0: Branch A (fcluster of the first scan, ICM_SLAM.py:160-165), then Branch B against the map it made,
    then the label cap (IndexError, :191)."""
    from icm_slam_b200.icm import Mapa
    g = golden("c1_ref.npz")
    z, odo, u_ = c1_inputs()
    from oracle import oracle as orc
    ocfg = orc.make_cfg(**CONFIG_ROS)
    zz, _ = orc.filtrar_z(orc.precondition(z[:, 0:1], ocfg.radio, ocfg.rango_laser_max)[:, 0], ocfg)
    wx, wy = orc.tras_rot(odo[:, 0], zz[:, 2], zz[:, 3])
    obs = np.stack([wx, wy], axis=1)
    m = Mapa(_cfg())
    y = np.zeros((2, 1000))
    y, c = m.actualizar(y, y, obs)
    k = int(c.max()) + 1
    assert m.landmarks_actuales == k and np.array_equal(c, g["p0_labels"][: len(c)])
    for i in range(k):
        assert np.max(np.abs(y[:, i] - obs[c == i].mean(axis=0))) <= 1e-13 and m.cant_obs_i[i] == (c == i).sum()
    y, c2 = m.actualizar(y, y, obs + 0.01)
    assert np.array_equal(c2, c)
    small = Mapa(_cfg(L=k))
    small.landmarks_actuales = k
    with pytest.raises(IndexError):
        small.actualizar(np.zeros((2, k)), y[:, :k], obs + 50.0)


def test_iterate_until_monitors_calc_cambio_and_stops():
    """The convergence monitor (sensors.py:302-315): every pass's calc_cambio triple from the sweep's own filter equals the
    full search of the reference's calc_cambio on the same maps, and the loop stops at the requested change."""
    from oracle import oracle as orc
    d, cfgd = _synthetic_case(625, 4000, 20181 + 23)
    z, odo, u = d["observations"], d["odometry"], d["velocities"]
    e = _engine(_cfg(**cfgd), z, odo, u)
    e.set_map(d["map_init"]); e.set_poses(d["x_init"])
    maps = [d["map_init"]]
    n, cam = e.iterate_until(odo[:, 0], 6, 0.0)
    assert n == 6 and cam.shape == (6, 3)
    e2 = _engine(_cfg(**cfgd), z, odo, u)
    e2.set_map(d["map_init"]); e2.set_poses(d["x_init"])
    for k in range(6):
        e2.iterate(None, odo[:, 0], 1)
        maps.append(e2.get_map())
        ref = orc.calc_cambio(maps[-1], maps[-2])
        assert np.max(np.abs(np.array(ref) - cam[k])) <= 1e-9, (k, ref, cam[k])
    assert np.array_equal(e.get_map(), e2.get_map()) and np.array_equal(e.get_poses(), e2.get_poses())
    e.set_map(d["map_init"]); e.set_poses(d["x_init"])
    tol = float(cam[2, 1]) * 1.0001
    expect = int(np.argmax(cam[:, 1] <= tol)) + 1          # the first pass whose largest landmark change is within tol
    n2, cam2 = e.iterate_until(odo[:, 0], 6, tol)
    assert n2 == expect <= 3 and np.allclose(cam2, cam[:expect], rtol=0, atol=1e-12)
    e.close(); e2.close()


# ---- C5 as ONE handle: K trajectories laid end to end (icmslam_set_batch) == the same trajectories run one at a time -----------
def test_concatenated_batch_equals_individual_runs():
    from icm_slam_b200.batch import ConcatBatch
    from icm_slam_b200.synthetic import make_synthetic_loop
    cfgd = dict(L=2048, cota=20.0)      # (room for the labels the scans of trajectory 2 create every sweep)
    cfg = _cfg(**cfgd)
    K, Tk, nsw = 7, 1024, 5
    data = [make_synthetic_loop(16, T=Tk, seed=20181 + 200 + i) for i in range(K)]
    data[2]["map_init"] = np.delete(data[2]["map_init"], 5, axis=1)            # a missing landmark: far observations, new labels
    data[4]["observations"][:, 300:310] = 10.0                                 # empty scans inside a trajectory
    b = ConcatBatch(cfg, 0, 1, device=0)
    for i, d in enumerate(data):
        b.add(i, d["observations"], d["odometry"], d["velocities"], d["map_init"], d["x_init"])
    b.finalize()
    b.iterate(nsw)
    got = b.results()
    b.close()
    for i, d in enumerate(data):
        e = _engine(cfg, d["observations"], d["odometry"], d["velocities"])
        e.set_map(d["map_init"])
        e.set_poses(d["x_init"])
        e.iterate(None, d["odometry"][:, 0], nsw)
        x1, m1 = e.get_poses(), e.get_map()
        e.close()
        dx = np.abs(got[i][0] - x1)
        assert dx[:2].max() <= 1e-9 and dx[2].max() <= 1e-11, (i, dx.max(axis=1))
        assert got[i][1].shape == m1.shape, (i, got[i][1].shape, m1.shape)
        assert np.abs(got[i][1] - m1).max() <= 1e-9, i


# ---- host-memory sweeps of a map chain: the chunk-pipelined path equals the one-piece path (bit for bit) ---------------------------
def _host_chain(z, odo, u, cfgd, map0, x_init, nsweeps, env):
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        e = _engine(_cfg(**cfgd), z, odo, u)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    x = np.ascontiguousarray(x_init.copy())
    mapa = np.ascontiguousarray(map0.copy())
    e.landmarks_actuales = mapa.shape[1]
    out = []
    for k in range(nsweeps):
        st, Lout, mapa = e.sweep(mapa, x, odo[:, 0])
        assert st == 0
        out.append((x.copy(), np.array(mapa, copy=True), e.associations().copy()))
    e.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", ["2", "4"])
def test_host_sweeps_in_chunks_equal_one_piece(chunks):
    g = golden("synth_b.npz")
    z, odo, u = g["observations"].astype(np.float64), g["odometry"], g["velocities"]
    cfgd = dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"]))
    a = _host_chain(z, odo, u, cfgd, g["map_init"], g["x_init"], 6, {"ICMSLAM_PIPE_CHUNKS": "1"})
    b = _host_chain(z, odo, u, cfgd, g["map_init"], g["x_init"], 6, {"ICMSLAM_PIPE_CHUNKS": chunks, "ICMSLAM_PIPE_TILES": "1"})
    for k, (ra, rb) in enumerate(zip(a, b)):
        assert np.array_equal(ra[2], rb[2]), k
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), k


@pytest.mark.gpu
def test_chunked_sweeps_on_a_longer_trajectory():
    """71 tiles, two landmarks missing from the initial map (labels are created in several chunks): host sweeps in 5 chunks equal
    the one-piece host sweeps and the device-resident sweeps bit for bit."""
    from icm_slam_b200.synthetic import make_synthetic
    d = make_synthetic(900, T=9000, seed=20181 + 21)
    z, odo, u = d["observations"], d["odometry"], d["velocities"]
    map0 = np.delete(d["map_init"], [17, 640], axis=1)
    cfgd = dict(CONFIG_ROS, L=2 * 900 + 64, cota=20.0)
    a = _chain(z, odo, u, cfgd, map0, d["x_init"], 6, {})
    ha = _host_chain(z, odo, u, cfgd, map0, d["x_init"], 5, {"ICMSLAM_PIPE_CHUNKS": "1"})
    hb = _host_chain(z, odo, u, cfgd, map0, d["x_init"], 5, {"ICMSLAM_PIPE_CHUNKS": "5", "ICMSLAM_PIPE_TILES": "8"})
    for k, (ra, rb) in enumerate(zip(ha, hb)):
        assert np.array_equal(ra[2], rb[2]), k
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1]), k
    for k in range(5):      # and the host path equals the device-resident one
        assert np.array_equal(ha[k][0], a[k][0]) and np.array_equal(ha[k][1], a[k][1]), k


# ---- the reference's user-configurable model functions, one pose at a time --------------------------------------
@pytest.mark.gpu
def test_model_hooks_match_the_oracle():
    from oracle import oracle as orc
    from icm_slam_b200.icm import ICM_SLAM
    g = golden("c1_ref.npz")
    z, odo, u = c1_inputs()
    cfg = _cfg(**CONFIG_ROS)
    _, ocfg, ext = _oracle(CONFIG_ROS, z, odo, u)
    icm = ICM_SLAM(cfg, x0=odo[:, 0])
    icm.mediciones, icm.odometria, icm.u = z, odo, u
    x = np.ascontiguousarray(g["p0_x"].copy())
    mapa = g["p0_map"]
    rng = np.random.default_rng(5)
    e = icm.engine
    T = z.shape[1]
    checked = 0
    for t in list(rng.integers(1, T - 1, 12)) + [T - 1]:
        o0, o1 = int(ext["off"][t]), int(ext["off"][t + 1])
        if o1 == o0:
            continue
        d, ang = ext["d"][o0:o1], ext["beam"][o0:o1] * np.pi / 180.0
        zt = np.stack([d, ang], axis=1)
        seen = mapa[:, rng.integers(0, mapa.shape[1], o1 - o0)].T.copy()       # any matched landmark per observation
        xq = x[:, t] + rng.normal(0.0, 0.01, 3)
        icm.t, icm.xt, icm.medicion_actual, icm.mapa_visto = int(t), x[:, t - 1], zt, seen
        # g and h
        assert np.allclose(icm.g(x[:, t - 1], u[:, t - 1]).reshape(3),
                           x[:, t - 1] + cfg.deltat * np.array([np.cos(x[2, t - 1]) * u[0, t - 1], np.sin(x[2, t - 1]) * u[0, t - 1], u[1, t - 1]]), rtol=0, atol=1e-15)
        alfa = ang + xq[2] - np.pi / 2.0
        dist = np.stack([xq[0] + d * np.cos(alfa), xq[1] + d * np.sin(alfa)], axis=1) - seen
        h_np = float(np.sum(np.matmul(dist, cfg.Q) * dist))
        assert abs(icm.h(xq, zt) - h_np) <= 1e-12 * max(1.0, abs(h_np))
        if t + 1 < T:
            icm.x_pos = x[:, t + 1]
            f_o = orc.fun_xn(ocfg, xq, x[:, t - 1], x[:, t + 1], u[:, t - 1], u[:, t], odo[:, t - 1:t + 2], d, ang, seen[:, 0], seen[:, 1])
            assert abs(icm.fun_xn(xq) - f_o) <= 1e-12 * max(1.0, abs(f_o)), t
            for solver, op in (("nm", "min_nm"), ("newton", "min_newton")):
                xo, nev = orc.solve_pose(ocfg, solver, x[:, t - 1], x[:, t + 1], u[:, t - 1], u[:, t], odo[:, t - 1:t + 2], d, ang, seen[:, 0], seen[:, 1])
                xg, fg, ng = e.pose_eval(op, z=zt, seen=seen, x_ant=x[:, t - 1], x_pos=x[:, t + 1], u_ant=u[:, t - 1], u_act=u[:, t],
                                         odo=odo[:, t - 1:t + 2], newton_tol=1e-14, newton_maxit=50)
                dd = np.abs(xg - xo)
                assert dd[:2].max() <= 1e-6 and dd[2] <= 1e-8, (t, solver, dd)
                if solver == "nm":
                    assert ng == nev, (t, ng, nev)
        else:
            f_o = orc.fun_x(ocfg, xq, x[:, t - 1], u[:, t - 1], odo[:, t - 1:t + 1], d, ang, seen[:, 0], seen[:, 1])
            assert abs(icm.fun_x(xq) - f_o) <= 1e-12 * max(1.0, abs(f_o))
            xo, nev = orc.solve_pose(ocfg, "nm", x[:, t - 1], None, u[:, t - 1], None, odo[:, t - 1:t + 1], d, ang, seen[:, 0], seen[:, 1])
            xg, fg, ng = e.pose_eval("min_nm", z=zt, seen=seen, x_ant=x[:, t - 1], u_ant=u[:, t - 1], odo=odo[:, t - 1:t + 1])
            dd = np.abs(xg - xo)
            assert dd[:2].max() <= 1e-6 and dd[2] <= 1e-8 and ng == nev, (t, dd, ng, nev)
        checked += 1
    assert checked >= 8
    with pytest.raises(Exception):
        e.pose_eval("g", x_ant=np.zeros(3), u_ant=np.zeros(2), model=1)       # only the reference's model exists


# ---- ROS ingestion end to end: a log replayed through the in-process publisher gives what the offline path gives
@pytest.mark.gpu
def test_online_ingestion_equals_offline_run():
    from icm_slam_b200 import ros_ingest as ri
    from icm_slam_b200.icm import ICM_SLAM, Mapa
    z, odo, u = c1_inputs()
    T = 400
    z, odo, u = z[:, :T], odo[:, :T], u[:, :T]
    cfgd = dict(CONFIG_ROS, N=3, topic_laser="/scan", topic_laser_msg="sensor_msgs/LaserScan", topic_odometry="/odom",
                topic_odometry_msg="nav_msgs/Odometry")
    cfg = _cfg(**cfgd)
    raw = np.where(z >= cfg.rango_laser_max, np.nan, z - cfg.radio)          # ranges as a lidar reports them (no return: NaN)
    online = ri.OnlineICM(cfg)
    client = ri.FakeRos()

    def pump(node):
        ri.publish_log(client, cfg, raw, odo, u)
        client.service("/icm_slam/iterative_flag").call({"data": True})       # "start iterating"
        return False

    mapa0, x0s = online.inicializar_online(client, pump)
    assert online.iterations_flag and online.mediciones.shape == (180, T)
    m_on, x_on = online.iterar(x0s.copy())
    # the same arrays through the offline surface
    off = ICM_SLAM(cfg, x0=np.array([online.odometria[:, 0]]).T)
    off.load_data(Mapa(cfg), online.mediciones.copy(), online.u.copy(), online.odometria.copy())
    mapa1, x1 = off.inicializar()
    assert np.array_equal(mapa0, mapa1) and np.array_equal(x0s, x1)
    m_off, x_off = off.iterar(x1.copy())
    assert np.array_equal(m_on, m_off) and np.array_equal(x_on, x_off)
    assert m_on.shape[1] >= 3


@pytest.mark.gpu
def test_a_map_set_again_starts_a_clean_chain():
    """A handle that has run a chain (steady tail armed, grid bookkeeping of ITS last build) and is then given a map with the same
    number of landmarks must behave like a fresh handle given that map."""
    g = golden("synth_b.npz")
    z, odo, u = g["observations"].astype(np.float64), g["odometry"], g["velocities"]
    cfgd = dict(L=int(g["cfg_L"]), cota=float(g["cfg_cota"]))
    e1 = _engine(_cfg(**cfgd), z, odo, u)
    e1.set_map(g["map_init"])
    e1.set_poses(np.ascontiguousarray(g["x_init"].copy()))
    e1.iterate(None, odo[:, 0], 8)
    m8, x8 = e1.get_map().copy(), e1.get_poses().copy()
    assert e1.sweep_stats()["steady_sweeps"] > 0          # the chain did reach the steady tail
    shift = np.zeros_like(m8)
    shift[0] += 0.01                                      # the same landmarks, moved within the grid's margin
    for mapa in (m8, m8 + shift):
        e1.set_map(mapa)
        e1.set_poses(x8.copy())
        e1.iterate(None, odo[:, 0], 4)
        e2 = _engine(_cfg(**cfgd), z, odo, u)
        e2.set_map(mapa)
        e2.set_poses(x8.copy())
        e2.iterate(None, odo[:, 0], 4)
        assert np.array_equal(e1.associations(), e2.associations())
        assert np.array_equal(e1.get_poses(), e2.get_poses()) and np.array_equal(e1.get_map(), e2.get_map())
        e2.close()
    e1.close()
