"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Library versions are recorded inside every fixture (the reference pins numpy 1.19.5 /
scipy 1.5.4, requisitos.txt:13,21; this container has newer ones).

Fixtures:
  c1_inputs.npz / c2_inputs.npz  the two shipped logs (observations stored as float32, which is
                                 exact for both files -- asserted below)
  c1_ref.npz   pass 0 + 2 sweeps of data_IJAC2018.mat through ICM_ROS (config_ros.yaml values)
  c2_ref.npz   pass 0 + 1 sweep of datos_palomar1.mat (raw observations)
  synth_a.npz, synth_b.npz  3 sweeps of small synthetic logs (with injected empty scans,
                                 new labels, a landmark pair that merges in Mapa.filtrar)
  units.npz    function-level vectors: filtrar_z, tras_rot_z, Mapa.actualizar, Mapa.filtrar,
               calc_cambio, fun_x, fun_xn, fmin results
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_runner as rr  # noqa: E402
from icm_slam_b200.synthetic import make_synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
VERS = dict(numpy=np.__version__, scipy=scipy.__version__)


def save(name, **arrs):
    arrs["versions"] = np.array([f"numpy={VERS['numpy']}", f"scipy={VERS['scipy']}"])
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, len(arrs), "arrays")


def pack_sweep(prefix, out, mapa, x, rec, cambio):
    lab = rec["labels"]
    out[prefix + "labels"] = np.concatenate([np.asarray(l, np.int32).ravel() for l in lab])
    out[prefix + "nt"] = np.array([len(l) for l in lab], np.int32)  # per actualizar call (non-empty scans)
    out[prefix + "x_in"] = rec["x_in"]
    out[prefix + "x_out"] = x.copy()
    out[prefix + "map_out"] = mapa.copy()
    out[prefix + "raw_map"] = rec["raw_map"]
    out[prefix + "raw_counts"] = rec["raw_counts"]
    out[prefix + "raw_L"] = np.int32(rec["raw_L"])
    out[prefix + "counts_out"] = rec["counts_out"]
    out[prefix + "nev"] = np.int64(rec["nev"])
    out[prefix + "cambio"] = np.array(cambio)


def run_log(name, z, odo, u, nsweeps, cfg_over=None):
    _, ref_icm = rr.load_reference()
    cfg = rr.make_config(**(cfg_over or {}))
    med = rr.precondition(z, cfg)
    s = rr.make_solver(cfg, med, odo, u)
    t0 = time.time()
    lab0 = []
    p0 = rr.pass0(s, log=lab0)
    print(name, "pass0 %.1fs raw_L=%d L=%d" % (time.time() - t0, p0["raw_L"], p0["mapa"].shape[1]))
    out = dict(p0_map=p0["mapa"], p0_x=p0["x"], p0_raw_L=np.int32(p0["raw_L"]), p0_raw_map=p0["raw_map"][:, :p0["raw_L"]],
               p0_raw_counts=p0["raw_counts"][:p0["raw_L"]],
               p0_labels=np.concatenate([np.asarray(l, np.int32).ravel() for l in lab0]),
               p0_nt=np.array([len(l) for l in lab0], np.int32))
    x = p0["x"].copy()
    m = p0["mapa"].copy()
    for k in range(1, nsweeps + 1):
        t0 = time.time()
        m1, x, rec = rr.sweep(s, m, x, record=True)
        cambio = ref_icm.calc_cambio(m1, m)
        print(name, "sweep %d %.1fs raw_L=%d L=%d nev=%d" % (k, time.time() - t0, rec["raw_L"], m1.shape[1], rec["nev"]))
        pack_sweep("s%d_" % k, out, m1, x, rec, cambio)
        m = m1.copy()
    out["nsweeps"] = np.int32(nsweeps)
    save(name, **out)


def run_synth(name, L_true, T, seed, nsweeps, empties, merge_pair, cfgL=200, cota=20.0, drop=None):
    _, ref_icm = rr.load_reference()
    d = make_synthetic(L_true, T=T, seed=seed)
    z = d["observations"].copy()
    for t in empties:
        z[:, t] = 10.0
    map_init = d["map_init"].copy()
    if merge_pair:
        # duplicate a landmark 0.4 m away so the two share observations and merge in filtrar
        j = merge_pair
        map_init = np.concatenate([map_init, map_init[:, [j]] + np.array([[0.3], [0.25]])], axis=1)
    if drop is not None:  # an unmapped trunk: its returns are "far" and spawn new labels every scan
        map_init = np.delete(map_init, drop, axis=1)
    cfg = rr.make_config(L=cfgL, cota=cota)
    med = rr.precondition(z, cfg)
    s = rr.make_solver(cfg, med, d["odometry"], d["velocities"])
    s.mapa_obj.landmarks_actuales = map_init.shape[1]
    x = d["x_init"].copy()
    m = map_init.copy()
    out = dict(observations=z.astype(np.float32), odometry=d["odometry"], velocities=d["velocities"], x_init=d["x_init"],
               map_init=map_init, cfg_L=np.int32(cfgL), cfg_cota=np.float64(cota))
    assert np.array_equal(out["observations"].astype(np.float64), z)
    for k in range(1, nsweeps + 1):
        t0 = time.time()
        m1, x, rec = rr.sweep(s, m, x, record=True)
        cambio = ref_icm.calc_cambio(m1, m)
        print(name, "sweep %d %.1fs raw_L=%d L=%d nev=%d" % (k, time.time() - t0, rec["raw_L"], m1.shape[1], rec["nev"]))
        pack_sweep("s%d_" % k, out, m1, x, rec, cambio)
        m = m1.copy()
    out["nsweeps"] = np.int32(nsweeps)
    save(name, **out)


def units():
    ref_sensors, ref_icm = rr.load_reference()
    rng = np.random.default_rng(7)
    cfg = rr.make_config()
    out = {}
    # ---- filtrar_z on crafted scans (B=181)
    B = 181
    scans = []
    for k in range(48):
        z = np.full(B, 10.0)
        kind = k % 8
        if kind == 0:   # a few clusters
            for _ in range(rng.integers(1, 6)):
                a = rng.integers(0, B - 6)
                w = rng.integers(1, 6)
                z[a:a + w] = np.float32(rng.uniform(0.7, 9.5)) + rng.normal(0, 0.02, w)
        elif kind == 1:  # single valid beam -> empty
            z[rng.integers(0, B)] = 3.0
        elif kind == 2:  # two isolated far apart -> (0,4)
            z[10] = 5.0
            z[150] = 6.0
        elif kind == 3:  # zeros (coincident points at the origin) + edges
            z[0:3] = 0.0
            z[178:181] = 2.0
            z[90] = 0.0
        elif kind == 4:  # everything valid
            z = rng.uniform(0.5, 9.9, B)
        elif kind == 5:  # salt noise that the median removes, edge beams
            z[rng.integers(0, B, 12)] = rng.uniform(1, 9, 12)
            z[0] = 1.0
            z[1] = 1.1
            z[B - 1] = 2.0
            z[B - 2] = 2.1
        elif kind == 6:  # near the distance threshold
            z[40:42] = [9.0, 9.0]
            z[47:49] = [8.15, 8.15]
            z[100] = 4.0
            z[101] = 4.0
            z[102] = 4.0
            z[109] = 5.2
            z[110] = 5.2
        else:            # dense random valid/invalid
            m = rng.random(B) < 0.3
            z[m] = rng.uniform(0.5, 9.9, int(m.sum()))
        scans.append(np.float32(z).astype(np.float64))
    scans = np.array(scans).T  # B x 48
    med = rr.precondition(scans, cfg)
    fz_off = [0]
    fz_rows = []
    for k in range(med.shape[1]):
        zz = ref_icm.filtrar_z(med[:, k].copy(), cfg)
        zz = zz.reshape(-1, 4) if zz.size else np.zeros((0, 4))
        fz_rows.append(zz)
        fz_off.append(fz_off[-1] + zz.shape[0])
    out["fz_scans"] = med
    out["fz_off"] = np.array(fz_off, np.int32)
    out["fz_rows"] = np.concatenate(fz_rows, axis=0)
    # ---- tras_rot_z
    poses = np.column_stack([rng.uniform(-50, 50, 12), rng.uniform(-50, 50, 12), rng.uniform(-7, 14, 12)])
    tr_in, tr_out = [], []
    for p in poses:
        zz = fz_rows[4].copy()
        tr_in.append(zz.copy())
        tr_out.append(ref_icm.tras_rot_z(p.copy(), zz).copy())
    out["tr_poses"] = poses
    out["tr_in"] = np.array(tr_in)
    out["tr_out"] = np.array(tr_out)
    # ---- Mapa.actualizar (branch B): sequences of calls on one Mapa
    cfg2 = rr.make_config(L=60)
    seqs = 6
    ac = {}
    for q in range(seqs):
        Lref = int(rng.integers(1, 12))
        ref_map = rng.uniform(-10, 10, (2, Lref))
        m = ref_icm.Mapa(cfg2)
        m.landmarks_actuales = Lref
        y = np.zeros((2, cfg2.L))
        ncall = 7
        for r in range(ncall):
            n = int(rng.integers(1, 14))
            pick = rng.integers(0, Lref, n)
            obs = ref_map[:, pick].T + rng.normal(0, 0.45, (n, 2))
            if r % 3 == 1:  # force some far observations
                obs[: max(1, n // 3)] += rng.uniform(3, 6)
            if r == 4 and n > 2:  # exact tie between two reference landmarks is impossible to craft
                obs[0] = obs[1]    # generally; duplicates at least exercise equal rows
            y, c = m.actualizar(y, ref_map, obs.copy())
            ac[f"ac{q}_{r}_obs"] = obs
            ac[f"ac{q}_{r}_c"] = np.asarray(c, np.int32)
            ac[f"ac{q}_{r}_y"] = y.copy()
            ac[f"ac{q}_{r}_cant"] = m.cant_obs_i.copy()
            ac[f"ac{q}_{r}_Lact"] = np.int32(m.landmarks_actuales)
        ac[f"ac{q}_ref"] = ref_map
        ac[f"ac{q}_ncall"] = np.int32(ncall)
    ac["ac_nseq"] = np.int32(seqs)
    out.update(ac)
    # ---- Mapa.filtrar: crafted maps with close pairs / chains, counts around cota
    cfg3 = rr.make_config(L=40, cota=10.0)
    nf = 10
    for q in range(nf):
        m = ref_icm.Mapa(cfg3)
        La = int(rng.integers(3, 30))
        pts = rng.uniform(-15, 15, (2, La))
        # chains: make some landmarks near others
        for _ in range(int(rng.integers(0, 6))):
            a, b = rng.integers(0, La, 2)
            if a != b:
                pts[:, a] = pts[:, b] + rng.uniform(-0.6, 0.6, 2)
        if q == 3 and La > 4:
            pts[:, 2] = pts[:, 4]  # coincident pair (a==0 -> max)
        cnt = rng.integers(0, 40, La).astype(float)
        cnt[rng.integers(0, La)] = 3.0  # guarantee at least one pruned (else the reference raises IndexError)
        cnt[rng.integers(0, La)] = 25.0
        y = np.zeros((2, cfg3.L))
        y[:, :La] = pts
        m.landmarks_actuales = La
        m.cant_obs_i = np.zeros(cfg3.L)
        m.cant_obs_i[:La] = cnt
        try:
            yy = m.filtrar(y.copy())
            ok = 1
        except Exception as e:  # noqa
            print("filtrar case", q, "raised", type(e).__name__, e)
            yy = np.zeros((2, cfg3.L))
            ok = 0
        out[f"fl{q}_in"] = pts
        out[f"fl{q}_cnt"] = cnt
        out[f"fl{q}_ok"] = np.int32(ok)
        out[f"fl{q}_out"] = yy
        out[f"fl{q}_Lact"] = np.int32(m.landmarks_actuales)
        out[f"fl{q}_cant"] = np.asarray(m.cant_obs_i, float).copy()
    out["fl_n"] = np.int32(nf)
    # ---- calc_cambio
    a = rng.uniform(-10, 10, (2, 9))
    b = a[:, :7] + rng.normal(0, 0.05, (2, 7))
    out["cc_new"] = b
    out["cc_old"] = a
    out["cc_out"] = np.array(ref_icm.calc_cambio(b, a))
    # ---- energies and fmin on small self-contained problems
    ne = 40
    E = dict(x=[], x_ant=[], x_pos=[], u=[], odo=[], z=[], seen=[], n=[], fxn=[], fx=[], min_xn=[], min_x=[], nev_xn=[],
             nev_x=[])
    for q in range(ne):
        n = int(rng.integers(1, 18))
        th = rng.uniform(-3, 12)
        x_ant = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), th])
        uu = np.array([[rng.uniform(0, 2.5), rng.uniform(0, 2.5), 0.0], [rng.uniform(-.6, .6), rng.uniform(-.6, .6), 0.0]])
        s = ref_sensors.ICM_ROS.__new__(ref_sensors.ICM_ROS)
        s.config = cfg if q % 4 else rr.make_config(Q=[0.7, 1.9], R=[1.3, 0.6, 2.2], cte_odom=0.45)
        xm = s.g(x_ant, uu[:, 0]).reshape(3) + rng.normal(0, [0.03, 0.03, 0.01])
        x_pos = s.g(xm, uu[:, 1]).reshape(3) + rng.normal(0, [0.03, 0.03, 0.01])
        odo = np.column_stack([x_ant + rng.normal(0, 0.3, 3), xm + rng.normal(0, 0.3, 3), x_pos + rng.normal(0, 0.3, 3)])
        if q % 5 == 0:  # exercise entrepi wrap: odometry heading jumps by 2*pi
            odo[2, 1:] += 2 * np.pi
        dd = rng.uniform(0.8, 9.8, n)
        al = np.sort(rng.integers(0, 181, n)) * np.pi / 180.0
        z = np.column_stack([dd, al])
        alfa = al + xm[2] - np.pi / 2
        seen = np.column_stack([xm[0] + dd * np.cos(alfa), xm[1] + dd * np.sin(alfa)]) + rng.normal(0, 0.08, (n, 2))
        x = xm + rng.normal(0, [0.05, 0.05, 0.02])
        s.u = uu
        s.odometria = odo
        s.t = 1
        s.x_pos = x_pos.reshape(3, 1)
        s.x_ant = x_ant.reshape(3, 1)
        s.xt = x_ant.reshape(3, 1)
        s.medicion_actual = z
        s.mapa_visto = seen
        fxn = float(np.asarray(s.fun_xn(x.copy())).ravel()[0])
        fx = float(np.asarray(s.fun_x(x.copy())).ravel()[0])
        xx = np.column_stack([x_ant, xm, x_pos])
        cnt = [0]
        orig = s.fun_xn
        s.fun_xn = lambda v: (cnt.__setitem__(0, cnt[0] + 1), orig(v))[1]
        mxn = np.asarray(s.minimizar_xn(z, seen, xx, 1)).reshape(3)
        nxn = cnt[0]
        del s.fun_xn
        s.xt = x_ant.reshape(3, 1)
        cnt = [0]
        orig2 = s.fun_x
        s.fun_x = lambda v: (cnt.__setitem__(0, cnt[0] + 1), orig2(v))[1]
        mx = np.asarray(s.minimizar_x(z, seen)).reshape(3)
        nx = cnt[0]
        del s.fun_x
        pad = np.zeros((17, 2))
        pad[:n] = z
        pads = np.zeros((17, 2))
        pads[:n] = seen
        cfgv = np.array([s.config.Q[0, 0], s.config.Q[1, 1], s.config.R[0, 0], s.config.R[1, 1], s.config.R[2, 2], s.config.cte_odom])
        for k, v in dict(x=x, x_ant=x_ant, x_pos=x_pos, u=uu[:, :2], odo=odo, z=pad, seen=pads, n=n, fxn=fxn, fx=fx, min_xn=mxn,
                         min_x=mx, nev_xn=nxn, nev_x=nx).items():
            E[k].append(v)
        E.setdefault("cfgv", []).append(cfgv)
    for k, v in E.items():
        out["en_" + k] = np.array(v)
    save("units.npz", **out)


def inputs():
    z, odo, u = rr.load_ijac()
    assert np.array_equal(z.astype(np.float32).astype(np.float64), z)
    save("c1_inputs.npz", observations=z.astype(np.float32), odometry=odo, velocities=u)
    z2, odo2, u2 = rr.load_palomar()
    assert np.array_equal(z2.astype(np.float32).astype(np.float64), z2)
    assert np.array_equal(odo, odo2) and np.array_equal(u, u2)
    save("c2_inputs.npz", observations=z2.astype(np.float32))  # odometry/velocities identical to c1
    return (z, odo, u), (z2, odo2, u2)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1:] or ["inputs", "units", "synth", "c1", "c2"]
    c1 = c2 = None
    if "inputs" in what or "c1" in what or "c2" in what:
        c1, c2 = inputs()
    if "units" in what:
        units()
    if "synth" in what:
        run_synth("synth_a.npz", 25, 260, 20181 + 100, 3, empties=[5, 6, 100], merge_pair=7)
        run_synth("synth_b.npz", 36, 380, 20181 + 101, 3, empties=[1, 200, 201, 202], merge_pair=0, cfgL=400, cota=12.0, drop=8)
    if "c1" in what:
        run_log("c1_ref.npz", *c1, nsweeps=2)
    if "c2" in what:
        run_log("c2_ref.npz", *c2, nsweeps=1)
