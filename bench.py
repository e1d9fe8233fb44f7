#!/usr/bin/env python
"""bench.py -- ICM sweeps/s on synthetic range-bearing data (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c1] [--impl reference]

One "step" = one ICM sweep (iterations_process_offline, sensors.py:125-168) over the whole
trajectory: projection, association, landmark statistics, pose update, map filter.  Scan
extraction (filtrar_z) is sweep-invariant and cached, as SURVEY.md 8(d) defines the metric.

* `value`    : sweeps/s with poses, observations and map resident in HBM (icmslam_iterate).
* `e2e`      : the same sweep through the reference-facing call
               ICM_SLAM.iterations_process_offline(mapa_viejo, x) with HOST numpy buffers:
               poses + map go host->device and come back inside the timed region.
* `roofline` : the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured
               HBM copy bandwidth (MEASURED_PEAKS.json).
* `cpu_baseline` / `--impl reference`: the CPU oracle port of the reference's own semantics
               (sequential Gauss-Seidel + Nelder-Mead + running means) on a bounded prefix.

Under torchrun (N > 1) the trajectory is split into N contiguous time segments, one per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (L_true, T, description)
    "c4": (316 * 316, 1_000_000, "synthetic 2D range-bearing: 1M poses / 99 856 (316^2) landmarks, 181-beam scans"),
    "c3": (100 * 100, 100_000, "synthetic 2D range-bearing: 100k poses / 10k landmarks, 181-beam scans"),
    "c1": (16 * 16, 2048, "synthetic 2D range-bearing: 2048 poses / 256 landmarks (smoke size)"),
}
SEED = 20181
METRIC = "ICM sweeps/sec (1M poses,100k lmk) at 1/2/4/8 B200; % HBM roofline"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config_for(L_true):
    from icm_slam_b200.config import ConfigICM
    return ConfigICM.from_values(N=1, L=2 * L_true, cota=20.0)


def make_data(name):
    from icm_slam_b200.synthetic import make_synthetic
    L_true, T, _ = WORKLOADS[name]
    t0 = time.time()
    d = make_synthetic(L_true, T=T, seed=SEED + {"c1": 1, "c3": 3, "c4": 4}[name])
    d["gen_s"] = time.time() - t0
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [v.strip() for v in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def sweep_bytes(T, n, L):
    """compulsory traffic of one sweep (fp64 SoA, int32 indices)."""
    return 92 * T + 20 * n + 40 * L + 4


# ------------------------------------------------------------------------------------------------
def cpu_reference_arm(name, d, budget_poses=None):
    """Times the CPU oracle in the reference's own mode on a prefix of the workload.  Returns
    (sweeps_per_s_extrapolated, sample description, seconds, P)."""
    from oracle import oracle as orc
    L_true, T, _ = WORKLOADS[name]
    cfg = config_for(L_true)
    ocfg = orc.make_cfg(cfg)
    P = budget_poses or {"c4": 1500, "c3": 12000, "c1": 2048}[name]
    P = min(P, T)
    z = d["observations"][:, :P]
    ext = orc.extract_all(orc.precondition(z, ocfg.radio, ocfg.rango_laser_max), ocfg)
    while P > 2 and ext["off"][P] == ext["off"][P - 1]:      # the reference needs a non-empty last scan
        P -= 1
    if P != z.shape[1]:
        z = d["observations"][:, :P]
        ext = orc.extract_all(orc.precondition(z, ocfg.radio, ocfg.rango_laser_max), ocfg)
    odo = np.ascontiguousarray(d["odometry"][:, :P])
    u = np.ascontiguousarray(d["velocities"][:, :P])
    x = np.ascontiguousarray(d["x_init"][:, :P].copy())
    m = orc.Mapa(ocfg)
    m.landmarks_actuales = d["map_init"].shape[1]
    t0 = time.perf_counter()
    try:
        orc.sweep(ocfg, m, ext, odo, u, odo[:, 0], d["map_init"], x, "sequential", "nm", "running")
    except ValueError:
        pass   # a short prefix may leave no landmark above cota; the timed work is already done
    dt = time.perf_counter() - t0
    sweeps_per_s = 1.0 / (dt * (T / P))
    sample = ("oracle C port of the reference semantics (sequential Gauss-Seidel, Nelder-Mead xtol=1e-3, running means, "
              "brute-force association against all %d landmarks) on the first %d of %d poses (%d observations), %.2f s; "
              "sweeps/s extrapolated linearly to the full trajectory" % (d["map_init"].shape[1], P, T, ext["n"], dt))
    return sweeps_per_s, sample, dt, P


def run_reference(args, rank, world):
    if rank != 0:
        return
    name = args.workload
    d = make_data(name)
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample, dt, P = cpu_reference_arm(name, d, args.ref_poses)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([a for a, _ in vals]))
    L_true, T, desc = WORKLOADS[name]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "sweeps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": desc, "T": T, "L_true": L_true, "seed": SEED},
           "cpu_baseline": {"value": v, "unit": "sweeps/s", "cores": 1, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from icm_slam_b200.config import ConfigICM   # noqa: F401
    from icm_slam_b200.engine import Engine
    from icm_slam_b200.icm import ICM_SLAM, Mapa, precondicionar

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libicmslam has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        from icm_slam_b200 import multigpu
        return multigpu.bench(args, rank, world, local_rank, WORKLOADS, SEED, METRIC, sweep_bytes, peaks, ClockSampler, make_data,
                              config_for)

    name = args.workload
    L_true, T, desc = WORKLOADS[name]
    d = make_data(name)
    cfg = config_for(L_true)
    x0 = d["odometry"][:, 0].copy()
    dev = torch.device("cuda", local_rank)

    # ---- device-resident arm --------------------------------------------------------------------
    eng = Engine(cfg, device=local_rank)
    stream = torch.cuda.Stream(device=dev)          # (the legacy default stream cannot capture CUDA graphs)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    t0 = time.time()
    eng.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
    n = eng.extract()
    prep_s = time.time() - t0
    eng.set_poses(d["x_init"])                 # ICM.positions resident on the device
    eng.set_map(d["map_init"])
    x_dev = None
    mode = dict(schedule="redblack", solver="newton", view="prev")
    for _ in range(args.warmup):
        eng.iterate(x_dev, x0, 1, **mode)
    torch.cuda.synchronize()
    lc0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.iterate(x_dev, x0, 1, **mode)        # steady state: one CUDA-graph replay per sweep
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    clocks = sampler.stop()
    launches = (eng.launch_count() - lc0) // max(args.steps, 1)
    L_now = eng.landmarks_actuales
    # kernel time of the dominant kernels, measured live with CUDA events on the handle's stream in a
    # separate short loop (reading the events synchronises, so it stays out of the timed region)
    kt = []
    for _ in range(max(3, min(args.steps, 10))):
        eng.iterate(x_dev, x0, 1, timing=True, **mode)
        kt.append(eng.kernel_ms())
    k_assoc_ms = float(np.mean([a for a, _ in kt]))
    k_pose_ms = float(np.mean([b for _, b in kt]))
    fused = True
    B_sweep = sweep_bytes(T, n, L_true)
    peak, peak_src = peaks()
    if fused and k_pose_ms > 0.0:
        # split mode: k_sweep_fused (association + moments + landmark statistics; reads poses 24T, offsets 4(T+1), observations
        # 16n, map 16L; writes labels 4n, statistics 24L) and two k_solve_colour launches (poses, odometry, controls in; poses
        # out).  The dominant kernel is reported with ITS share of the sweep's algorithmic bytes.
        dom_name, dom_ms = "k_runs+k_assoc_tiles", k_assoc_ms
        dom_bytes = 24 * T + 4 * (T + 1) + 16 * n + 16 * L_true + 4 * n + 24 * L_true
    elif fused:
        dom_name, dom_ms, dom_bytes = "k_sweep_fused", k_assoc_ms, B_sweep
    else:
        # two kernels share the sweep's bytes: association (reads poses 24T, obs 16n, map; writes c 4n) and the
        # pose kernels (read poses/odometry/controls/obs/c, write poses).  The dominant one is reported.
        b_assoc = 24 * T + 4 * (T + 1) + 16 * n + 16 * L_true + 4 * n
        b_pose = 24 * T + 24 * T + 16 * T + 4 * (T + 1) + 16 * n + 4 * n + 16 * L_true + 24 * T
        if k_assoc_ms >= k_pose_ms:
            dom_name, dom_ms, dom_bytes = "k_assoc", k_assoc_ms, b_assoc
        else:
            dom_name, dom_ms, dom_bytes = "k_pose_colour(x2)", k_pose_ms, b_pose
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    sweep_achieved = B_sweep / (ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(name, {}).get(dom_name)
        except Exception:
            traffic = None

    # ---- end-to-end arm: the reference-facing call with HOST buffers --------------------------------
    icm = ICM_SLAM(cfg, x0=x0)
    icm._engine.close()
    icm._engine = eng                                      # same device dataset; only the call path differs
    icm.mediciones, icm.odometria, icm.u = d["observations"], d["odometry"], d["velocities"]
    icm._loaded = (id(icm.mediciones), id(icm.u), id(icm.odometria), np.shape(icm.mediciones))
    icm.mapa_obj = Mapa(cfg)
    icm.mapa_obj._attach(eng)
    x_host_t = torch.from_numpy(d["x_init"].copy()).pin_memory()
    x_host = x_host_t.numpy()
    mapa = d["map_init"].copy()
    icm.mapa_obj.landmarks_actuales = mapa.shape[1]
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        mapa_w, _ = icm.iterations_process_offline(mapa, x_host)
    torch.cuda.synchronize()
    tb0 = eng.transfer_bytes()                 # bytes the library actually copies (a map fed back unchanged is not re-uploaded)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mapa, _ = icm.iterations_process_offline(mapa, x_host)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    tb1 = eng.transfer_bytes()
    h2d, d2h = tb1[0] - tb0[0], tb1[1] - tb0[1]

    # ---- CPU baseline (rank 0, N = 1): bounded sample -----------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        v, sample, dt, P = cpu_reference_arm(name, d, args.ref_poses)
        cpu = {"value": v, "unit": "sweeps/s", "cores": 1, "kind": "port", "sample": sample}

    out = {
        "metric": METRIC, "value": 1000.0 / ms, "unit": "sweeps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "T": T, "L_true": L_true, "n_obs": int(n), "landmarks_after": int(L_now), "beams": 181,
                   "seed": SEED, "mode": "redblack/newton/prev", "partition": "1 time segment",
                   "l2": "inputs larger than L2 (observations %.0f MB + poses/odometry/controls %.0f MB vs 126 MB L2)"
                         % (16 * n / 1e6, 64 * T / 1e6),
                   "prep_s": {"synthetic_gen": round(d["gen_s"], 2), "load_extract": round(prep_s, 2)}},
        "clocks": clocks,
        "e2e": {"value": 1.0 / e2e_s, "unit": "sweeps/s", "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                "steps": e2e_steps, "call": "ICM_SLAM.iterations_process_offline(mapa_viejo, x) with pinned host numpy buffers"},
        "gpu_launches": int(launches) * args.steps,
        "gpu_launches_per_step": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel_ms": dom_ms, "algorithmic_bytes": int(dom_bytes),
                     "sweep": {"algorithmic_bytes": int(B_sweep), "achieved": sweep_achieved, "frac": sweep_achieved / peak,
                               "kernel_ms": {"assoc_or_fused": k_assoc_ms, "pose": k_pose_ms},
                               "note": "whole sweep (all kernels of the graph replay) against the sweep's algorithmic bytes"}},
        "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-poses", type=int, default=None, help="prefix length for the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-times", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
