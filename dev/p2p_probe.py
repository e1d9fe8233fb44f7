"""per-sweep device times of the segmented solver under torchrun (p2p vs nccl exchange): where does a slow step come from?"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from icm_slam_b200.config import ConfigICM
from icm_slam_b200.multigpu import SegmentedSolver
from icm_slam_b200.synthetic import make_synthetic
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = 316 * 316
d = make_synthetic(L, T=1_000_000, seed=20181 + 4)
cfg = ConfigICM.from_values(N=1, L=2 * L, cota=20.0)
stream = torch.cuda.Stream(device=local, priority=-1)
torch.cuda.set_stream(stream)
variants = [a for a in sys.argv[1:]] or ["p2p", "nccl"]
for variant in variants:
    # "p2p", "nccl", or "p2p:ENV=V,ENV=V" (library switches are read when the handle is created)
    exchange, _, envs = variant.partition(":")
    saved = {}
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=")
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    sol = SegmentedSolver(cfg, rank, world, device=local, exchange=exchange)
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    sol.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
    sol.set_map(d["map_init"]); sol.set_poses(d["x_init"])
    for _ in range(8): sol.sweep()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    for rep in range(2):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        host = []
        evs[0].record(stream)
        for k in range(20):
            t0 = time.perf_counter()
            sol.sweep(2 if exchange == "nccl" else 1)
            host.append((time.perf_counter() - t0) * 1e3)
            evs[k + 1].record(stream)
            if exchange == "nccl" and k >= 9: break
        torch.cuda.synchronize()
        nrec = 10 if exchange == "nccl" else 20
        dts = [evs[k].elapsed_time(evs[k + 1]) for k in range(nrec)]
        per = 2 if exchange == "nccl" else 1
        tot = torch.tensor([sum(dts) / (nrec * per), max(dts) / per, max(host)], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("%s rep %d: mean %.4f ms/sweep  worst interval %.4f ms/sweep  worst host call %.3f ms   rank0 intervals: %s" % (
                variant, rep, tot[0].item(), tot[1].item(), tot[2].item(), " ".join("%.3f" % (v / per) for v in dts)), flush=True)
        dist.barrier()
    if os.environ.get("ICMSLAM_TRACE") or "ICMSLAM_TRACE=1" in variant:
        import ctypes as C
        buf = (C.c_uint64 * 256)(); cn = (C.c_uint32 * 3)()
        sol.engine.lib.icmslam_get_trace(sol.engine._h, buf, cn)
        tr = np.array(buf[:], dtype=np.uint64).reshape(32, 8).astype(np.int64)
        k_last = int(cn[0]) - 1                      # last closed sweep
        names = ["runs", "assoc", "labels", "reduce", "steady", "solve", "halo", "steady_end"]
        lines = []
        for k in range(k_last - 5, k_last + 1):
            r, r1 = tr[k & 31], tr[(k + 1) & 31]
            t0 = r[0]
            vals = [r[0], r[1], r[2], r1[3], r[4], r1[5], r1[6], r[7]]
            nxt = tr[(k + 1) & 31][0] if k < k_last else 0
            lines.append("sweep %d: " % k + "  ".join("%s %+.1f" % (n, (v - t0) / 1e3) for n, v in zip(names, vals)) + ("  next_runs %+.1f" % ((nxt - t0) / 1e3) if nxt else ""))
        for rr in range(world):
            if rank == rr:
                print("rank %d trace (us after k_runs entry)\n  " % rank + "\n  ".join(lines), flush=True)
            dist.barrier()
    st = sol.engine.sweep_stats()
    if rank == 0: print(variant, "stats", {k: st[k] for k in ("dirty_tiles", "steady_sweeps", "status")}, flush=True)
    sol.close()
dist.barrier()
import threading
threading.Timer(20.0, lambda: os._exit(0)).start()
dist.destroy_process_group()
