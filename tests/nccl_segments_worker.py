"""torchrun worker of tests/test_gpu_scale.py::test_two_nccl_ranks_equal_the_single_engine_bit_for_bit.

Every rank runs the time-segmented sweep over REAL NCCL collectives (SegmentedSolver, no stream pre-binding by the caller);
rank 0 also runs the same sweeps on one engine and writes both result hashes (poses + map + labels) to argv[1]."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def digest(x, mapa):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(x, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(mapa, dtype=np.float64).tobytes())
    return h.hexdigest()


def main():
    if os.environ.get("ICMSLAM_TEST_WATCHDOG"):      # (debugging aid: dump every thread's stack and exit instead of hanging)
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["ICMSLAM_TEST_WATCHDOG"]), exit=True)
    import torch
    import torch.distributed as dist
    from helpers import CONFIG_ROS
    from icm_slam_b200.config import ConfigICM
    from icm_slam_b200.engine import Engine
    from icm_slam_b200.multigpu import SegmentedSolver
    from icm_slam_b200.synthetic import make_synthetic

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L_true, T, nsweeps = 900, 9000, 7
    d = make_synthetic(L_true, T=T, seed=20181 + 21)
    # two landmarks missing from the initial map: their observations are far, scans create labels on both segments
    map0 = np.delete(d["map_init"], [17, 640], axis=1)
    cfg = ConfigICM.from_values(**dict(CONFIG_ROS, L=2 * L_true + 64, cota=20.0))
    res = {}
    sols = []
    for exchange in ("p2p", "nccl", "fallback"):   # the library's peer-memory kernels / NCCL collectives between the segment calls /
        if exchange == "fallback":                 # ... and p2p requested but one rank cannot: every rank must end up on NCCL
            os.environ["ICMSLAM_P2P_FAIL"] = "1"
        sol = SegmentedSolver(cfg, rank, world, device=local, exchange="p2p" if exchange == "fallback" else exchange)
        sol.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
        sol.set_map(map0)
        sol.set_poses(d["x_init"])
        sol.sweep(3)                         # eager sweeps, then the captured graphs
        sol.sweep(nsweeps - 3)
        torch.cuda.synchronize()
        x_seg = sol.gather_poses()
        m_seg = sol.get_map()
        res["segmented_" + exchange] = digest(x_seg, m_seg)
        if exchange == "fallback":
            os.environ.pop("ICMSLAM_P2P_FAIL", None)
            res["fallback_exchange"] = sol.exchange
        sols.append(sol)
    res["segmented"] = res["segmented_p2p"]
    if rank == 0:
        e = Engine(cfg, device=local)
        e.load(d["observations"], d["odometry"], d["velocities"], precondition=True)
        e.extract()
        e.set_map(map0)
        e.set_poses(d["x_init"])
        e.iterate(None, d["odometry"][:, 0], nsweeps)
        res["single"] = digest(e.get_poses(), e.get_map())
        res["max_pose_diff"] = float(np.max(np.abs(e.get_poses() - x_seg)))
        with open(sys.argv[1], "w") as f:
            json.dump(res, f)
        e.close()
    dist.barrier()
    for sol in sols:
        sol.close()                  # (drops the captured graphs and the solver's own halo group before the default group goes)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
