"""C5 of BASELINE.json / SURVEY.md section 8(e): a batch of INDEPENDENT trajectories (Monte-Carlo multi-start).

The path shards trivially: trajectory i belongs to rank i % world, every trajectory is one libicmslam handle (its own device
buffers, stream and CUDA graphs), and there is no data-path collective -- ranks never exchange anything while sweeping.  The
sweeps of a rank's handles are enqueued back to back on their own streams, so small trajectories overlap on the GPU.

    batch = TrajectoryBatch(config, rank, world)
    for i in batch.owned(n_traj):
        batch.add(i, z_i, odo_i, u_i, map_i, x_i)
    batch.iterate(30)
    res = batch.results()          # {i: (x (3 x T), mapa (2 x L))}
    allres = gather_results(res)   # optional, for output only (torch.distributed all_gather_object)
"""
from __future__ import annotations

import numpy as np

from .engine import Engine


def owned_indices(n_traj: int, rank: int, world: int):
    """Round-robin shard: trajectory i runs on rank i % world (weak scaling: fixed work per trajectory)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, int(n_traj), world))


class TrajectoryBatch:
    def __init__(self, config, rank: int = 0, world: int = 1, device=None):
        self.config, self.rank, self.world = config, int(rank), int(world)
        self.device = int(rank if device is None else device)
        self._eng = {}
        self._x0 = {}

    def owned(self, n_traj: int):
        return owned_indices(n_traj, self.rank, self.world)

    def add(self, idx: int, z, odometria, u, map_init, x_init, precondition=True):
        """Loads trajectory `idx` (raw ranges B x T, odometry 3 x T, controls 2 x T) with its initial map / poses."""
        if idx % self.world != self.rank:
            raise ValueError("trajectory %d belongs to rank %d" % (idx, idx % self.world))
        e = Engine(self.config, device=self.device)
        e.load(z, odometria, u, precondition=precondition)
        e.extract()
        e.set_map(map_init)
        e.set_poses(np.ascontiguousarray(x_init, dtype=np.float64))
        self._eng[idx] = e
        self._x0[idx] = np.ascontiguousarray(np.asarray(odometria, dtype=np.float64)[:, 0].copy())

    def iterate(self, n_sweeps: int = 1, **mode):
        """n_sweeps ICM sweeps of every owned trajectory; sweep k of all handles is enqueued before sweep k+1 (the calls
        are asynchronous: the handles' streams run concurrently)."""
        for _ in range(int(n_sweeps)):
            for idx, e in self._eng.items():
                e.iterate(None, self._x0[idx], 1, **mode)

    def results(self):
        return {idx: (e.get_poses(), e.get_map()) for idx, e in self._eng.items()}

    def n_observations(self):
        return sum(e.n for e in self._eng.values())

    def close(self):
        for e in self._eng.values():
            e.close()
        self._eng.clear()


def gather_results(local: dict, group=None):
    """All ranks' results as one dict (output only -- not part of the timed path)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return dict(local)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, local, group=group)
    out = {}
    for p in parts:
        out.update(p)
    return out
