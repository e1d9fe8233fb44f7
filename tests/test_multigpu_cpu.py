"""Host logic of the time-segment partition (icm_slam_b200/multigpu.py) on CPU: two gloo ranks exchange the
same records / statistics the GPUs exchange over NCCL, on data produced by the CPU oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import CONFIG_ROS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_segments_properties():
    from icm_slam_b200.multigpu import plan_segments, local_columns
    rng = np.random.default_rng(3)
    for T, world in [(1833, 2), (1833, 8), (100000, 8), (64, 4), (10, 2)]:
        w = rng.integers(0, 40, T)
        seg = plan_segments(T, world, w)
        assert seg[0][0] == 0 and seg[-1][1] == T and len(seg) == world
        assert all(a[1] == b[0] for a, b in zip(seg[:-1], seg[1:]))
        assert all(s[0] % 2 == 0 and s[1] > s[0] for s in seg)
        loads = [w[a:b].sum() for a, b in seg]
        if T >= 1000:
            assert max(loads) <= 1.25 * (w.sum() / world) + 80
        for r in range(world):
            c_lo, c_hi, t_lo, t_hi = local_columns(seg, r, world)
            assert c_lo >= 0 and c_hi <= T and t_lo % 2 == 0
            assert (t_lo == 0) == (r == 0) and (c_hi == seg[r][1]) == (r == world - 1)
    with pytest.raises(ValueError):
        plan_segments(5, 4)


def test_label_bases_and_halo_records():
    from icm_slam_b200.multigpu import label_bases, make_record, halo_from_records
    base, total = label_bases([3, 0, 5, 1])
    assert list(base) == [0, 3, 3, 8] and total == 9
    recs = [make_record([r, 1, 2], [r, 3, 4], [r, 5, 6], r) for r in range(3)]
    left, right = halo_from_records(recs, 1, 3)
    assert np.array_equal(left, np.array([[0, 0], [3, 5], [4, 6]])) and np.array_equal(right, [2, 1, 2])
    assert halo_from_records(recs, 0, 3)[0] is None and halo_from_records(recs, 2, 3)[1] is None


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from icm_slam_b200.multigpu import (plan_segments, local_columns, label_bases, make_record, gather_records,
                                            reduce_statistics, halo_from_records, SEG_REC)
        from icm_slam_b200.synthetic import make_synthetic
        from oracle import oracle as orc
        d = make_synthetic(256, T=1500, seed=20181 + 5)
        cfgd = dict(CONFIG_ROS, L=2048, cota=20.0)      # room for one new label per scan
        ocfg = orc.make_cfg(**cfgd)
        z, odo, u = d["observations"], d["odometry"], d["velocities"]
        ext = orc.extract_all(orc.precondition(z, ocfg.radio, ocfg.rango_laser_max), ocfg)
        m = orc.Mapa(ocfg)
        map0 = d["map_init"].copy()
        map0[:, ::23] += 5.0                      # push some landmarks away: scans now create new labels
        L0 = map0.shape[1]
        m.landmarks_actuales = L0
        x = np.ascontiguousarray(d["x_init"].copy())
        full = orc.sweep(ocfg, m, ext, odo, u, odo[:, 0], map0, x, "redblack", "newton", "prev")
        T = x.shape[1]
        off, c = ext["off"], full["c"]
        seg = plan_segments(T, world, np.diff(off))
        g_lo, g_hi = seg[rank]
        # this rank's share of what the fused kernel produces on its segment
        obs = slice(off[g_lo], off[g_hi])
        lab = c[obs]
        cnt = torch.from_numpy(np.bincount(lab[lab < L0], minlength=int(ocfg.L)).astype(np.int32))
        far_scans = [t for t in range(g_lo, g_hi) if off[t + 1] > off[t] and (c[off[t]:off[t + 1]] >= L0).any()]
        rec = torch.from_numpy(make_record(x[:, g_lo], x[:, max(g_hi - 2, g_lo)], x[:, g_hi - 1], len(far_scans)))
        allrec = gather_records(rec).numpy()
        assert allrec.shape == (world, SEG_REC)
        base, total = label_bases(allrec[:, 9])
        # global numbering of the new labels == the single-process numbering
        if far_scans:
            t = far_scans[0]
            assert c[off[t]:off[t + 1]].max() == L0 + base[rank]
        assert total == full["raw_L"] - L0
        # halo poses
        left, right = halo_from_records(allrec, rank, world)
        if rank > 0:
            assert np.array_equal(left, x[:, g_lo - 2:g_lo])
        if rank < world - 1:
            assert np.array_equal(right, x[:, g_hi])
        # statistics: integer sums are exact for any partition
        fixed = torch.from_numpy(np.round(np.bincount(lab[lab < L0], weights=np.arange(lab.size)[lab < L0] % 97, minlength=int(ocfg.L))).astype(np.int64))
        newl = torch.zeros(2 * int(ocfg.L), dtype=torch.float64)
        for k, t in enumerate(far_scans):
            newl[L0 + base[rank] + k] = float(t)
        reduce_statistics([cnt, fixed, newl])
        assert np.array_equal(cnt.numpy()[:L0], full["raw_counts"][:L0].astype(np.int32))
        lab_all = c
        want = np.round(np.bincount(lab_all[lab_all < L0], weights=np.concatenate(
            [np.arange(off[b] - off[a]) % 97 for a, b in seg])[lab_all < L0], minlength=int(ocfg.L))).astype(np.int64)
        assert np.array_equal(fixed.numpy(), want)
        all_far = [t for t in range(T) if off[t + 1] > off[t] and (c[off[t]:off[t + 1]] >= L0).any()]
        assert np.array_equal(newl.numpy()[L0:L0 + total], np.array(all_far, dtype=np.float64))
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


# ---- C5: a batch of independent trajectories shards round-robin with no data-path collective ----------------------
def test_trajectory_batch_sharding_is_a_partition():
    from icm_slam_b200.batch import owned_indices
    for n, world in ((4096, 8), (10, 3), (5, 8), (0, 2)):
        parts = [owned_indices(n, r, world) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        owned_indices(4, 2, 2)


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from icm_slam_b200.batch import owned_indices, gather_results
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    local = {i: (np.full((3, 4), float(i)), np.full((2, 2), -float(i))) for i in owned_indices(7, rank, world)}
    allr = gather_results(local)
    q.put((rank, sorted(allr), all(float(allr[i][0][0, 0]) == float(i) for i in allr)))
    dist.destroy_process_group()


def test_trajectory_batch_gather_two_ranks_gloo():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 37
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, keys, ok in got:
        assert keys == list(range(7)) and ok


# ---- the statistics exchange sums ONE block of int64 words in which new-label means (fp64) share words with the old labels'
#      fixed-point sums: a double contributed by exactly one rank survives an integer sum with the other ranks' zero words ----
def _alias_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    L, ls = 64, 40                                   # 40 old labels (int64 sums from every rank), 24 slots for new labels
    rng = np.random.default_rng(100 + rank)
    words = np.zeros(2 * L, dtype=np.int64)
    words[:ls] = rng.integers(-2**45, 2**45, ls)      # fsum_x of the old labels
    words[L:L + ls] = rng.integers(-2**45, 2**45, ls)
    mine = np.arange(ls + rank, L, world)             # new labels are numbered disjointly across ranks
    words.view(np.float64)[mine] = rng.normal(size=mine.size) * 50.0
    words.view(np.float64)[L + mine] = rng.normal(size=mine.size) * 50.0
    t = torch.from_numpy(words.copy())
    dist.all_reduce(t)
    q.put((rank, words, t.numpy().copy(), mine))
    dist.destroy_process_group()


def test_exchange_block_int64_sum_preserves_disjoint_doubles_gloo():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + 71
    world = 2
    ps = [ctx.Process(target=_alias_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = sorted([q.get(timeout=120) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    L, ls = 64, 40
    total = got[0][2]
    assert np.array_equal(total, got[1][2])
    assert np.array_equal(total[:ls], got[0][1][:ls] + got[1][1][:ls])                      # old labels: exact integer sums
    assert np.array_equal(total[L:L + ls], got[0][1][L:L + ls] + got[1][1][L:L + ls])
    for rank, words, _, mine in got:                                                        # new labels: the owner's double, bit for bit
        assert np.array_equal(total.view(np.float64)[mine], words.view(np.float64)[mine])
        assert np.array_equal(total.view(np.float64)[L + mine], words.view(np.float64)[L + mine])
