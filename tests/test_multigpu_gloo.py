"""The N>1 host logic on CPU: two gloo ranks shard the buckets of one box, all-gather their
particle / moment slices and between them produce exactly the single-process lists."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from changa_b200.multigpu import (shard_rows, gather_rows, bucket_cuts_by_particles, bucket_range_by_starts,
                                  bucket_range_by_active)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from changa_b200.workloads import config_workload, interaction_counts
        wl = config_workload("cube300", n=14 ** 3, bucket_range_of=(rank, world))
        parts = np.ascontiguousarray(wl["parts"], dtype=np.float32)
        mom = np.ascontiguousarray(wl["moments"], dtype=np.float32)
        mine_p, _ = shard_rows(parts, rank, world)
        mine_m, _ = shard_rows(mom, rank, world)
        full_p = gather_rows(dist, torch, torch.from_numpy(mine_p), world).numpy()
        full_m = gather_rows(dist, torch, torch.from_numpy(mine_m), world).numpy()
        ok = np.array_equal(full_p[: len(parts)], parts) and np.array_equal(full_m[: len(mom)], mom)
        ok = ok and not full_p[len(parts):].any() and not full_m[len(mom):].any()
        cnt = interaction_counts(wl)
        t = torch.tensor([cnt["cell"], cnt["part"], len(wl["ewald"]["active"]), wl["cell"][3].sum()], dtype=torch.int64)
        dist.all_reduce(t)
        q.put((rank, ok, wl["bucket_range"], t.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_ranks_cover_the_single_process_step():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from changa_b200.workloads import config_workload, interaction_counts
    wl = config_workload("cube300", n=14 ** 3)
    cnt = interaction_counts(wl)
    assert all(r[1] for r in res)                                   # gathered arrays == replicated arrays
    assert res[0][2][0] == 0 and res[0][2][1] == res[1][2][0] and res[1][2][1] == wl["tree"].num_buckets
    total = res[0][3]
    assert total == [cnt["cell"], cnt["part"], 14 ** 3, 14 ** 3]    # same work, split not duplicated


def test_shard_rows_and_cuts():
    a = np.arange(35, dtype=np.float32).reshape(7, 5)
    pieces = [shard_rows(a, r, 4)[0] for r in range(4)]
    full = np.concatenate(pieces)
    assert full.shape == (8, 5) and np.array_equal(full[:7], a) and not full[7:].any()
    sizes = np.random.default_rng(1).integers(1, 13, 1000)
    for w in (1, 2, 4, 8):
        cuts = bucket_cuts_by_particles(sizes, w)
        assert cuts[0] == 0 and cuts[-1] == 1000 and len(cuts) == w + 1 and np.all(np.diff(cuts) > 0)
        loads = np.add.reduceat(sizes, cuts[:-1])
        assert loads.max() - loads.min() <= 24


def test_bucket_ranges_by_starts_tile_the_box():
    """the device driver's cut rule (RawParticleStep): ranges are contiguous, disjoint, cover every
    bucket and particle, and differ by at most one bucket's particles from n / world"""
    sizes = np.random.default_rng(2).integers(1, 13, 5000)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    n = int(sizes.sum())
    for w in (1, 2, 3, 8):
        r = [bucket_range_by_starts(starts, n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == len(sizes) and r[0][2] == 0 and r[-1][3] == n
        for a, b in zip(r, r[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        for b0, b1, p0, p1 in r:
            assert p1 - p0 == int(sizes[b0:b1].sum())
            assert abs((p1 - p0) - n / w) <= 12


def test_bucket_ranges_by_active_particles():
    """multistep cut rule: contiguous, disjoint, covering, and balanced in ACTIVE particles even when the
    active set is concentrated in a corner of the SFC order"""
    rng = np.random.default_rng(3)
    sizes = rng.integers(1, 13, 5000)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    n = int(sizes.sum())
    active = np.zeros(n, dtype=bool)
    active[: n // 5] = rng.random(n // 5) < 0.9      # dense region: most of the active particles
    active[n // 5:] = rng.random(n - n // 5) < 0.02
    markers = np.nonzero(active)[0]
    for w in (1, 2, 3, 8):
        r = [bucket_range_by_active(starts, markers, n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == len(sizes) and r[0][2] == 0 and r[-1][3] == n
        for a, b in zip(r, r[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        counts = [int(active[p0:p1].sum()) for _, _, p0, p1 in r]
        assert sum(counts) == len(markers)
        assert max(counts) - min(counts) <= 2 * 12 + 1
    # nothing active: the particle-count rule
    assert bucket_range_by_active(starts, [], n, 1, 2) == bucket_range_by_starts(starts, n, 1, 2)

