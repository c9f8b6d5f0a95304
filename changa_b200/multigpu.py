"""Host logic of the multi-GPU force step: one process per GPU, buckets cut into
contiguous SFC ranges, particle and moment records replicated with one all-gather each per step.
Backend-agnostic (`torch.distributed`: nccl on the GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_rows(a, rank, world):
    """equal, zero-padded row slices of a 2-D array: what each rank contributes to the all-gather.
    Returns (my_rows, rows_per_rank)."""
    a = np.ascontiguousarray(a)
    chunk = -(-a.shape[0] // world)
    padded = np.zeros((chunk * world,) + a.shape[1:], dtype=a.dtype)
    padded[: a.shape[0]] = a
    return padded[rank * chunk:(rank + 1) * chunk].copy(), chunk


def gather_rows(dist, torch, mine, world, out=None):
    """one all_gather_into_tensor; `mine` is this rank's (chunk, cols) tensor"""
    if out is None:
        out = torch.empty((mine.shape[0] * world,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, mine)
    return out


def bucket_cuts_by_particles(bucket_sizes, world):
    """contiguous bucket ranges holding (nearly) equal particle counts; never splits a bucket"""
    csum = np.cumsum(bucket_sizes)
    n = int(csum[-1])
    cuts = np.searchsorted(csum, np.arange(1, world) * n / world, side="left") + 1
    cuts = np.concatenate([[0], np.minimum(cuts, len(bucket_sizes)), [len(bucket_sizes)]])
    return np.maximum.accumulate(cuts).astype(np.int64)


def bucket_range_by_starts(bucket_starts, n, rank, world):
    """what RawParticleStep does on the device with torch.searchsorted: rank r owns the buckets whose
    first particle lies in [r n / world, (r+1) n / world) -- contiguous in SFC order, never splits a
    bucket, at most one bucket of imbalance.  Returns (b0, b1, p0, p1): bucket and particle ranges."""
    starts = np.asarray(bucket_starts)
    nb = len(starts)
    b0 = 0 if rank == 0 else int(np.searchsorted(starts, rank * n // world, side="left"))
    b1 = nb if rank == world - 1 else int(np.searchsorted(starts, (rank + 1) * n // world, side="left"))
    p0 = int(starts[b0]) if b0 < nb else n
    p1 = int(starts[b1]) if b1 < nb else n
    return b0, b1, p0, p1


def bucket_range_by_active(bucket_starts, markers, n, rank, world):
    """multistep form of bucket_range_by_starts (what RawParticleStep does on the device when rungs are
    given): the cut points are the particles holding the r * nActive / world -th Ewald markers, so every
    rank gets about the same number of ACTIVE particles.  markers: ascending tree-order indices of the
    active particles.  Falls back to the particle-count rule when nothing is active."""
    markers = np.asarray(markers)
    n_act = len(markers)
    if n_act == 0:
        return bucket_range_by_starts(bucket_starts, n, rank, world)
    starts = np.asarray(bucket_starts)
    nb = len(starts)
    at0 = min(n_act - 1, rank * n_act // world)
    at1 = min(n_act - 1, (rank + 1) * n_act // world)
    b0 = 0 if rank == 0 else int(np.searchsorted(starts, markers[at0], side="left"))
    b1 = nb if rank == world - 1 else int(np.searchsorted(starts, markers[at1], side="left"))
    p0 = int(starts[b0]) if b0 < nb else n
    p1 = int(starts[b1]) if b1 < nb else n
    return b0, b1, p0, p1

