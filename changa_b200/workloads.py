"""Synthetic inputs for the gravity hot path.

A workload is a dict in the shape DataManager/TreePiece hand to the GPU entry
points after serialisation:
  parts    (Np,5) float64  {mass, soft, x, y, z}        CompactPartData order
  moments  (Nn,27) float64 CudaMultipoleMoments order
  cell     (ilist (Li,2) int32 {index, offsetID}, markers (nb+1), starts (nb), sizes (nb))
  part     same, entries index particles (already expanded per source particle)
  fperiod  scalar period applied on all axes
  ewald    None | {root (27), momc (32), ewt (nh,5), L, fEwCut, nReps, active}
Arrays are float64 here; the C-ABI front end converts to cudatype on staging.
"""
import numpy as np

from .hostcuda import encode_offset


def _ragged(rng, nb, lo, hi):
    return rng.integers(lo, hi + 1, nb)


def random_workload(seed=0, n_buckets=64, max_bucket=12, n_cells=400, cell_len=(0, 200), part_len=(0, 120),
                    periodic=True, min_bucket=1):
    """Random buckets, cells and lists with no tree behind them: exercises every
    code path of the list kernels (ragged and empty lists, self pairs, both
    spline branches, replica offsets, multi-pass buckets) at sizes a CPU check
    finishes instantly."""
    rng = np.random.default_rng(seed)
    sizes = _ragged(rng, n_buckets, min_bucket, max_bucket).astype(np.int32)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32)
    n = int(sizes.sum())
    centres = rng.uniform(-0.45, 0.45, (n_buckets, 3))
    pos = np.repeat(centres, sizes, axis=0) + rng.normal(0, 0.01, (n, 3))
    parts = np.column_stack([rng.uniform(0.5, 1.5, n) / n, rng.uniform(0.004, 0.02, n), pos])

    mom = np.zeros((n_cells, 27))
    mom[:, 0] = rng.uniform(0.01, 0.03, n_cells)            # radius
    mom[:, 1] = rng.uniform(0.004, 0.02, n_cells)           # soft
    mom[:, 2] = rng.uniform(0.5, 2.0, n_cells) * 8 / n      # mass
    mom[:, 3:6] = rng.uniform(-0.5, 0.5, (n_cells, 3))      # cm
    mom[:, 6:] = rng.normal(0, 0.15, (n_cells, 21)) * mom[:, 2:3]

    fperiod = 1.0 if periodic else 0.0
    reps = (-1, 0, 1) if periodic else (0,)

    cl, cm_ = [], [0]
    for b in range(n_buckets):
        want = int(rng.integers(cell_len[0], cell_len[1] + 1))
        idx = rng.integers(0, n_cells, want)
        off = rng.choice(reps, (want, 3))
        d = mom[idx, 3:6] + off * fperiod - centres[b]
        keep = np.sqrt((d ** 2).sum(1)) > 0.12                # well-separated: expansion converges
        idx, off = idx[keep], off[keep]
        codes = np.array([encode_offset(int(o[0]), int(o[1]), int(o[2]), b & 0x3fffff) for o in off], dtype=np.int64)
        cl.append(np.column_stack([idx, codes]) if len(idx) else np.zeros((0, 2), dtype=np.int64))
        cm_.append(cm_[-1] + len(idx))
    cell_list = np.concatenate(cl).astype(np.int64)
    cell_list = (cell_list & 0xffffffff).astype(np.uint32).view(np.int32).reshape(-1, 2)

    pl, pm = [], [0]
    for b in range(n_buckets):
        want = int(rng.integers(part_len[0], part_len[1] + 1))
        own = np.arange(starts[b], starts[b] + sizes[b])     # includes self pairs (r = 0 -> skipped)
        others = rng.integers(0, n, max(want - len(own), 0))
        idx = np.concatenate([own, others])[:max(want, 0)] if want else np.zeros(0, dtype=np.int64)
        off = rng.choice(reps, (len(idx), 3))
        off[: min(len(own), len(idx))] = 0
        codes = np.array([encode_offset(int(o[0]), int(o[1]), int(o[2])) for o in off], dtype=np.int64)
        pl.append(np.column_stack([idx, codes]) if len(idx) else np.zeros((0, 2), dtype=np.int64))
        pm.append(pm[-1] + len(idx))
    part_list = np.concatenate(pl).astype(np.int64)
    part_list = (part_list & 0xffffffff).astype(np.uint32).view(np.int32).reshape(-1, 2)

    def filled(ilist, markers):
        """serialize() drops buckets with empty lists (Compute.cpp:1075-1100)"""
        markers = np.asarray(markers)
        ln = np.diff(markers)
        k = ln > 0
        m = np.concatenate([[0], np.cumsum(ln[k])]).astype(np.int32)
        return (np.ascontiguousarray(ilist, dtype=np.int32), m, starts[k].copy(), sizes[k].copy())

    return {
        "parts": parts, "moments": mom, "fperiod": fperiod,
        "cell": filled(cell_list, cm_), "part": filled(part_list, pm),
        "ewald": None, "name": f"random(seed={seed},nb={n_buckets},maxb={max_bucket})",
    }


def interaction_counts(wl):
    """pair interactions the way ChaNGa counts them (Compute.cpp:1643-1651):
    list entries x target particles of the bucket."""
    out = {}
    for key in ("cell", "part"):
        if wl.get(key):
            _, m, _, sz = wl[key]
            out[key] = int((np.diff(m).astype(np.int64) * sz.astype(np.int64)).sum())
        else:
            out[key] = 0
    return out
