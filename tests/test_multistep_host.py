"""Host-side pieces of the multistep force step that need no GPU: the rung
recipe, the active sets the device code must reproduce, and the walk restricted to them."""
import numpy as np

from changa_b200.workloads import clustered_box, density_rungs
from changa_b200.tree import Tree


def test_density_rungs_follow_the_recipe():
    pos, mass, soft = clustered_box(20000, seed=2, n_halos=12)
    rung = density_rungs(pos)
    assert rung.dtype == np.uint8 and rung.shape == (len(pos),)
    assert rung.min() == 0 and 2 <= rung.max() <= 6
    # rung = clamp(floor(log2(rho / rho_mean) / 2), 0, 6): denser cells never get a lower rung
    g = max(1, int(round((len(pos) / 8.0) ** (1.0 / 3.0))))
    ijk = np.floor((pos + 0.5) * g).astype(np.int64) % g
    cell = (ijk[:, 0] * g + ijk[:, 1]) * g + ijk[:, 2]
    cnt = np.bincount(cell, minlength=g ** 3)[cell]
    order = np.argsort(cnt, kind="stable")
    assert np.all(np.diff(rung[order].astype(int)) >= 0)
    # a uniform box has no deep rungs
    uni = np.random.default_rng(1).uniform(-0.5, 0.5, (20000, 3))
    assert density_rungs(uni).max() <= 1


def test_active_sets_and_masked_walk():
    """what cb200_active_sets_device / cb200_walk_device_active have to reproduce (tests/test_gpu_parity.py
    ::test_multistep_device_path compares the device against exactly this)"""
    pos, mass, soft = clustered_box(6000, seed=3, n_halos=6)
    rung = density_rungs(pos)
    t = Tree(pos, mass, soft, max_bucket=12)
    rs = rung[t.order]
    active_rung = 2
    pact = rs >= active_rung
    bact = np.array([pact[s:s + z].any() for s, z in zip(t.bucket_starts, t.bucket_sizes)])
    assert 0 < bact.sum() < t.num_buckets
    # Ewald markers (Ewald.cpp:416-437): ascending tree-order indices of the active PARTICLES, a subset of the
    # particles of the active BUCKETS (Compute.cpp:1278)
    markers = np.nonzero(pact)[0]
    in_active_bucket = np.repeat(bact, t.bucket_sizes)
    assert in_active_bucket[markers].all() and in_active_bucket.sum() >= len(markers)
    full = t.walk(theta=0.7, n_replicas=1, period=1.0)
    part = t.walk(theta=0.7, n_replicas=1, period=1.0, bucket_active=bact)
    fm, pm = full["cell_mark"], part["cell_mark"]
    assert np.all(np.diff(pm)[~bact] == 0)                       # inactive buckets: empty lists
    mask = ~np.int32((1 << 22) - 1)                               # low offsetID bits name the walk target
    for b in np.nonzero(bact)[0][:200]:
        a = full["cell"][fm[b]:fm[b + 1]]
        c = part["cell"][pm[b]:pm[b + 1]]
        assert np.array_equal(a[:, 0], c[:, 0]) and np.array_equal(a[:, 1] & mask, c[:, 1] & mask)


def test_gpu_local_affinity_is_harmless_without_nvml():
    import os
    import bench
    before = os.sched_getaffinity(0)
    n = bench.gpu_local_affinity(0)
    assert n is None or n == len(os.sched_getaffinity(0))
    os.sched_setaffinity(0, before)
