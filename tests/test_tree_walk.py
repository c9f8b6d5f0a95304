"""Host side of the hot path's input (changa_b200/csrc/treewalk.cpp, ewald_tables.py):
tree, moments, interaction lists and Ewald tables against the CPU oracle.  No GPU."""
import os

import numpy as np
import pytest

from changa_b200.tree import Tree, tree_workload
from changa_b200.ewald_tables import ewald_tables
from changa_b200 import workloads
from oracle import oracle as orc
from oracle.walk_oracle import SequentialWalk


def particles(n, seed, clustered=True, soft=None):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-0.5, 0.5, (n, 3))
    if clustered:
        k = n // 3
        pos[:k] = rng.uniform(-0.3, 0.3, 3) + 0.01 * rng.normal(size=(k, 3))
        pos = (pos + 0.5) % 1.0 - 0.5
    mass = rng.uniform(0.5, 1.5, n) / n
    soft = np.full(n, soft if soft is not None else n ** (-1 / 3) / 20)
    return pos, mass, soft


@pytest.fixture(scope="module")
def small_tree():
    pos, mass, soft = particles(3000, 11)
    return Tree(pos, mass, soft, max_bucket=12)


def test_tree_topology(small_tree):
    t = small_tree
    # particles sorted, permutation valid
    assert sorted(t.order.tolist()) == list(range(t.n))
    # breadth-first numbering: children after parents, child0 before child1, levels contiguous
    nn = t.num_nodes
    seen = 1
    for i in range(nn):
        for c in (t.child0[i], t.child1[i]):
            if c >= 0:
                assert c == seen
                seen += 1
    assert seen == nn
    leaf = (t.child0 < 0) & (t.child1 < 0)
    cnt = t.last - t.first + 1
    assert cnt[leaf].max() <= 12 and cnt[~leaf].min() > 12      # Compute.cpp:2476-2477
    # buckets tile the particle array in order
    assert t.bucket_starts[0] == 0 and np.array_equal(t.bucket_starts[1:], np.cumsum(t.bucket_sizes)[:-1])
    assert t.bucket_sizes.sum() == t.n
    # children partition the parent's range; geometric boxes halve along level%3
    for i in np.nonzero(~leaf)[0][:200]:
        kids = [c for c in (t.child0[i], t.child1[i]) if c >= 0]
        assert t.first[kids[0]] == t.first[i] and t.last[kids[-1]] == t.last[i]
    # every particle sits inside its bucket's geometric box and tight box
    for b in range(0, t.num_buckets, 7):
        nd = t.bucket_node[b]
        p = t.parts[t.first[nd]:t.last[nd] + 1, 2:5]
        assert np.all(p >= t.geolo[nd] - 1e-12) and np.all(p <= t.geohi[nd] + 1e-12)
        assert np.allclose(p.min(0), t.boxlo[nd]) and np.allclose(p.max(0), t.boxhi[nd])


def test_moments_equal_oracle_build(small_tree):
    t = small_tree
    want = orc.build_moments(t.parts[:, 2:5], t.parts[:, 0], t.parts[:, 1], t.child0, t.child1, t.first,
                             t.last, t.geolo, t.geohi, t.boxlo, t.boxhi)
    assert np.array_equal(want, t.moments)          # same arithmetic, same order: bit-exact
    assert np.isclose(t.moments[0, 2], t.parts[:, 0].sum())
    cm = (t.parts[:, 0:1] * t.parts[:, 2:5]).sum(0) / t.parts[:, 0].sum()
    assert np.allclose(t.moments[0, 3:6], cm, atol=1e-14)


def _as_i32(rows, k):
    if not len(rows):
        return np.zeros((0, k), dtype=np.int32)
    return np.array(rows, dtype=np.int64).reshape(-1, k).astype(np.uint32).view(np.int32)


@pytest.mark.parametrize("nrep,active_frac,soft", [(0, None, None), (1, None, 0.02), (1, 0.3, 0.02)])
def test_lists_bit_exact_vs_sequential_walk(nrep, active_frac, soft):
    """the parallel recursive walk emits, per bucket, exactly the entries and order of the
    reference's bucket-after-bucket walk (TreeWalk.cpp:297-397 + Compute.cpp:1608-1863)"""
    pos, mass, s = particles(1200, 5, soft=soft)
    t = Tree(pos, mass, s, max_bucket=8)
    act = None
    if active_frac is not None:
        act = np.random.default_rng(9).random(t.num_buckets) < active_frac
    w = t.walk(theta=0.7, n_replicas=nrep, period=1.0, bucket_active=act)
    ref = SequentialWalk(t.child0, t.child1, t.first, t.last, t.boxlo, t.boxhi, t.moments, t.bucket_node,
                         0.7, nrep, 1.0, act).run()
    for b in range(t.num_buckets):
        rc, rp, rs = ref.get(b, ([], [], []))
        assert np.array_equal(w["cell"][w["cell_mark"][b]:w["cell_mark"][b + 1]], _as_i32(rc, 2)), b
        assert np.array_equal(w["part"][w["part_mark"][b]:w["part_mark"][b + 1]], _as_i32(rp, 3)), b
        assert np.array_equal(w["soft"][w["soft_mark"][b]:w["soft_mark"][b + 1]], _as_i32(rs, 2)), b
    if soft:
        assert len(w["soft"]) > 0        # the softened split is exercised


def test_lists_cover_every_replica_mass_once(small_tree):
    t = small_tree
    w = t.walk(theta=0.7, n_replicas=1, period=1.0)
    total = t.parts[:, 0].sum()
    cmass = t.moments[:, 2]
    pm = np.concatenate([[0], np.cumsum(t.parts[:, 0])])
    for b in range(0, t.num_buckets, 5):
        c = w["cell"][w["cell_mark"][b]:w["cell_mark"][b + 1]]
        p = w["part"][w["part_mark"][b]:w["part_mark"][b + 1]]
        s = w["soft"][w["soft_mark"][b]:w["soft_mark"][b + 1]]
        m = cmass[c[:, 0]].sum() + cmass[s[:, 0]].sum() + (pm[p[:, 0] + p[:, 2]] - pm[p[:, 0]]).sum()
        assert np.isclose(m, 27 * total, rtol=1e-12)


def test_serialized_bucket_range_matches_full_walk(small_tree):
    """a rank's share of the buckets gets exactly the lists the full walk gives those buckets"""
    t = small_tree
    full = t.walk(theta=0.7, n_replicas=1, period=1.0)
    nb = t.num_buckets
    b0, b1 = nb // 3, 2 * nb // 3
    part = t.walk(theta=0.7, n_replicas=1, period=1.0, bucket_range=(b0, b1))
    for key, w in (("cell", 2), ("part", 3)):
        lo, hi = full[key + "_mark"][b0], full[key + "_mark"][b1]
        got = part[key]
        want = full[key][lo:hi]
        assert got.shape == want.shape
        # the low 22 bits name the walk's target bucket at decision time, which depends on where
        # the walk started; the replica code and the node/particle index must agree
        assert np.array_equal(got[:, 0], want[:, 0])
        assert np.array_equal(got[:, 1] & (0x1ff << 22), want[:, 1] & (0x1ff << 22))
        if w == 3:
            assert np.array_equal(got[:, 2], want[:, 2])
    assert part["cell_mark"][b0] == 0 and part["cell_mark"][-1] == len(part["cell"])


def test_expand_part_list(small_tree):
    t = small_tree
    w = t.walk(theta=0.7, n_replicas=0)
    ex, em = t.expand_part_list(w["part"], w["part_mark"])
    p = w["part"]
    want_idx = np.concatenate([np.arange(s, s + n) for s, n in zip(p[:, 0], p[:, 2])])
    assert np.array_equal(ex[:, 0], want_idx)
    assert np.array_equal(ex[:, 1], np.repeat(p[:, 1], p[:, 2]))
    assert em[-1] == len(ex) == p[:, 2].sum()


def test_tree_force_close_to_direct_sum():
    """lists + oracle evaluation reproduce the direct softened sum to tree-code accuracy
    (theta = 0.7, hexadecapole: ~1e-4 typical)"""
    pos, mass, soft = particles(2500, 21, clustered=True, soft=0.004)
    wl = tree_workload(pos, mass, soft, theta=0.7, n_replicas=0, max_bucket=12)
    parts = wl["parts"]
    v = np.zeros((len(parts), 5))
    orc.cell_list(parts, wl["moments"], *wl["cell"], 0.0, v)
    orc.part_list(parts, parts, *wl["part"], 0.0, v)
    if wl["softcell"]:
        orc.part_list(parts, np.ascontiguousarray(wl["softcell"][4]), *wl["softcell"][:4], 0.0, v)
    # direct sum through the same oracle: one bucket = all particles, list = all particles
    n = len(parts)
    il = np.column_stack([np.arange(n), np.full(n, (3 | 3 << 3 | 3 << 6) << 22)]).astype(np.int32)
    d = np.zeros((n, 5))
    orc.part_list(parts, parts, il, np.array([0, n], dtype=np.int32), np.array([0], dtype=np.int32),
                  np.array([n], dtype=np.int32), 0.0, d)
    rel = np.linalg.norm(v[:, :3] - d[:, :3], axis=1) / np.linalg.norm(d[:, :3], axis=1)
    assert np.median(rel) < 3e-4 and np.percentile(rel, 99) < 5e-3, (np.median(rel), rel.max())
    assert np.allclose(v[:, 3], d[:, 3], rtol=2e-3)


def test_ewald_tables_match_oracle(small_tree):
    root = small_tree.moments[0]
    momc, ewt = ewald_tables(root, 1.0, 2.8)
    momc_o, ewt_o = orc.ewald_tables(root, 1.0, 2.8)
    assert len(ewt) == len(ewt_o) == 80                       # dEwhCut = 2.8 -> NEWH (EwaldCUDA.h:6)
    np.testing.assert_allclose(momc, momc_o, rtol=1e-13, atol=1e-18)
    np.testing.assert_allclose(ewt, ewt_o, rtol=1e-12, atol=1e-20 + 1e-14 * np.abs(ewt_o).max())
    momc2, ewt2 = ewald_tables(root, 2.5, 2.0)
    momc2_o, ewt2_o = orc.ewald_tables(root, 2.5, 2.0)
    np.testing.assert_allclose(ewt2, ewt2_o, rtol=1e-12, atol=1e-14 * np.abs(ewt2_o).max())


def test_uniform_box_follows_ppartt_recipe():
    pos, mass, soft = workloads.uniform_box(1000, seed=1)
    assert pos.shape == (1000, 3) and pos.min() >= -0.5 and pos.max() <= 0.5
    assert np.isclose(mass.sum(), 1.0) and np.isclose(soft[0], 1000 ** (-1 / 3) / 20)
    # glibc rand() with seed 1: first draw is 1804289383
    assert np.isclose(pos[0, 0], -0.5 + 1804289383 / 2147483647.0)


@pytest.mark.skipif(not os.path.exists("/root/reference/teststep/king_soft.bin"), reason="reference fixtures absent")
def test_read_tipsy_reference_fixtures():
    pos, mass, soft = workloads.read_tipsy("/root/reference/teststep/king_soft.bin")
    assert len(pos) == 36000 and np.isclose(soft[0], 0.20421657)
    pos, mass, soft = workloads.read_tipsy("/root/reference/testcosmo/cube300.tbin")
    assert len(pos) == 110592 and np.isclose(mass.sum(), 0.3, rtol=1e-4)


def test_config_workload_shards_buckets():
    wl0 = workloads.config_workload("cube300", n=16 ** 3, bucket_range_of=(0, 2))
    wl1 = workloads.config_workload("cube300", n=16 ** 3, bucket_range_of=(1, 2))
    assert wl0["bucket_range"][1] == wl1["bucket_range"][0]
    n0 = wl0["cell"][3].sum()
    n1 = wl1["cell"][3].sum()
    assert n0 + n1 == 16 ** 3 and abs(int(n0) - int(n1)) <= 24
    assert np.array_equal(wl0["parts"], wl1["parts"])
    assert len(wl0["ewald"]["active"]) == n0


def test_walk_decisions_by_the_reference_gravity_h():
    """the sequential walk with every opening decision taken by the REFERENCE's own openCriterionNode /
    openSoftening (gravity.h compiled unmodified, oracle/_ref/libgravity_ref.so) builds the same lists as with the
    oracle's restatements, and so does the product's host walk -- on a clustered periodic tree with softened cells
    (live: needs /root/reference; the golden vectors of tests/golden/gravity_kat.npz carry the pin elsewhere)"""
    G = orc.ref_gravity()
    if G is None:
        pytest.skip("oracle/_ref/libgravity_ref.so needs /root/reference")

    class ReferenceDecisions(SequentialWalk):
        calls = 0

        def open_criterion(self, node, my, shift):
            ReferenceDecisions.calls += 1
            return G.gref_open_criterion_node(self.mom[node], int(self.last[node] - self.first[node] + 1),
                                              np.ascontiguousarray(shift), self.mom[my], self.boxlo[my], self.boxhi[my],
                                              int(self.is_bucket(my)))

        def open_softening(self, node, my, shift):
            return G.gref_open_softening(self.mom[node], np.ascontiguousarray(shift), self.mom[my], self.boxlo[my],
                                         self.boxhi[my])

    pos, mass, s = particles(1500, 21, soft=0.02)
    t = Tree(pos, mass, s, max_bucket=8)
    theta = 0.7
    G.gref_set_theta(theta, theta ** 4)
    args = (t.child0, t.child1, t.first, t.last, t.boxlo, t.boxhi, t.moments, t.bucket_node, theta, 1, 1.0)
    ref = ReferenceDecisions(*args).run()
    ours = SequentialWalk(*args).run()
    assert ReferenceDecisions.calls > 50000
    assert ref == ours
    w = t.walk(theta=theta, n_replicas=1, period=1.0)
    nsoft = 0
    for b in range(t.num_buckets):
        rc, rp, rs = ref[b]
        assert np.array_equal(w["cell"][w["cell_mark"][b]:w["cell_mark"][b + 1]], _as_i32(rc, 2)), b
        assert np.array_equal(w["part"][w["part_mark"][b]:w["part_mark"][b + 1]], _as_i32(rp, 3)), b
        assert np.array_equal(w["soft"][w["soft_mark"][b]:w["soft_mark"][b + 1]], _as_i32(rs, 2)), b
        nsoft += len(rs)
    assert nsoft > 0
