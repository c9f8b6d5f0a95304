"""Generates tests/golden/gravity_kat.npz from the REFERENCE's own gravity.h
(oracle/_ref/libgravity_ref.so = /root/reference/gravity.h compiled unmodified through oracle/gravity_ref.cpp, with
the reference's moments.c linked in; see oracle/Makefile).  Run here, in the container that has /root/reference:

    python tests/golden/make_gravity_golden.py

Known answers of SPLINE, partBucketForce, nodeBucketForce (hexadecapole and softened branch), openSoftening,
openCriterionNode and openCriterionBucket, of the reference's own Ewald.cpp (EwaldInit, BucketEwald; same library,
oracle/ewald_ref.cpp) and of the moment build by the reference's own MultipoleMoments.h: the fixture keeps the oracle pinned to them on machines without the reference (the GPU box).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

THETA = 0.7
GEOM = 2.0 / np.sqrt(3.0)
HOME = 0xDB << 22


def f32(a):
    """inputs of the force cases are float-representable, so that the float CUDA path can be handed exactly the
    numbers the reference's double code saw (tests/test_gpu_parity.py)"""
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def random_cell(L, rng, centre, size, soft):
    """27-double cell record (gravity_oracle.c CM_* order) of a random cloud: centre of mass, radius, hexadecapole
    moments about the centre of mass scaled by the radius"""
    k = 24
    pos = centre + rng.uniform(-size / 2, size / 2, (k, 3))
    m = rng.uniform(0.5, 1.5, k)
    cm = (pos * m[:, None]).sum(0) / m.sum()
    rad = float(np.sqrt(((pos - cm) ** 2).sum(1)).max())
    fm = np.zeros(orc.FM_N)
    for i in range(k):
        t = np.zeros(orc.FM_N)
        d = pos[i] - cm
        L.orc_fm_make(t, m[i], rad, d[0], d[1], d[2])
        L.orc_fm_add(fm, t)
    c = np.zeros(orc.CM_N)
    c[0], c[1], c[2] = rad, soft, m.sum()
    c[3:6] = cm
    # FM order: m xx yy xy xz yz | xxx xyy xxy yyy xxz yyz xyz | xxxx ...;  CM order: xx xy xz yy yz | same | same
    c[6], c[7], c[8], c[9], c[10] = fm[1], fm[3], fm[4], fm[2], fm[5]
    c[11:27] = fm[6:22]
    return f32(c)


def bucket(rng, n, spread):
    part = np.ascontiguousarray(f32(np.column_stack([rng.uniform(0.5, 2, n), 10 ** rng.uniform(-3, -1.3, n),
                                                     rng.uniform(-spread, spread, (n, 3))])))
    lo, hi = part[:, 2:].min(0), part[:, 2:].max(0)
    my = np.zeros(orc.CM_N)
    my[2] = part[:, 0].sum()
    my[3:6] = (part[:, 2:] * part[:, :1]).sum(0) / my[2]
    my[1] = (part[:, 1] * part[:, 0]).sum() / my[2]
    my[0] = float(np.sqrt(((part[:, 2:] - my[3:6]) ** 2).sum(1)).max())
    return part, my, lo, hi


def criterion_case(rng, near):
    """(node6 = radius, soft, mass, cm; npart; shift; my soft + cm; box; isBucket).  near: the opening sphere is
    tuned to graze the box (or to graze containing it), where a different rounding would flip the answer."""
    node = np.zeros(orc.CM_N)
    my = np.zeros(orc.CM_N)
    node[0] = 10 ** rng.uniform(-2, -0.3)
    node[1] = 10 ** rng.uniform(-4, -1)
    node[2] = 1.0
    node[3:6] = rng.uniform(-0.5, 0.5, 3)
    c = rng.uniform(-0.5, 0.5, 3)
    h = 10 ** rng.uniform(-3, -0.5, 3)
    lo, hi = c - h, c + h
    my[3:6] = c + rng.uniform(-1, 1, 3) * h
    my[1] = 10 ** rng.uniform(-4, -1)
    my[2] = 1.0
    shift = rng.integers(-1, 2, 3).astype(float) if rng.random() < 0.6 else np.zeros(3)
    if near:
        s = node[3:6] + shift
        gap = np.maximum(np.maximum(lo - s, s - hi), 0.0)
        far = np.maximum(np.abs(lo - s), np.abs(hi - s))
        d = np.sqrt((gap ** 2).sum()) if near == 1 else np.sqrt((far ** 2).sum())
        if d > 0:
            ropen = d * (1.0 + rng.uniform(-1e-9, 1e-9) * rng.integers(0, 2))
            node[0] = ropen / max(GEOM / THETA, 1.0)
    return node, int(rng.integers(1, 40)), shift, my, lo, hi, int(rng.integers(0, 2))


def main():
    orc.build()
    L, G = orc.lib(), orc.ref_gravity()
    assert G is not None, "needs /root/reference (oracle/_ref/libgravity_ref.so)"
    G.gref_set_theta(THETA, THETA ** 4)
    rng = np.random.default_rng(20261018)
    out = {"theta": np.array([THETA, THETA ** 4])}

    # SPLINE: inside h, between h and 2h, outside, and on the joints
    n = 3000
    twoh = 10 ** rng.uniform(-3, 0, n)
    r = twoh * 10 ** rng.uniform(-2, 0.5, n)
    r[:50] = twoh[:50]
    r[50:100] = 0.5 * twoh[50:100]
    ab = np.zeros((n, 2))
    for i in range(n):
        a, b = np.zeros(1), np.zeros(1)
        G.gref_spline(r[i] * r[i], twoh[i], a, b)
        ab[i] = a[0], b[0]
    out.update(spline_r2=r * r, spline_twoh=twoh, spline_ab=ab)

    # partBucketForce: one source on a 12-particle bucket, with replica shifts, a coincident pair (skipped),
    # pairs inside the softening length, and accumulation on top of earlier values
    n = 96
    pb_part, pb_src, pb_shift, pb_in, pb_out = [], [], [], [], []
    for i in range(n):
        part, _, _, _ = bucket(rng, 12, 0.05)
        src = np.array([rng.uniform(0.5, 2), 10 ** rng.uniform(-3, -1.3), *rng.uniform(-0.2, 0.2, 3)])
        if i % 4 == 1:
            src[2:] = part[5, 2:] + rng.uniform(-1, 1, 3) * 0.3 * src[1]
        src = f32(src)
        if i % 4 == 2:
            src[:] = part[7]
        shift = rng.integers(-1, 2, 3).astype(float) if i % 3 == 0 else np.zeros(3)
        v = rng.normal(size=(12, 5)) if i % 5 == 0 else np.zeros((12, 5))
        v[:, 4] = np.abs(v[:, 4])
        pb_in.append(v.copy())
        G.gref_part_bucket_force(src, shift, part, 0, 11, None, 0, v)
        pb_part.append(part); pb_src.append(src); pb_shift.append(shift); pb_out.append(v)
    out.update(pb_part=np.array(pb_part), pb_src=np.array(pb_src), pb_shift=np.array(pb_shift), pb_in=np.array(pb_in),
               pb_out=np.array(pb_out))

    # nodeBucketForce: a hexadecapole cell on a bucket, from touching distance to far away; some cells soft enough
    # to take the particle branch
    n = 384
    nb_cell, nb_shift, nb_my, nb_lo, nb_hi, nb_part, nb_out, nb_soft = [], [], [], [], [], [], [], []
    for i in range(n):
        part, my, lo, hi = bucket(rng, 12, 0.05)
        dist = 10 ** rng.uniform(-1.2, 0.5)
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        shift = rng.integers(-1, 2, 3).astype(float) if i % 3 == 0 else np.zeros(3)
        cell = random_cell(L, rng, d * dist - shift, 0.05, 10 ** rng.uniform(-3, -0.8))
        v = np.zeros((12, 5))
        nb_soft.append(G.gref_open_softening(cell, shift, my, lo, hi))
        G.gref_node_bucket_force(cell, shift, my, lo, hi, part, 0, 11, None, 0, v)
        nb_cell.append(cell); nb_shift.append(shift); nb_my.append(my); nb_lo.append(lo); nb_hi.append(hi)
        nb_part.append(part); nb_out.append(v)
    out.update(nb_cell=np.array(nb_cell), nb_shift=np.array(nb_shift), nb_my=np.array(nb_my), nb_lo=np.array(nb_lo),
               nb_hi=np.array(nb_hi), nb_part=np.array(nb_part), nb_out=np.array(nb_out), nb_soft=np.array(nb_soft, dtype=np.int32))

    # opening criteria: random cases, and cases tuned to graze
    n = 6000
    oc_node, oc_np, oc_shift, oc_my, oc_lo, oc_hi, oc_isb, oc_node_out, oc_bucket_out, oc_soft_out = ([] for _ in range(10))
    for i in range(n):
        node, npart, shift, my, lo, hi, isb = criterion_case(rng, near=(0 if i % 2 == 0 else (1 if i % 4 == 1 else 2)))
        oc_node.append(node[:6]); oc_np.append(npart); oc_shift.append(shift); oc_my.append(my[:6]); oc_lo.append(lo)
        oc_hi.append(hi); oc_isb.append(isb)
        oc_node_out.append(G.gref_open_criterion_node(node, npart, shift, my, lo, hi, isb))
        oc_bucket_out.append(G.gref_open_criterion_bucket(node, npart, shift, my, lo, hi))
        oc_soft_out.append(G.gref_open_softening(node, shift, my, lo, hi))
    i32 = lambda a: np.array(a, dtype=np.int32)
    out.update(oc_node=np.array(oc_node), oc_npart=i32(oc_np), oc_shift=np.array(oc_shift), oc_my=np.array(oc_my),
               oc_lo=np.array(oc_lo), oc_hi=np.array(oc_hi), oc_isb=i32(oc_isb), oc_node_out=i32(oc_node_out),
               oc_bucket_out=i32(oc_bucket_out), oc_soft_out=i32(oc_soft_out))

    # Ewald.cpp: EwaldInit (complete root moments, h-loop table) and BucketEwald on particles all over the box,
    # next to the root's centre of mass (the small-r expansion, Ewald.cpp:141-152) and in every replica setting
    ew_root, ew_L, ew_hcut, ew_cut, ew_nrep, ew_momc, ew_ewt, ew_newt, ew_part, ew_out = ([] for _ in range(10))
    for i, (Lbox, hcut, fcut, nrep) in enumerate([(1.0, 2.8, 2.6, 1), (1.0, 2.8, 2.6, 0), (1.0, 1.5, 1.7, 2), (2.5, 2.8, 2.6, 1),
                                                   (0.4, 3.2, 3.1, 1), (1.0, 2.8, 2.6, 1)]):
        root = random_cell(L, rng, rng.uniform(-0.05, 0.05, 3) * Lbox, Lbox, 0.01 * Lbox)
        momc = np.zeros(orc.MC_N)
        ewt = np.zeros((256, 5))
        nh = G.eref_init(root, Lbox, hcut, momc, ewt.reshape(-1), 256)
        assert nh <= 256
        n = 48
        part = np.ascontiguousarray(np.column_stack([np.full(n, 1.0 / n), np.full(n, 0.01 * Lbox),
                                                     rng.uniform(-0.5, 0.5, (n, 3)) * Lbox]))
        part[0, 2:] = root[3:6] + 1e-3 * Lbox * rng.normal(size=3)
        part[1, 2:] = root[3:6] + np.array([0.03, 0.0, 0.0]) * Lbox
        part[2, 2:] = root[3:6] + np.array([0.0, 0.035, 0.0]) * Lbox
        v = np.zeros((n, 5))
        G.eref_bucket_ewald(root, Lbox, hcut, fcut, nrep, part, 0, n - 1, None, 0, v)
        ew_root.append(root); ew_L.append(Lbox); ew_hcut.append(hcut); ew_cut.append(fcut); ew_nrep.append(nrep)
        ew_momc.append(momc); ew_ewt.append(ewt); ew_newt.append(nh); ew_part.append(part); ew_out.append(v)
    out.update(ew_root=np.array(ew_root), ew_L=np.array(ew_L), ew_hcut=np.array(ew_hcut), ew_cut=np.array(ew_cut),
               ew_nrep=i32(ew_nrep), ew_momc=np.array(ew_momc), ew_ewt=np.array(ew_ewt), ew_newt=i32(ew_newt),
               ew_part=np.array(ew_part), ew_out=np.array(ew_out))

    # MultipoleMoments.h: the moment build of two small trees by the reference's own class (one with 40 coincident
    # particles: zero-size boxes take the first-particle radius rule, GenericTreeNode.h:226-233).  The topology comes
    # from the product's host tree; the known answer is what the reference's operations make of it.
    from changa_b200.tree import Tree
    for tag, n, mb, dup in (("a", 600, 12, False), ("b", 300, 8, True)):
        pos = rng.uniform(-0.5, 0.5, (n, 3))
        if dup:
            pos[100:140] = pos[99]
        t = Tree(pos, rng.uniform(0.5, 1.5, n) / n, rng.uniform(0.001, 0.01, n), max_bucket=mb)
        nn = len(t.child0)
        mom = np.zeros((nn, orc.CM_N))
        f64, ii = orc.as_f64, orc.as_i32
        G.gref_build_moments(f64(t.parts[:, 2:5]).reshape(-1), f64(t.parts[:, 0]), f64(t.parts[:, 1]), ii(t.child0), ii(t.child1),
                             ii(t.first), ii(t.last), f64(t.geolo).reshape(-1), f64(t.geohi).reshape(-1),
                             f64(t.boxlo).reshape(-1), f64(t.boxhi).reshape(-1), nn, mom.reshape(-1))
        out.update({f"mb_{tag}_parts": t.parts.copy(), f"mb_{tag}_child0": t.child0.copy(), f"mb_{tag}_child1": t.child1.copy(),
                    f"mb_{tag}_first": t.first.copy(), f"mb_{tag}_last": t.last.copy(), f"mb_{tag}_geolo": t.geolo.copy(),
                    f"mb_{tag}_geohi": t.geohi.copy(), f"mb_{tag}_boxlo": t.boxlo.copy(), f"mb_{tag}_boxhi": t.boxhi.copy(),
                    f"mb_{tag}_moments": mom})

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gravity_kat.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", "criterion outcomes (-1, 0, 1):",
          [int((out["oc_node_out"] == k).sum()) for k in (-1, 0, 1)], "softened cells:", int(out["nb_soft"].sum()))


if __name__ == "__main__":
    main()
