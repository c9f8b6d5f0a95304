"""cb200_step_run on tiny and degenerate particle sets (open boundary, no Ewald), each case in its own process
with a timeout: n = 1, 2, 12 (one bucket), 13, 100, 1000 against a double direct sum; 50 particles at one point.
usage: python tools/edge_probe.py [case ...]   (no argument: every case, one JSON line each; run it under `timeout`)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = ["1000", "100", "13", "12", "pair_in_soft", "2", "same50", "1"]


def direct(pos, mass, soft):
    """double direct sum with the spline of gravity.h:147-182 (through the oracle's pair routine)"""
    from oracle import oracle as orc
    n = len(pos)
    parts = np.ascontiguousarray(np.column_stack([np.full(n, mass), np.full(n, soft), pos]).astype(np.float32).astype(np.float64))
    il = np.zeros((n, 2), dtype=np.int32)
    il[:, 0] = np.arange(n)
    il[:, 1] = 0xDB << 22
    v = np.zeros((n, 5))
    # one "bucket" per particle, each with the whole set as its list
    lists = np.ascontiguousarray(np.tile(il, (n, 1)))
    marks = (np.arange(n + 1) * n).astype(np.int32)
    starts = np.arange(n, dtype=np.int32)
    sizes = np.ones(n, dtype=np.int32)
    orc.part_list(parts, parts, lists, marks, starts, sizes, 0.0, v)
    return v


_HC = []


def host():
    if not _HC:
        from changa_b200.hostcuda import HostCUDA
        _HC.append(HostCUDA(double=False, device=0))
    return _HC[0]


def run_case(case):
    from changa_b200.step import NativeStep
    hc = host()
    rng = np.random.default_rng(11)
    soft = 1e-4
    if case == "same50":
        n = 50
        pos = np.tile(np.array([[0.1, -0.2, 0.3]]), (n, 1))
    elif case == "pair_in_soft":
        n = 40
        pos = rng.uniform(-0.45, 0.45, (n, 3))
        pos[1] = pos[0] + 0.3 * soft  # inside the softening length of each other
    else:
        n = int(case)
        pos = rng.uniform(-0.45, 0.45, (n, 3))
    mass = 1.0 / n
    st = NativeStep(hc, n, theta=0.7, n_replicas=0, period=1.0, ewald=None)
    try:
        st.set_particles(pos, mass, soft)
        res = st.run()
        got = st.out.array[:n].astype(np.float64).copy()
        out = {"case": case, "n": n, "nodes": res.numNodes, "buckets": res.numBuckets, "levels": res.numLevels,
               "pc_pairs": res.pcPairs, "pp_pairs": res.ppPairs}
    finally:
        st.free()
    want = direct(pos, mass, soft)
    amag = np.linalg.norm(want[:, :3], axis=1)
    da = np.linalg.norm(got[:, :3] - want[:, :3], axis=1)
    scale = max(float(np.sqrt((amag ** 2).mean())), 1e-300)
    out.update(finite=bool(np.isfinite(got).all()), max_da_over_rms_a=float(da.max() / scale) if amag.max() > 0 else float(da.max()),
               max_abs_a=float(np.abs(got[:, :3]).max()),
               max_dpot_over_pot=float((np.abs(got[:, 3] - want[:, 3]) / np.maximum(np.abs(want[:, 3]), 1e-300)).max())
               if np.abs(want[:, 3]).max() > 0 else float(np.abs(got[:, 3]).max()))
    print(json.dumps(out))


if __name__ == "__main__":
    for c in (sys.argv[1:] or CASES):
        run_case(c)
        sys.stdout.flush()
