#!/usr/bin/env python
"""Force step with the lists built on the GPU (DeviceTreeStep) on a large box:
  python tools/bench_device_tree.py --workload uniform --n 4194304 [--steps 3]
prints one JSON line: phase times, pair counts, rates."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="uniform", choices=["uniform", "clustered", "cosmo"])
    ap.add_argument("--n", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--theta", type=float, default=0.7)
    a = ap.parse_args()
    import torch
    from changa_b200.hostcuda import HostCUDA
    from changa_b200.device_step import DeviceTreeStep
    from changa_b200.tree import Tree
    from changa_b200 import workloads as W
    t0 = time.time()
    if a.workload == "uniform":
        pos, mass, soft = W.uniform_box(a.n, seed=1)
    elif a.workload == "clustered":
        pos, mass, soft = W.clustered_box(a.n, seed=2)
    else:
        pos, mass, soft = W.cosmo_box(int(round(a.n ** (1 / 3))))
    t1 = time.time()
    tree = Tree(pos, mass, soft, max_bucket=12)
    t2 = time.time()
    hc = HostCUDA(double=False, device=0)
    st = DeviceTreeStep(hc, tree, theta=a.theta, n_replicas=1, period=1.0, ewald={"dEwCut": 2.6, "dEwhCut": 2.8})
    st.run()  # warm-up (pool growth)
    hc.timing(True)
    phases = {}
    w0 = time.perf_counter()
    for _ in range(a.steps):
        out = st.run(phases=phases)
    wall = (time.perf_counter() - w0) / a.steps
    taps = hc.timing_read()
    info = st.lists_info
    bs = tree.bucket_sizes.astype(np.int64)
    # pair counts need the markers: keep them from one more run
    st.run(keep_lists=True)
    k = st.kept
    pc = int((np.diff(k["cell_mark"].astype(np.int64)) * bs).sum())
    pp = int((np.diff(k["part_mark"].astype(np.int64)) * bs).sum()) + int((np.diff(k["soft_mark"].astype(np.int64)) * bs).sum())
    ph = {n: v / a.steps for n, v in phases.items()}
    pc_ms = taps["cell_ms"] / max(taps["cell_launches"], 1)
    line = {"workload": a.workload, "n": tree.n, "buckets": tree.num_buckets, "nodes": tree.num_nodes,
            "gen_s": round(t1 - t0, 2), "host_tree_s": round(t2 - t1, 2), "wall_ms_per_step": wall * 1e3,
            "phases_ms": {n: round(v, 3) for n, v in ph.items()}, "lists": info, "pc_pairs": pc, "pp_pairs": pp,
            "pc_ms": pc_ms, "pp_ms": taps["part_ms"] / a.steps, "ewald_ms": taps["ewald_ms"] / max(taps["ewald_launches"], 1),
            "pc_pairs_per_s": pc / (pc_ms * 1e-3), "pc_tflops_198": pc * 198 / (pc_ms * 1e-3) / 1e12,
            "interactions_per_s_wall": (pc + pp) / wall, "h2d_bytes": st.h2d_bytes, "d2h_bytes": st.d2h_bytes,
            "acc_finite": bool(np.isfinite(out).all())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
