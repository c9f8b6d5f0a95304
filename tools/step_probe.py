#!/usr/bin/env python
"""Time the in-library force step (cb200_step_run) on one box: one JSON line with the phases, the kernel
times, the pair counts and interactions/s, resident (records already in HBM) and end to end (pinned host
records in, rows out).  Under torchrun every rank joins one NCCL communicator and the box is shared.

  python tools/step_probe.py --n 4194304 [--kind clustered] [--steps 5] [--no-overlap] [--no-cost-cuts]
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/step_probe.py --n ..."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--particles", dest="n", type=int, default=1 << 22)
    ap.add_argument("--kind", default="uniform", choices=["uniform", "clustered"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-cost-cuts", action="store_true")
    ap.add_argument("--active-rung", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    from changa_b200.hostcuda import HostCUDA
    from changa_b200.step import Comm, NativeStep
    from changa_b200.workloads import uniform_box, clustered_box, density_rungs
    hc = HostCUDA(double=False, device=local)
    comm = Comm.from_env(hc.L)
    pos, mass, soft = uniform_box(a.n, seed=1) if a.kind == "uniform" else clustered_box(a.n, seed=2)
    rung = density_rungs(pos) if a.active_rung > 0 else None
    st = NativeStep(hc, a.n, theta=0.7, n_replicas=1, period=1.0, ewald={"dEwCut": 2.6, "dEwhCut": 2.8}, comm=comm,
                    active_rung=a.active_rung, overlap_ewald=not a.no_overlap, cost_cuts=not a.no_cost_cuts)
    st.set_particles(pos, float(mass[0]), float(soft[0]), rung)
    del pos
    for _ in range(3):
        res = st.run()
    out = {"n": a.n, "kind": a.kind, "world": world, "rank": rank, "nodes": res.numNodes, "buckets": res.numBuckets}
    for mode in ("e2e", "resident"):
        if mode == "resident":
            st.upload()
            st.run(resident=True)
        ph = {}
        hc.timing(True)
        comm.barrier(st.stream)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            res = st.run(resident=(mode == "resident"))
            for k, v in st.phases().items():
                ph[k] = ph.get(k, 0.0) + v / a.steps
        wall = (time.perf_counter() - t0) / a.steps
        taps = hc.timing_read()
        hc.timing(False)
        agg = comm.allreduce([wall], "max")
        tot = comm.allreduce([res.pcPairs, res.ppPairs, res.rows], "sum")
        out[mode] = {"ms_per_step": float(agg[0]) * 1e3, "interactions_per_s": float(tot[0] + tot[1]) / float(agg[0]),
                     "rank_phases_ms": {k: round(v, 3) for k, v in ph.items()},
                     "rank_pc_ms": taps["cell_ms"] / a.steps, "rank_pp_ms": taps["part_ms"] / a.steps,
                     "rank_ewald_ms": taps["ewald_ms"] / a.steps,
                     "rank_pc_tflops": res.pcPairs * 198.0 / max(taps["cell_ms"] / a.steps, 1e-9) / 1e9,
                     "rank_pp_pairs_per_s": res.ppPairs / max(taps["part_ms"] / a.steps, 1e-9) * 1e3}
        out["pc_pairs"], out["pp_pairs"], out["rows"] = float(tot[0]), float(tot[1]), float(tot[2])
        out["rank_range"] = [res.bucketLo, res.bucketHi, res.partLo, res.partHi]
    st.free()
    comm.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.destroy()


if __name__ == "__main__":
    main()
