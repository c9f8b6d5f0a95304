#!/usr/bin/env python
"""Kernel-level numbers of BASELINE.json configs 1, 2 and 5 (king_soft.bin, cube300.tbin,
adiabtophat collapse): the reference's own particle sets, lists from the host walk, everything resident
in HBM, the list kernels and the Ewald kernel timed with CUDA events around each launch
(changa_b200.resident.ResidentStep; the device-pointer entry points of the C ABI).  One JSON line.

  python tools/resident_probe.py --workload cube300 [--double] [--steps 50]
  CB200_PC64_VARIANT=4 python tools/resident_probe.py --workload collapse --double"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cube300", choices=["cube300", "king", "collapse"])
    ap.add_argument("--double", action="store_true")
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    import torch
    from changa_b200.hostcuda import HostCUDA, ForceStep
    from changa_b200.resident import ResidentStep
    from changa_b200.workloads import config_workload, interaction_counts
    hc = HostCUDA(double=a.double, device=0)
    wl = config_workload(a.workload)
    cnt = interaction_counts(wl)
    rs = ResidentStep(hc, wl, torch)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(rs.ext):
        for _ in range(5):
            rs.step()
        torch.cuda.synchronize()
        evs = []
        for _ in range(a.steps):
            flush.zero_()  # evict lists / moments / particles from the 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(rs.ext)
            rs.step()
            e1.record(rs.ext)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(x.elapsed_time(y) for x, y in evs) / a.steps
        hc.timing(True)
        for _ in range(a.steps):
            flush.zero_()
            rs.step()
        torch.cuda.synchronize()
        taps = hc.timing_read()
        hc.timing(False)
    fs = ForceStep(hc, wl)
    for _ in range(3):
        fs.run()
    import time
    t0 = time.perf_counter()
    for _ in range(10):
        fs.run()
    e2e = (time.perf_counter() - t0) / 10
    fs.free()
    pc_ms = taps["cell_ms"] / max(taps["cell_launches"], 1)
    pp_ms = taps["part_ms"] / a.steps
    ew_ms = taps["ewald_ms"] / max(taps["ewald_launches"], 1)
    peak = json.load(open(os.path.join(ROOT, "profiles", "r01_fp32_peak.json")))
    peak_tf = float(peak["dfma_tflops" if a.double else "ffma_tflops"])
    pairs = cnt["cell"] + cnt["part"] + cnt["softcell"]
    pc_tf = cnt["cell"] * 198.0 / (pc_ms * 1e-3) / 1e12 if pc_ms else 0.0
    print(json.dumps({
        "workload": wl["name"], "dtype": "f64" if a.double else "f32", "pc_variant": os.environ.get("CB200_PC64_VARIANT"),
        "particles": len(wl["parts"]), "pc_pairs": cnt["cell"], "pp_pairs": cnt["part"] + cnt["softcell"],
        "resident_step_ms": ms, "interactions_per_s": pairs / (ms * 1e-3),
        "e2e_ms_reference_facing_abi": e2e * 1e3, "e2e_interactions_per_s": pairs / e2e,
        "pc_ms": pc_ms, "pp_ms": pp_ms, "ewald_ms": ew_ms,
        "pc_tflops": pc_tf, "pc_frac_of_fma_peak": pc_tf / peak_tf, "peak_tflops": peak_tf,
        "pp_pairs_per_s": (cnt["part"] + cnt["softcell"]) / (pp_ms * 1e-3) if pp_ms else None,
        "ewald_particles_per_s": (len(wl["ewald"]["active"]) if wl.get("ewald") and wl["ewald"]["active"] is not None
                                  else len(wl["parts"])) / (ew_ms * 1e-3) if ew_ms else None,
        "l2": "flushed between steps (256 MiB device write)"}), flush=True)


if __name__ == "__main__":
    main()
