#!/usr/bin/env python
"""bench.py -- one force step of ChaNGa's GPU gravity hot path on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cube300] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one force evaluation of one synthetic box: (all-gather of particle and
moment slices when N > 1) -> layout pack -> zero accumulators -> particle-cell lists
(hexadecapole) -> particle-particle lists -> softened cells -> Ewald.  Default workload
= BASELINE.json configs[1] (testcosmo cube300.tbin: 48^3 periodic box, theta 0.7, nReplicas 1,
Ewald) on the reference's own particle set (positions committed under tests/golden/).  At N > 1
GPUs the box holds N x 48^3 particles of a synthetic stand-in (a Zel'dovich-displaced grid) and
every rank owns a contiguous SFC range of buckets (weak scaling).

  value   pair interactions (p-c + p-p, counted like Compute.cpp:1643-1651) per second of
          the whole step with raw inputs and lists already in HBM, CUDA-event timed on
          the launching stream, L2 flushed between steps, max over ranks
  e2e     same metric through the reference-facing C ABI (DataManagerTransferLocalTree,
          TreePiece*ListDataTransferLocal, EwaldHost, TransferParticleVarsBack) from
          pinned HOST buffers: H2D + kernels + D2H inside the timed region, wall clock
  roofline  the p-c kernel against the measured FP32 FFMA peak (198 flop/pair convention,
          SURVEY.md 8d); also the HBM side of list streaming
  cpu_baseline  the oracle port (oracle/gravity_oracle.c, OpenMP) on the host cores

--impl reference times the CPU restatement of the reference's own gravity
(nodeBucketForce / partBucketForce / BucketEwald) with all host threads on the same
workload.  No part of the GPU arm runs through oracle/.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PC = 198.0      # per p-c pair (6 shift + 5 r^2 + 1 rsqrt + 181 eval + 5 idt2)
FLOP_PC_REF = 170.0  # the reference's own count (moments.c:1466)
FLOP_PP = 30.0       # per p-p pair, unsoftened branch
FLOP_EW_REAL, FLOP_EW_K = 350.0, 58.0


def measured_peaks():
    out = {"fp64_tflops": 36.98, "fp64_source": "fallback: DFMA run of round 1 (profiles/r01_fp32_peak.json)",
           "fp32_tflops": 71.64, "fp32_source": "fallback: tools/fp32_peak.cu FFMA run of round 1 (profiles/r01_fp32_peak.json)",
           "hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "profiles", "r01_fp32_peak.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            out["fp32_tflops"] = float(j["ffma_tflops"])
            out["fp32_source"] = "measured: tools/fp32_peak.cu scalar FFMA chains on this pool's B200 (profiles/r01_fp32_peak.json)"
            out["fp64_tflops"] = float(j["dfma_tflops"])
            out["fp64_source"] = "measured: tools/fp32_peak.cu DFMA chains on this pool's B200 (profiles/r01_fp32_peak.json)"
        except Exception:
            pass
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            out["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            out["hbm_source"] = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    # DRAM bytes of one p-c launch on the default workload, from the committed `ncu --set full`
    # capture of this same command (tools/ncu_summary.py); null when no capture is committed
    out["pc_traffic"], out["pc_traffic_source"] = None, None
    import glob
    caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_cell_list_x2.json")))
    if caps:
        try:
            out["pc_traffic"] = float(json.load(open(caps[-1]))["dram_bytes"])
            out["pc_traffic_source"] = os.path.relpath(caps[-1], ROOT)
        except Exception:
            pass
    return out


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[4:8]) if v.lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def ewald_real_terms(wl):
    """replicas that pass the real-space cut per active particle (EwaldKernel, HostCUDA.cu:2039-2060)"""
    ew = wl.get("ewald")
    if not ew:
        return 0, 0
    parts = wl["parts"]
    act = ew["active"] if ew["active"] is not None else np.arange(len(parts))
    d = parts[act][:, 2:5] - np.asarray(ew["root"][3:6])
    L, nE, nR = ew["L"], int(np.ceil(ew["fEwCut"])), ew["nReps"]
    cut2 = (ew["fEwCut"] * L) ** 2
    total = 0
    r = np.arange(-nE, nE + 1)
    for ix in r:
        for iy in r:
            for iz in r:
                hole = max(abs(ix), abs(iy), abs(iz)) <= nR
                if hole:
                    total += len(d)
                    continue
                q = d + np.array([ix, iy, iz]) * L
                total += int(((q ** 2).sum(1) <= cut2).sum())
    return total, len(act)


def cpu_force_step(wl, repeats=1):
    """the oracle port on the host cores: returns (seconds per step, threads)"""
    from oracle import oracle as orc
    orc.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))  # torchrun pins OMP_NUM_THREADS=1
    parts = np.ascontiguousarray(wl["parts"])
    mom = np.ascontiguousarray(wl["moments"])
    best = None
    for _ in range(repeats):
        v = np.zeros((len(parts), 5))
        t0 = time.perf_counter()
        orc.cell_list(parts, mom, *wl["cell"], wl["fperiod"], v)
        orc.part_list(parts, parts, *wl["part"], wl["fperiod"], v)
        if wl.get("softcell"):
            orc.part_list(parts, np.ascontiguousarray(wl["softcell"][4]), *wl["softcell"][:4], wl["fperiod"], v)
        ew = wl.get("ewald")
        if ew:
            orc.ewald(parts, ew["active"], ew["root"], ew["momc"], ew["L"], ew["fEwCut"], ew["nReps"],
                      int(np.ceil(ew["fEwCut"])), 1.2e-3, ew["ewt"], v)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, int(orc.lib().orc_num_threads())


class ResidentStep:
    """the force step with every input already in HBM (device pointers through the C ABI)"""

    def __init__(self, hc, wl, torch, dist, rank, world):
        self.hc, self.torch, self.dist, self.rank, self.world = hc, torch, dist, rank, world
        L = hc.L
        self.stream = hc.stream_create()
        self.ext = torch.cuda.ExternalStream(self.stream)
        f32 = hc.np_real  # float32, or float64 under --double (the CUDA_USE_DOUBLE build)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.n, self.nn = len(wl["parts"]), len(wl["moments"])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        with torch.cuda.stream(self.ext):
            parts = np.ascontiguousarray(wl["parts"], dtype=f32)
            mom = np.ascontiguousarray(wl["moments"], dtype=f32)
            pb, mb = L.cb200_packed_particle_bytes(), L.cb200_packed_moment_bytes()
            if world > 1:
                # every rank owns an equal, padded slice of both arrays.  It packs ITS slice into the
                # kernels' layout and the all-gather lands the slices directly in the replicated packed
                # arrays: no rank ever repacks the other ranks' records
                from changa_b200.multigpu import shard_rows
                mine_p, self.pc = shard_rows(parts, rank, world)
                mine_m, self.mc = shard_rows(mom, rank, world)
                self.my_parts, self.my_mom = up(mine_p), up(mine_m)
                self.send_p = torch.empty(self.pc * pb, dtype=torch.uint8, device=dev)
                self.send_m = torch.empty(self.mc * mb, dtype=torch.uint8, device=dev)
                self.npk, self.nmk = self.pc * world, self.mc * world
            else:
                self.raw_parts, self.raw_mom = up(parts), up(mom)
                self.npk, self.nmk = self.raw_parts.shape[0], self.raw_mom.shape[0]
            self.pk_parts = torch.empty(self.npk * pb, dtype=torch.uint8, device=dev)
            self.pk_mom = torch.empty(self.nmk * mb, dtype=torch.uint8, device=dev)
            self.vars = torch.zeros((self.n, 5), dtype=torch.float64 if f32 == np.float64 else torch.float32, device=dev)
            self.lists = {}
            for key in ("cell", "part", "softcell"):
                if wl.get(key) and len(wl[key][0]):
                    il, m, st, sz = wl[key][:4]
                    self.lists[key] = (up(il), up(m), up(st), up(sz), len(st), int(sz.max()))
            self.soft_src = None
            if "softcell" in self.lists:
                raw = up(np.ascontiguousarray(wl["softcell"][4], dtype=f32))
                self.soft_src = torch.empty(raw.shape[0] * L.cb200_packed_particle_bytes(), dtype=torch.uint8, device=dev)
                L.cb200_pack_particles_device(raw.data_ptr(), self.soft_src.data_ptr(), raw.shape[0], self.stream)
            self.ew = None
            ew = wl.get("ewald")
            if ew:
                act = ew["active"] if ew["active"] is not None else np.arange(self.n, dtype=np.int32)
                self.ew_markers = up(np.ascontiguousarray(act, dtype=np.int32))
                e = hc.EwaldHostMemorySetup(len(act), len(ew["ewt"]), 1)
                hc.fill_ewald(e, ew["root"], ew["momc"], ew["ewt"], ew["L"], ew["fEwCut"], ew["nReps"], active=act)
                self.ew, self.ew_n = e, len(act)
            self.fperiod = float(wl.get("fperiod", 0.0))
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()

    def step(self):
        L, s = self.hc.L, self.stream
        pending = None
        if self.world > 1:
            # particles first: Ewald needs nothing else, so the (larger) moment all-gather runs on
            # NCCL's stream under the Ewald kernel and is waited for only in front of the p-c lists
            L.cb200_pack_particles_device(self.my_parts.data_ptr(), self.send_p.data_ptr(), self.pc, s)
            L.cb200_pack_moments_device(self.my_mom.data_ptr(), self.send_m.data_ptr(), self.mc, s)
            self.dist.all_gather_into_tensor(self.pk_parts, self.send_p)
            pending = self.dist.all_gather_into_tensor(self.pk_mom, self.send_m, async_op=True)
        else:
            L.cb200_pack_moments_device(self.raw_mom.data_ptr(), self.pk_mom.data_ptr(), self.nmk, s)
            L.cb200_pack_particles_device(self.raw_parts.data_ptr(), self.pk_parts.data_ptr(), self.npk, s)
        L.cb200_zero_vars_device(self.vars.data_ptr(), self.n, s)
        P, V, M = self.pk_parts.data_ptr(), self.vars.data_ptr(), self.pk_mom.data_ptr()
        if self.ew is not None:  # same order as ForceStep.run: Ewald needs only the particles
            L.cb200_ewald_device(P, V, self.ew_markers.data_ptr(), self.ew_n, self.ew.cachedData, self.ew.ewt, s)
        if pending is not None:
            pending.wait()  # stream-level: the step's stream waits for the moment all-gather
        if "cell" in self.lists:
            il, m, st, sz, nb, mx = self.lists["cell"]
            L.cb200_cell_list_device_ex(P, V, M, il.data_ptr(), m.data_ptr(), st.data_ptr(), sz.data_ptr(), nb,
                                        self.fperiod, mx, s)
        if "part" in self.lists:
            il, m, st, sz, nb, mx = self.lists["part"]
            L.cb200_part_list_device_ex(P, V, P, il.data_ptr(), m.data_ptr(), st.data_ptr(), sz.data_ptr(), nb,
                                        self.fperiod, mx, s)
        if "softcell" in self.lists:
            il, m, st, sz, nb, mx = self.lists["softcell"]
            L.cb200_part_list_device_ex(P, V, self.soft_src.data_ptr(), il.data_ptr(), m.data_ptr(), st.data_ptr(),
                                        sz.data_ptr(), nb, self.fperiod, mx, s)

    def capture(self):
        """one step as a CUDA graph (single GPU: the step is kernels and one memset on one stream,
        all operands resident): replaying it removes the host's launch cost and most of the gap
        between the step's short kernels.  Returns the kernels the graph holds."""
        torch = self.torch
        before = self.hc.kernel_launches()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.ext):
            self.step()
        return self.hc.kernel_launches() - before

    def timed(self, steps, warmup, graph=False):
        torch = self.torch
        run = self.graph.replay if graph else self.step
        with torch.cuda.stream(self.ext):
            for _ in range(warmup):
                self.flush.zero_()
                run()
            torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier()
            torch.cuda.synchronize()
            evs = []
            for _ in range(steps):
                self.flush.zero_()  # evict the lists / moments / particles from the 126 MB L2
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(self.ext)
                run()
                b.record(self.ext)
                evs.append((a, b))
            torch.cuda.synchronize()
            if self.world > 1:
                self.dist.barrier()
            torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)  # ms over exactly `steps` steps


def large_box_step(hc, n, torch, dist, rank, world, steps=3, kind="uniform", active_rung=0):
    """The same force step on a box that fills a GPU (uniform, SURVEY 8d recipe C3 at a size one
    default run can afford), from UNSORTED host particles: upload 40 B/particle, then keys, sort,
    tree, moments, double walk, p-c / p-p / Ewald all on the device, accelerations back in the
    caller's order (changa_b200.device_step.RawParticleStep).  At N > 1 the SAME box is shared
    (strong scaling): every rank uploads 1/N of the records, one NCCL all-gather replicates them,
    tree and moments are built redundantly, each rank walks and evaluates its own SFC range of
    buckets.  Reported next to the headline numbers: kernel rates without the launch ramp and
    tail of the 110k-particle box."""
    from changa_b200.device_step import RawParticleStep
    from changa_b200.workloads import uniform_box, clustered_box
    pos, mass, soft = uniform_box(n, seed=1) if kind == "uniform" else clustered_box(n, seed=2)
    mass, soft = float(mass[0]), float(soft[0])  # equal-mass box: scalars (1/N, N^(-1/3)/20)
    rung = None
    if active_rung > 0:  # multistep force step: rungs by local density
        from changa_b200.workloads import density_rungs
        rung = density_rungs(pos)
    st = RawParticleStep(hc, pos, mass, soft, theta=0.7, n_replicas=1, period=1.0,
                         ewald={"dEwCut": 2.6, "dEwhCut": 2.8}, max_bucket=12, dist=dist, rank=rank, world=world,
                         rung=rung, active_rung=active_rung)
    del pos
    st.run(count_pairs=True)  # warm-up (pool growth); the markers give the pair counts
    info = dict(st.info)
    hc.timing(True)
    phases = {}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = st.run(phases=phases)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    rows = out[1] if world > 1 else out
    finite = bool(np.isfinite(rows).all())
    taps = hc.timing_read()
    hc.timing(False)
    h2d, d2h = st.h2d_bytes, st.d2h_bytes
    free_b, total_b = torch.cuda.mem_get_info()
    st_active = {"active_" + k: v for k, v in getattr(st, "active", {}).items()}
    st.free()
    agg = torch.tensor([wall], dtype=torch.float64, device="cuda")
    tot = torch.tensor([info["pc_pairs"], info["pp_pairs"], h2d, d2h, len(rows)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    wall = float(agg[0])
    pc, pp, h2d, d2h, nrows = (float(x) for x in tot.tolist())
    pc_ms = taps["cell_ms"] / max(taps["cell_launches"], 1)
    multistep = {"active_rung": active_rung, **st_active} if active_rung > 0 else None
    return {"workload": f"{kind}(N={n},theta=0.7,nReplicas=1,bucket=12), tree and lists built on the device",
            "multistep": multistep,
            "scaling": "strong" if world > 1 else None, "n_gpus": world,
            "ms_per_step": wall * 1e3, "steps": steps, "interactions_per_s": (pc + pp) / wall,
            "nodes": info["nodes"], "buckets": info["buckets"], "pc_pairs": pc, "pp_pairs": pp, "result_rows": nrows,
            "rank0_phases_ms": {a: round(b / steps, 3) for a, b in phases.items()},
            "rank0_pc_ms": pc_ms, "rank0_pc_tflops": info["pc_pairs"] * FLOP_PC / (pc_ms * 1e-3) / 1e12,
            "rank0_pp_ms": taps["part_ms"] / steps, "rank0_ewald_ms": taps["ewald_ms"] / max(taps["ewald_launches"], 1),
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "finite": finite,
            "rank0_hbm_in_use_gb": round((total_b - free_b) / 1e9, 1),
            "timing": "wall clock around RawParticleStep.run() (max over ranks); phases and kernels by CUDA events on rank 0's stream"}


def gpu_local_affinity(index):
    """pin this process to the CPUs NVML names as local to GPU `index`; returns the core count or None"""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def workload_size(args, world):
    """None = the config's own particle set (at N=1 the reference's Tipsy fixture, committed under
    tests/golden/); at N > 1 the box grows with the GPU count (weak scaling), which only the
    synthetic generators can do: cube300 then means the 48^3-per-GPU Zel'dovich stand-in"""
    if args.n:
        return args.n * world
    if world > 1 and args.workload == "cube300":
        return 48 ** 3 * world
    return None


def run_reference(args, rank, world):
    """CPU arm: the oracle port (kind "port": gravity.h / Ewald.cpp need Charm++ and do not
    compile here) with all host threads; rank 0 only."""
    if rank != 0:
        return
    from changa_b200.workloads import config_workload, interaction_counts
    wl = config_workload(args.workload, n=workload_size(args, world),
                         bucket_range_of=(0, world) if world > 1 else None)
    cnt = interaction_counts(wl)
    pairs = cnt["cell"] + cnt["part"] + cnt["softcell"]
    cpu_force_step(wl)  # untimed: loads (and if needed builds) the checker library, spins up the thread pool
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    threads = 1
    for _ in range(steps):
        _, threads = cpu_force_step(wl)
    dt = (time.perf_counter() - t0) / steps
    val = pairs / dt
    sample = f"{steps} full force steps of rank 0's share ({pairs} pair interactions + Ewald on {len(wl['parts']) // world} particles)"
    print(json.dumps({
        "impl": "reference", "metric": "gravity_interactions_per_s", "value": val, "unit": "interactions/s",
        "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "note": "CPU restatement of nodeBucketForce/partBucketForce/BucketEwald, OpenMP over buckets"},
        "cpu_baseline": {"value": val, "unit": "interactions/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cube300", choices=["cube300", "king", "uniform", "clustered", "collapse"])
    ap.add_argument("--n", type=int, default=0, help="particles per GPU (default: the config's own size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--double", action="store_true",
                    help="the CUDA_USE_DOUBLE build (cudatype = double): roofline against the measured DFMA peak")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 50)")
    ap.add_argument("--large-kind", default="uniform", choices=["uniform", "clustered"],
                    help="particle distribution of the extra box (SURVEY 8d recipes C3 / C4)")
    ap.add_argument("--large-n", type=int, default=1 << 22,
                    help="particles of the extra box whose tree and lists are built on the device; "
                         "shared by all ranks at N > 1 (0: skip)")
    ap.add_argument("--no-graph", action="store_true", help="launch the resident step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--large-active-rung", type=int, default=0,
                    help="> 0: the extra box runs a multistep force step at this activeRung (rungs by local "
                         "density, SURVEY C4): active-bucket lists + Ewald markers made on the device")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = None
    if world > 1 and not os.environ.get("CB200_NO_AFFINITY"):
        # one process per GPU: run on the cores next to that GPU, so the pinned buffers (first touch)
        # and the copies they feed stay on the GPU's own socket instead of crossing the host fabric
        numa = gpu_local_affinity(local)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner (NCCL_DEBUG=VERSION/WARN) to STDOUT when the communicator
        # comes up; the JSON line must be the only thing there, so fd 1 points at stderr meanwhile
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from changa_b200.hostcuda import HostCUDA, ForceStep
    from changa_b200.workloads import config_workload, interaction_counts
    hc = HostCUDA(double=args.double, device=local)
    if args.double:
        args.large_n = 0  # the device-built-lists box is a float pipeline

    wl = config_workload(args.workload, n=workload_size(args, world),
                         bucket_range_of=(rank, world) if world > 1 else None)
    cnt = interaction_counts(wl)
    pairs = cnt["cell"] + cnt["part"] + cnt["softcell"]
    ew_real, ew_n = ewald_real_terms(wl)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident throughput (value, roofline) --------------------------------
    rs = ResidentStep(hc, wl, torch, dist, rank, world)
    rs.timed(0, max(args.warmup, 3))
    use_graph = world == 1 and not args.no_graph
    if use_graph:
        try:
            per_step = rs.capture()
            rs.timed(0, 3, graph=True)
        except Exception as e:  # a capture that fails must not take the line down: launch eagerly
            sys.stderr.write(f"CUDA graph capture failed ({e!r}); launching eagerly\n")
            torch.cuda.synchronize()
            use_graph = False
    if use_graph:
        ms_total = rs.timed(args.steps, 0, graph=True)
        launches = per_step * args.steps
    else:
        launches1 = hc.kernel_launches()
        ms_total = rs.timed(args.steps, 0)
        launches = hc.kernel_launches() - launches1
    # per-kernel times: a second pass with an event pair around every launch (not the timed one:
    # the event records sit between the kernels)
    tap_steps = max(1, min(args.steps, 50))
    hc.timing(True)
    rs.timed(tap_steps, 0)
    taps = hc.timing_read()
    hc.timing(False)

    # ---- end to end through the reference-facing ABI, host buffers -----------------------
    e2e_how = "wall clock around ForceStep.run() (C-ABI entry points, pinned host buffers)"
    if world > 1:
        # every rank uploads its slice only; NVLink replicates (ShardedForceStep).  Checked once against
        # the rank pushing everything through its own PCIe link: same rows, bit for bit.
        from changa_b200.hostcuda import ShardedForceStep
        ref = ForceStep(hc, wl)
        want = ref.run().copy()
        ref.free()
        fs, good = None, 0
        try:
            fs = ShardedForceStep(hc, wl, torch, dist, rank, world)
            got = fs.run()
            good = 1 if np.array_equal(got[fs.p0:fs.p1], want[fs.p0:fs.p1]) else 0
        except Exception as e:
            sys.stderr.write(f"rank {rank}: sharded step failed: {e!r}\n")
        same = torch.tensor([good], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)  # the ranks decide together: the sharded step holds collectives
        if int(same.item()) == 0:
            sys.stderr.write(f"rank {rank}: sharded upload unusable or different somewhere; timing the full upload\n")
            if fs is not None:
                fs.free()
            fs = ForceStep(hc, wl)
        else:
          e2e_how = ("wall clock around ShardedForceStep.run(): slice upload from pinned host buffers, packed arrays "
                     "all-gathered over NVLink, the reference's list requests, own rows back; equals ForceStep bitwise")
    else:
        fs = ForceStep(hc, wl)
    e2e_steps = args.e2e_steps or max(1, min(args.steps, 50))
    for _ in range(3):
        fs.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fs.run()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d, d2h = fs.h2d_bytes, fs.d2h_bytes
    fs.free()

    clocks = sampler.stop() if rank == 0 else None  # the sampled region ends here: nvidia-smi polling stalls the driver for
    # milliseconds at a time, which the long kernels of the extra box below would show
    large = None
    if args.large_n > 0:
        if world > 1:
            large = large_box_step(hc, args.large_n, torch, dist, rank, world, kind=args.large_kind,
                                       active_rung=args.large_active_rung)
        else:
            try:
                large = large_box_step(hc, args.large_n, torch, dist, rank, world, kind=args.large_kind,
                                       active_rung=args.large_active_rung)
            except Exception as e:  # extra information: never takes the headline line down
                large = {"error": repr(e)}

    # ---- aggregate over ranks ------------------------------------------------------------
    agg = torch.tensor([ms_total, e2e_s, taps["cell_ms"], taps["part_ms"], taps["ewald_ms"]], dtype=torch.float64, device="cuda")
    tot = torch.tensor([pairs, cnt["cell"], cnt["part"] + cnt["softcell"], ew_n, ew_real, launches, h2d, d2h],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, e2e_s, cell_ms, part_ms, ewald_ms = (float(x) for x in agg.tolist())
    g_pairs, g_pc, g_pp, g_ewn, g_ewreal, g_launch, g_h2d, g_d2h = (float(x) for x in tot.tolist())

    if rank == 0:
        peaks = measured_peaks()
        K = args.steps
        ms_step = ms_total / K
        value = g_pairs / (ms_step * 1e-3)
        # dominant kernel: particle-cell.  Per-launch figures of THIS rank (rank 0).
        cell_launches = max(taps["cell_launches"], 1)
        pc_ms = taps["cell_ms"] / cell_launches
        pc_tflops = cnt["cell"] * FLOP_PC / (pc_ms * 1e-3) / 1e12 if pc_ms > 0 else 0.0
        il_c = wl["cell"][0]
        pc_bytes = len(il_c) * (8 + 128) + int(wl["cell"][3].sum()) * (32 + 40)  # list + packed-cell gather + targets
        pp_ms = taps["part_ms"] / max(taps["part_launches"], 1)
        ew_ms = taps["ewald_ms"] / max(taps["ewald_launches"], 1)
        n_ewh = len(wl["ewald"]["ewt"]) if wl.get("ewald") else 0
        line = {
            "metric": "gravity_interactions_per_s", "value": value, "unit": "interactions/s",
            "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.double else "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "particles_total": len(wl["parts"]), "particles_per_gpu": len(wl["parts"]) // world,
                       "theta": 0.7, "expansion": "hexadecapole", "bucket_size": 12,
                       "pc_pairs": g_pc, "pp_pairs": g_pp, "ewald_particles": g_ewn,
                       "l2": "flushed between steps (256 MiB device write)",
                       "launch": "CUDA graph replay of one resident step" if use_graph else "eager launches",
                       "cpu_affinity": f"{numa} GPU-local cores per rank (NVML)" if numa else None,
                       "parallelism": f"buckets sharded by SFC range x{world}; packed particle and moment slices all-gathered per step" if world > 1 else "single GPU"},
            "force_step_ms": ms_step,
            "kernels": {"pc_ms": pc_ms, "pp_ms": pp_ms, "ewald_ms": ew_ms,
                        "pc_interactions_per_s": cnt["cell"] / (pc_ms * 1e-3) if pc_ms else None,
                        "pp_interactions_per_s": (cnt["part"] + cnt["softcell"]) / (pp_ms * 1e-3) if pp_ms else None,
                        "ewald_particles_per_s": ew_n / (ew_ms * 1e-3) if ew_ms else None,
                        "ewald_real_terms_per_particle": ew_real / max(ew_n, 1),
                        # the other two kernels against the same FMA peak, by the stated conventions
                        # (p-p 30 flop/pair; Ewald 350 per evaluated replica + 58 per h-vector)
                        "pp_tflops": (cnt["part"] + cnt["softcell"]) * FLOP_PP / (pp_ms * 1e-3) / 1e12 if pp_ms else None,
                        "ewald_tflops": (ew_real * FLOP_EW_REAL + ew_n * n_ewh * FLOP_EW_K) / (ew_ms * 1e-3) / 1e12 if ew_ms else None,
                        "note": "rank 0, CUDA events around each launch, separate eager pass"},
            "roofline": {"bound": "fp64_fma" if args.double else "fp32_fma",
                         "kernel": "cell_list_kernel (p-c hexadecapole, scalar FP64)" if args.double else "cell_list_x2_kernel (p-c hexadecapole, packed f32x2)",
                         "achieved": pc_tflops,
                         "peak": peaks["fp64_tflops" if args.double else "fp32_tflops"], "unit": "TFLOP/s",
                         "frac": pc_tflops / peaks["fp64_tflops" if args.double else "fp32_tflops"],
                         "traffic": peaks["pc_traffic"] if args.workload == "cube300" and not args.n and world == 1 and not args.double else None,
                         "traffic_source": peaks["pc_traffic_source"], "flop_per_pair": FLOP_PC, "achieved_ref170": pc_tflops * FLOP_PC_REF / FLOP_PC,
                         "peak_source": peaks["fp64_source" if args.double else "fp32_source"],
                         "hbm": {"algorithmic_bytes": pc_bytes, "achieved_gbs": pc_bytes / (pc_ms * 1e-3) / 1e9 if pc_ms else None,
                                 "peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["hbm_source"]}},
            "e2e": {"value": g_pairs / (e2e_s / e2e_steps), "unit": "interactions/s", "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "steps": e2e_steps, "h2d_bytes_per_step": g_h2d, "d2h_bytes_per_step": g_d2h,
                    "timing": e2e_how},
            "gpu_launches": int(g_launch),
            "clocks": clocks,
        }
        if large is not None:
            if "rank0_pc_tflops" in large:
                large["rank0_pc_frac_of_fp32_peak"] = large["rank0_pc_tflops"] / peaks["fp32_tflops"]
            line["large_box"] = large
        if world == 1 and not args.no_cpu_baseline:
            try:
                dt, threads = cpu_force_step(wl, repeats=2)
                line["cpu_baseline"] = {"value": pairs / dt, "unit": "interactions/s", "cores": threads, "kind": "port",
                                        "ms_per_step": dt * 1e3,
                                        "sample": f"full force step ({pairs} pair interactions + Ewald on {ew_n} particles), best of 2"}
            except Exception as e:  # the checker is optional for the measurement
                line["cpu_baseline"] = {"value": None, "error": repr(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
