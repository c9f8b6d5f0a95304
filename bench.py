#!/usr/bin/env python
"""bench.py -- one force step of ChaNGa's GPU gravity hot path on N B200s, on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload uniform256] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

Workload (default): BASELINE.json configs[2] -- the synthetic uniform 256^3 dark-matter box of the
reference's testdata/ppartt.c (srand(1); x, y, z = -0.5 + rand()/RAND_MAX; m = 1/N; eps = N^(-1/3)/20),
periodic, nReplicas = 1 + Ewald, theta = 0.7, hexadecapole, bucket size 12.  At N > 1 the SAME box is
shared (strong scaling); at N = 8 a `target` block adds BASELINE's target, the clustered 512^3 box.

A "step" is one whole force evaluation from UNSORTED particles through the in-library step
(cb200_step_run): [upload of the rank's 40-byte records ->] one NCCL all-gather -> keys, sort, tree,
boxes, moments -> interaction lists (double walk) -> particle-cell (hexadecapole) -> particle-particle ->
softened cells -> Ewald [-> rows back in the caller's order].

  value     pair interactions (p-c + p-p, counted like Compute.cpp:1643-1651) per second of the whole
            step with the particle records already in HBM, CUDA events on the step's stream, max over ranks
  e2e       the same through the C ABI from pinned HOST buffers: H2D of the records and D2H of the
            accelerations inside the timed region, wall clock, max over ranks
  roofline  the p-c kernel (and the p-p kernel) against the measured FP32 FMA peak, 198 / 30 flop per pair
  parity    buckets re-evaluated by the CPU oracle from the step's own lists and moments: median / max |da|/|a|
  cpu_baseline                      the reference's own CPU gravity (kind "reference": the --impl reference arm as a
                                    CPU-only subprocess on a bounded bucket range of the same box, all host cores), with
                                    the oracle port on a bucket range of the step's own lists next to it ("port": also the
                                    fallback, and the parity_range check)
  --impl reference                  the reference's OWN CPU gravity (gravity.h / Ewald.cpp compiled unmodified into
                                    oracle/_ref/libgravity_ref.so; kind "reference") on all host cores, on a bounded
                                    bucket range of the same box; the oracle port where that library is absent
  ref_cuda  the reference's own HostCUDA.cu kernels (compiled unmodified) on the same requests, next to
            this library's reference-facing entry points on them

No part of the GPU arm runs through oracle/.
"""
import argparse
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
THETA, BUCKET = 0.7, 12
EWALD = {"dEwCut": 2.6, "dEwhCut": 2.8}
WORKLOADS = {  # name -> (generator, particles)
    "uniform256": ("uniform", 1 << 24),     # BASELINE configs[2]
    "clustered512": ("clustered", 1 << 27),  # BASELINE configs[3] / the north-star target
    "clustered256": ("clustered", 1 << 24),
    "uniform4m": ("uniform", 1 << 22),
    "clustered4m": ("clustered", 1 << 22),
}


def measured_peaks():
    out = {"fp32_tflops": 71.64, "fp32_source": "fallback: tools/fp32_peak.cu FFMA run of round 1",
           "hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "profiles", "r01_fp32_peak.json")
    if os.path.exists(p):
        try:
            out["fp32_tflops"] = float(json.load(open(p))["ffma_tflops"])
            out["fp32_source"] = "measured: tools/fp32_peak.cu scalar FFMA chains on this pool's B200 (profiles/r01_fp32_peak.json)"
        except Exception:
            pass
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            out["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            out["hbm_source"] = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    out["pc_traffic"], out["pc_traffic_source"] = None, None
    import glob
    # the latest ncu capture of the p-c kernel in the 256^3 step (session tags sort as r02a < ... < r02z < r02aa < ...)
    caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02*_ncu_cell_list_x2_256.json")) +
                  glob.glob(os.path.join(ROOT, "profiles", "r02*_ncu_pc_256.json")),
                  key=lambda f: (len(os.path.basename(f).split("_")[0]), os.path.basename(f)))
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
        num = lambda s: s.replace(".", "").isdigit()
        sm = [float(r[1]) for r in self.rows if num(r[1])]
        mx = [float(r[2]) for r in self.rows if num(r[2])]
        pw = [float(r[3]) for r in self.rows if num(r[3])]
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] if pw and len(pw) == len(sm) else sm
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis and all(x.strip().isdigit() for x in vis.split(",")) and local < len(vis.split(",")):
        return int(vis.split(",")[local])
    return local


def gpu_local_affinity(index):
    """pin this process to the CPUs NVML names as local to GPU `index`; returns the core count or None"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(index))
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


def make_rows(kind, n, lo, hi):
    """rows [lo, hi) of the box: (positions, mass, softening); equal masses 1/N, eps = N^(-1/3)/20"""
    from changa_b200.workloads import uniform_box, clustered_box_rows
    if kind == "uniform":  # glibc rand() is one sequential stream (ppartt.c): the whole box is generated
        pos, mass, soft = uniform_box(n, seed=1)
        return np.ascontiguousarray(pos[lo:hi]), float(mass[0]), float(soft[0])
    return clustered_box_rows(n, lo, hi, seed=2)


# ------------------------------------------------------------------------------------------------
# the CPU oracle as checker and as baseline (test infrastructure: only these legs touch oracle/)
# ------------------------------------------------------------------------------------------------
class DeviceProducts:
    """what the step left on the device, brought to the host for the oracle: sorted particles, double
    moments, accumulators, and the interaction lists of chosen bucket ranges"""

    def __init__(self, st):
        self.st = st
        tr = st.L.cb200_step_tree(st.handle).contents
        n, nn = tr.numParticles, tr.numNodes
        self.n, self.nn, self.nb = n, nn, tr.numBuckets
        pos = st._download(tr.d_pos, 3 * n, np.float64).reshape(n, 3)
        mass = st._download(tr.d_mass, n, np.float64)
        soft = st._download(tr.d_soft, n, np.float64)
        f32 = lambda a: a.astype(np.float32).astype(np.float64)  # inputs as the float kernels see them
        self.parts = np.ascontiguousarray(np.column_stack([f32(mass), f32(soft), f32(pos)]))
        del pos, mass, soft
        mom = st.moments()
        self.root = mom[0].copy()
        self.moments = np.ascontiguousarray(f32(mom))
        del mom
        # softened cells travel as source particles {M, soft, cm} (Compute.cpp:1683-1699)
        self.node_parts = np.ascontiguousarray(np.column_stack([self.moments[:, 2], self.moments[:, 1], self.moments[:, 3:6]]))
        self.vars = st.vars()
        self.vbuf = np.zeros((self.n, 5))

    def oracle_range(self, b0, b1, orc, period, n_reps):
        """oracle accelerations of buckets [b0, b1): returns (particle indices, oracle rows, pairs, seconds)"""
        li = self.st.bucket_lists(b0, b1)
        starts, sizes = li["starts"], li["sizes"]
        idx = np.concatenate([np.arange(s, s + z) for s, z in zip(starts, sizes)]) if len(starts) else np.zeros(0, np.int64)
        v = self.vbuf
        ln = lambda key: np.diff(li[key + "_mark"]).astype(np.int64)
        pairs = int(((ln("cell") + ln("part") + ln("soft")) * sizes).sum())
        momc, ewt = orc.ewald_tables(self.root, period, EWALD["dEwhCut"])
        f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
        t0 = time.perf_counter()
        orc.cell_list(self.parts, self.moments, li["cell"], li["cell_mark"], starts, sizes, period, v)
        orc.part_list(self.parts, self.parts, li["part"], li["part_mark"], starts, sizes, period, v)
        if len(li["soft"]):
            orc.part_list(self.parts, self.node_parts, li["soft"], li["soft_mark"], starts, sizes, period, v)
        orc.ewald(self.parts, idx.astype(np.int32), f32(self.root), f32(momc), float(np.float32(period)), EWALD["dEwCut"],
                  n_reps, int(np.ceil(EWALD["dEwCut"])), 1.2e-3, f32(ewt), v)
        dt = time.perf_counter() - t0
        rows = v[idx].copy()
        v[idx] = 0.0
        return idx, rows, pairs, dt, li


def parity_stats(got, want):
    got = np.asarray(got, dtype=np.float64)
    amag = np.linalg.norm(want[:, :3], axis=1)
    da = np.linalg.norm(got[:, :3] - want[:, :3], axis=1)
    rel = da / np.maximum(amag, 1e-300)
    # in a nearly uniform periodic box the net force on a particle is the small residue of cancelling
    # terms: the worst-particle figure divides by max(|a|, 0.1 rms|a|) as the tests do
    rel_floor = da / np.maximum(amag, 0.1 * np.sqrt((amag ** 2).mean()))
    pot = np.abs(got[:, 3] - want[:, 3]) / np.maximum(np.abs(want[:, 3]), 1e-300)
    dt = np.abs(got[:, 4] - want[:, 4]) / np.maximum(np.abs(want[:, 4]), 1e-300)
    return {"particles": int(len(got)), "median_da_over_a": float(np.median(rel)), "p99_da_over_a": float(np.quantile(rel, 0.99)),
            "max_da_over_max_a_floor": float(rel_floor.max()), "median_dpot_over_pot": float(np.median(pot)),
            "max_ddtgrav_over_dtgrav": float(dt.max()), "finite": bool(np.isfinite(got).all())}


def parity_block(st, prod, orc, period, n_reps, rng_buckets, n_runs=16, run_len=128, seed=7):
    """n_runs runs of run_len consecutive buckets at random places of this rank's range, re-evaluated by the
    oracle from the step's own lists and moments"""
    b_lo, b_hi = rng_buckets
    rng = np.random.default_rng(seed)
    run_len = min(run_len, max(1, (b_hi - b_lo) // n_runs))
    got, want, pairs = [], [], 0
    for b0 in sorted(rng.integers(b_lo, max(b_lo + 1, b_hi - run_len), n_runs).tolist()):
        idx, rows, p, _, _ = prod.oracle_range(b0, b0 + run_len, orc, period, n_reps)
        got.append(prod.vars[idx])
        want.append(rows)
        pairs += p
    out = parity_stats(np.concatenate(got), np.concatenate(want))
    out.update(buckets=n_runs * run_len, pairs=pairs,
               how="CPU oracle (double, oracle/gravity_oracle.c) on the step's own lists and moments, float-rounded inputs; "
                   "tolerance of the north star: median |da|/|a| <= 1e-4 in float")
    out["within_tolerance"] = bool(out["median_da_over_a"] <= 1e-4 and out["finite"])
    return out


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU gravity (or the port of it) on the host cores, CPU only
# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """the reference's own CPU gravity (kind "reference": gravity.h / Ewald.cpp compiled unmodified into
    oracle/_ref/libgravity_ref.so; the oracle port, kind "port", where that library is absent) with all host threads
    on a bounded bucket range of the same box; tree and lists from the host walk (changa_b200/csrc/treewalk.cpp) --
    no GPU is touched.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import oracle as orc
    from changa_b200.tree import Tree
    from changa_b200.ewald_tables import ewald_tables
    kind, n = WORKLOADS[args.workload]
    n = args.n or n
    threads = len(os.sched_getaffinity(0))
    orc.lib().orc_set_num_threads(threads)  # torchrun pins OMP_NUM_THREADS=1
    pos, mass, soft = make_rows(kind, n, 0, n)
    t = Tree(pos, np.full(n, mass), np.full(n, soft), max_bucket=BUCKET)
    del pos
    nb = t.num_buckets
    frac = min(1.0, args.ref_pairs / (650.0 * n))  # ~650 pair interactions per particle at theta = 0.7
    b0 = nb // 3
    b1 = min(nb, b0 + max(64, int(frac * nb)))
    w = t.walk(theta=THETA, n_replicas=1, period=1.0, bucket_range=(b0, b1))
    from changa_b200.tree import serialize
    cell = serialize(w["cell"], w["cell_mark"], t.bucket_starts, t.bucket_sizes)[:4]
    ex, em = t.expand_part_list(w["part"], w["part_mark"])
    part = serialize(ex, em, t.bucket_starts, t.bucket_sizes)[:4]
    soft_l = None
    if len(w["soft"]):
        soft_l = serialize(w["soft"], w["soft_mark"], t.bucket_starts, t.bucket_sizes)[:4]
        node_parts = np.ascontiguousarray(np.column_stack([t.moments[:, 2], t.moments[:, 1], t.moments[:, 3:6]]))
    parts, mom = np.ascontiguousarray(t.parts), np.ascontiguousarray(t.moments)
    act = np.concatenate([np.arange(s, s + z) for s, z in zip(t.bucket_starts[b0:b1], t.bucket_sizes[b0:b1])]).astype(np.int32)
    momc, ewt = ewald_tables(t.moments[0], 1.0, EWALD["dEwhCut"])
    pairs = sum(int((np.diff(l[1]).astype(np.int64) * l[3].astype(np.int64)).sum()) for l in (cell, part, soft_l) if l)

    v = np.zeros((n, 5))

    def port_step():
        v[act] = 0.0
        orc.cell_list(parts, mom, *cell, 1.0, v)
        orc.part_list(parts, parts, *part, 1.0, v)
        if soft_l:
            orc.part_list(parts, node_parts, *soft_l, 1.0, v)
        orc.ewald(parts, act, t.moments[0], momc, 1.0, EWALD["dEwCut"], 1, int(np.ceil(EWALD["dEwCut"])), 1.2e-3, ewt, v)

    # kind "reference": the reference's OWN nodeBucketForce / partBucketForce / BucketEwald (gravity.h and Ewald.cpp
    # compiled unmodified into oracle/_ref/libgravity_ref.so, which travels with the repo) on the same lists, OpenMP
    # over buckets; the oracle port is the fallback where that library was never built
    kind_ran, step, check = "port", port_step, None
    if not args.ref_port and orc.ref_gravity() is not None:
        rs = orc.ReferenceStep(t, 1.0, ewald=EWALD, n_replicas=1)
        kind_ran, step = "reference", (lambda: rs.run(w, b0, b1, threads=threads))
        step()
        port_step()
        got, want = rs.vars(int(act[0]), int(act[-1]) + 1), v[int(act[0]):int(act[-1]) + 1]
        amag = np.maximum(np.linalg.norm(want[:, :3], axis=1), 1e-300)
        check = float((np.linalg.norm(got[:, :3] - want[:, :3], axis=1) / amag).max())  # port vs reference: rounding level

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = pairs / dt
    sample = (f"buckets [{b0}, {b1}) of {nb} ({len(act)} particles, {pairs} pair interactions + their Ewald sums) per step; "
              f"throughput per pair is what scales to the whole box")
    print(json.dumps({
        "impl": "reference", "metric": "gravity_interactions_per_s", "value": val, "unit": "interactions/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(kind, n), "theta": THETA, "expansion": "hexadecapole", "bucket_size": BUCKET,
                   "note": ("the reference's own nodeBucketForce / partBucketForce / BucketEwald (gravity.h, Ewald.cpp compiled "
                            "unmodified, scalar double path; oracle/_ref/libgravity_ref.so)" if kind_ran == "reference" else
                            "CPU restatement of nodeBucketForce / partBucketForce / BucketEwald") +
                           ", OpenMP over buckets, on lists from the host walk; the same box as the GPU arm, a bounded bucket "
                           "range per step",
                   "port_vs_reference_max_da_over_a": check},
        "cpu_baseline": {"value": val, "unit": "interactions/s", "cores": threads, "kind": kind_ran, "sample": sample},
        "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def reference_cpu_baseline(pairs, kind, n, timeout=600):
    """cpu_baseline of kind "reference" for the GPU arm's line: the reference arm (this file, `--impl reference`:
    the reference's own gravity.h / Ewald.cpp compiled unmodified, all host cores) run as a CPU-only subprocess on a
    bounded bucket range of the same box.  None when it is not available (no oracle/_ref/libgravity_ref.so) or fails:
    the caller then keeps the oracle port's numbers."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgravity_ref.so")):
        return None
    wl = [k for k, v in WORKLOADS.items() if v[0] == kind]
    if not wl:
        return None
    try:
        env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK")}
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", wl[0], "--n", str(n),
                            "--steps", "1", "--warmup", "1", "--ref-pairs", str(pairs)], capture_output=True, text=True,
                           timeout=timeout, env=env, cwd=ROOT)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        j = json.loads(line[-1])
        c = j["cpu_baseline"]
        if c.get("kind") != "reference" or not c.get("value"):
            return None
        return {"value": c["value"], "unit": c["unit"], "cores": c["cores"], "kind": "reference", "seconds": j["ms_per_step"] * 1e-3,
                "sample": c["sample"] + "; " + j["config"]["note"]}
    except Exception:
        return None


def workload_name(kind, n):
    side = round(n ** (1.0 / 3.0))
    cube = f"{side}^3 = " if side ** 3 == n else ""
    recipe = "testdata/ppartt.c recipe, seed 1" if kind == "uniform" else "SURVEY C4 recipe: 30% uniform + 70% in Plummer halos"
    return f"{kind} periodic box, {cube}{n} particles ({recipe}), theta={THETA}, nReplicas=1 + Ewald, hexadecapole, bucket={BUCKET}"


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def time_steps(st, comm, steps, resident):
    """K steps bracketed by a barrier + synchronize on both sides: (wall seconds, sum of the steps' CUDA-event
    milliseconds, last result), wall / events as max over ranks"""
    hc = st.hc
    hc.stream_synchronize(st.stream)
    comm.barrier(st.stream)
    hc.L.cb200_device_synchronize()
    ev = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        res = st.run(resident=resident, sfc_order=True)  # rows in SFC order + caller indices at every N
        ev += float(res.ms[len(res.ms) - 1])
    hc.L.cb200_device_synchronize()
    wall = time.perf_counter() - t0
    comm.barrier(st.stream)
    hc.L.cb200_device_synchronize()
    agg = comm.allreduce([wall, ev], "max", st.stream)
    return float(agg[0]), float(agg[1]), res


def bench_box(hc, comm, kind, n, steps, warmup, local, want_parity=True, want_cpu=True, want_refcuda=True,
              want_single_gpu_check=True, cpu_pairs=2.5e9, sample_clocks=True):
    """everything measured on one box; returns the dict the JSON line is made of"""
    from changa_b200.step import NativeStep
    rank, world = comm.rank, comm.world
    st = NativeStep(hc, n, theta=THETA, n_replicas=1, period=1.0, ewald=EWALD, max_bucket=BUCKET, comm=comm)
    lo, hi = st.my_rows()
    pos, mass, soft = make_rows(kind, n, lo, hi)
    st.set_rows(pos, mass, soft)
    del pos
    launches0 = hc.kernel_launches()
    for _ in range(max(warmup, 3)):  # the first steps also size the memory pools
        res = st.run()
    launches_per_step = (hc.kernel_launches() - launches0) / max(warmup, 3)
    sampler = ClockSampler(physical_gpu_index(local))
    if rank == 0 and sample_clocks:
        sampler.start()
    e2e_wall, e2e_ev, res = time_steps(st, comm, steps, resident=False)
    h2d, d2h = int(res.h2dBytes), int(res.d2hBytes)
    st.upload()
    st.run(resident=True)
    res_wall, res_ev, res = time_steps(st, comm, steps, resident=True)
    clocks = sampler.stop() if (rank == 0 and sample_clocks) else None
    # per-kernel times: a pass with an event pair around the three force kernels; phases from the step's own events
    tap_steps = max(1, min(steps, 5))
    hc.timing(True)
    ph = {}
    for _ in range(tap_steps):
        res = st.run(resident=True)
        for k, v in st.phases().items():
            ph[k] = ph.get(k, 0.0) + v / tap_steps
    taps = hc.timing_read()
    hc.timing(False)
    tot = comm.allreduce([res.pcPairs, res.ppPairs, res.rows, h2d, d2h, res.nCell, res.nPart + res.nSoft], "sum", st.stream)
    mx = comm.allreduce([res.pcPairs, res.ppPairs, taps["cell_ms"] / tap_steps, taps["part_ms"] / tap_steps], "max", st.stream)
    out = {"n": n, "kind": kind, "world": world, "steps": steps,
           "pc_pairs": float(tot[0]), "pp_pairs": float(tot[1]), "rows": float(tot[2]), "h2d": float(tot[3]), "d2h": float(tot[4]),
           "list_entries": float(tot[5] + tot[6]),
           "nodes": res.numNodes, "buckets": res.numBuckets, "levels": res.numLevels,
           "let_block_level": int(res.letBlockLevel), "let_fallback": int(res.letFallback),
           "e2e_s_per_step": e2e_wall / steps, "e2e_event_ms_per_step": e2e_ev / steps,
           "resident_ms_per_step": res_ev / steps, "resident_wall_ms_per_step": res_wall / steps * 1e3,
           "launches_per_step": launches_per_step, "clocks": clocks,
           "rank0_phases_ms": {k: round(v, 3) for k, v in ph.items()},
           "rank0": {"pc_pairs": int(res.pcPairs), "pp_pairs": int(res.ppPairs),
                     "pc_ms": taps["cell_ms"] / tap_steps, "pp_ms": taps["part_ms"] / tap_steps,
                     "ewald_ms": taps["ewald_ms"] / tap_steps, "ewald_particles": int(res.partHi - res.partLo),
                     "list_entries_cell": int(res.nCell), "list_entries_part": int(res.nPart + res.nSoft),
                     "bucket_range": [res.bucketLo, res.bucketHi], "particle_range": [res.partLo, res.partHi]},
           "max_rank": {"pc_pairs": float(mx[0]), "pp_pairs": float(mx[1]), "pc_ms": float(mx[2]), "pp_ms": float(mx[3])}}
    free_b = None
    try:
        import torch
        free_b, total_b = torch.cuda.mem_get_info(local)
        out["rank0_hbm_in_use_gb"] = round((total_b - free_b) / 1e9, 1)
    except Exception:
        pass

    # ---- parity against the oracle (rank 0's range) and against the single-GPU step (every rank) ----
    if want_parity or want_cpu or want_refcuda:
        res = st.run(resident=True, keep_lists=True)
        prod = DeviceProducts(st) if rank == 0 else None
    if want_parity and rank == 0:
        from oracle import oracle as orc
        orc.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
        try:
            out["parity"] = parity_block(st, prod, orc, 1.0, 1, (res.bucketLo, res.bucketHi))
        except Exception as e:
            out["parity"] = {"error": repr(e)}
    if want_cpu and rank == 0:
        from oracle import oracle as orc
        threads = len(os.sched_getaffinity(0))
        orc.lib().orc_set_num_threads(threads)
        try:
            frac = min(1.0, cpu_pairs / max(res.pcPairs + res.ppPairs, 1))
            span = max(64, int(frac * (res.bucketHi - res.bucketLo)))
            b0 = res.bucketLo + (res.bucketHi - res.bucketLo - span) // 2
            idx, rows, pairs, dt, li = prod.oracle_range(b0, b0 + span, orc, 1.0, 1)
            out["cpu_baseline"] = {"value": pairs / dt, "unit": "interactions/s", "cores": int(orc.lib().orc_num_threads()),
                                   "kind": "port", "seconds": dt,
                                   "sample": f"buckets [{b0}, {b0 + span}) of {res.numBuckets}: {len(idx)} particles, {pairs} pair "
                                             f"interactions + their Ewald sums, one pass of the oracle port (OpenMP over buckets)"}
            ref = reference_cpu_baseline(min(cpu_pairs, 1e9), kind, n)
            if ref is not None:  # the reference's own CPU code did the same kind of sample: that is the baseline
                out["cpu_baseline"] = dict(ref, port=out["cpu_baseline"])
            out["parity_range"] = parity_stats(prod.vars[idx], rows)
            out["parity_range"]["buckets"] = span
            if want_refcuda:
                out["ref_cuda"] = ref_cuda_block(hc, prod, li, b0, min(b0 + span, b0 + max(64, span // 8)), rows_of=(idx, rows))
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "error": repr(e)}
    prod = None
    if world > 1 and want_single_gpu_check:
        # the sharded step against the same box stepped by ONE GPU (this one): my rows, bit for bit
        res = st.run()  # e2e form: rows + caller indices on the host
        rows_idx, rows = st.idx.array[: res.rows].copy(), st.out.array[: res.rows].copy()
        st.free()
        one = NativeStep(hc, n, theta=THETA, n_replicas=1, period=1.0, ewald=EWALD, max_bucket=BUCKET, comm=None)
        pos, mass, soft = make_rows(kind, n, 0, n)
        one.set_rows(pos, mass, soft)
        del pos
        one.run()
        same = bool(np.array_equal(one.out.array[rows_idx].view(np.uint32), rows.view(np.uint32)))
        one.free()
        flags = comm.allreduce([1.0 if same else 0.0, res.rows], "sum")
        out["parity_vs_n1"] = {"bitwise_equal": bool(flags[0] == world), "ranks_equal": int(flags[0]), "rows_checked": int(flags[1]),
                               "how": "every rank re-steps the whole box alone on its own GPU and compares its rows bit for bit"}
    else:
        st.free()
    return out


def ref_cuda_block(hc, prod, li_all, b0, b1, rows_of):
    """the reference's own HostCUDA.cu + CUDAMoments.cu, compiled unmodified for sm_100a (oracle/_ref), and this
    library's reference-facing entry points (DataManagerTransferLocalTree, TreePiece*ListDataTransferLocal,
    EwaldHost, TransferParticleVarsBack) on the SAME requests: buckets [b0, b1) of the box with their lists as
    host arrays, whole particle and moment arrays uploaded"""
    from oracle import ref_cuda
    if not ref_cuda.available():
        return {"unavailable": "oracle/_ref/libhostcuda_ref.so not built"}
    from changa_b200.hostcuda import ForceStep
    from changa_b200.tree import ewald_tables_fast
    li = prod.st.bucket_lists(b0, b1)
    starts, sizes = li["starts"], li["sizes"]
    act = np.concatenate([np.arange(s, s + z) for s, z in zip(starts, sizes)]).astype(np.int32)
    momc, ewt = ewald_tables_fast(prod.root, 1.0, EWALD["dEwhCut"])
    wl = {"name": "range", "parts": prod.parts, "moments": prod.moments, "fperiod": 1.0,
          "cell": (li["cell"], li["cell_mark"], starts, sizes), "part": (li["part"], li["part_mark"], starts, sizes),
          "softcell": None,
          "ewald": {"root": prod.root, "momc": momc, "ewt": ewt, "L": 1.0, "fEwCut": EWALD["dEwCut"], "nReps": 1, "active": act}}
    ln = lambda key: np.diff(li[key + "_mark"]).astype(np.int64)
    pairs = int(((ln("cell") + ln("part")) * sizes).sum())
    ref_rows, ref_s = ref_cuda.RefCuda().force_step(wl, repeats=3)
    fs = ForceStep(hc, wl)
    fs.run()
    t0 = time.perf_counter()
    for _ in range(3):
        ours = fs.run()
    ours_s = (time.perf_counter() - t0) / 3
    ours = np.asarray(ours, dtype=np.float64)[act].copy()
    h2d = fs.h2d_bytes
    fs.free()
    idx_all, want_all = rows_of
    where = np.searchsorted(idx_all, act)
    want = want_all[where]
    # softened cells are not part of these requests (the reference evaluates them on the host CPU,
    # Compute.cpp:1683-1699): compare the two GPU paths with each other
    ref_rows = np.asarray(ref_rows, dtype=np.float64)[act]
    amag = np.maximum(np.linalg.norm(want[:, :3], axis=1), 1e-300)
    diff = np.linalg.norm(ours[:, :3] - ref_rows[:, :3], axis=1) / amag
    return {"what": "HostCUDA.cu + CUDAMoments.cu compiled unmodified (-use_fast_math) vs this library's reference-facing "
                    "entry points, same requests, end to end from pinned host buffers (upload, list requests, EwaldHost, copy back)",
            "buckets": int(b1 - b0), "pair_interactions": pairs, "h2d_bytes": int(h2d),
            "reference_cuda_ms": ref_s * 1e3, "ours_ms": ours_s * 1e3, "speedup": ref_s / ours_s,
            "reference_cuda_interactions_per_s": pairs / ref_s, "ours_interactions_per_s": pairs / ours_s,
            "median_rel_diff_between_the_two": float(np.median(diff))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uniform256", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--particles", dest="n", type=int, default=0,
                    help="particles of the box (default: the workload's own size); under torchrun use --particles "
                         "(its parser takes --n for an abbreviation of its own options)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-target", action="store_true", help="skip the 512^3 clustered block of an 8-GPU run")
    ap.add_argument("--target", action="store_true", help="add the target block at any GPU count")
    ap.add_argument("--target-n", type=int, default=1 << 27)
    ap.add_argument("--cpu-pairs", type=float, default=2.5e9, help="pair interactions of the cpu_baseline sample")
    ap.add_argument("--ref-pairs", type=float, default=6e8, help="pair interactions per step of --impl reference")
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port even where oracle/_ref/libgravity_ref.so exists")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa = gpu_local_affinity(local) if (world > 1 and not os.environ.get("CB200_NO_AFFINITY")) else None
    kind, n = WORKLOADS[args.workload]
    n = args.n or n
    # NCCL writes its version banner to STDOUT when the communicator comes up; the JSON line must be the
    # only thing there, so fd 1 points at stderr until the line is printed
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    from changa_b200.hostcuda import HostCUDA
    from changa_b200.step import Comm
    hc = HostCUDA(double=False, device=local)  # aborts without a CUDA device: there is no CPU fallback
    comm = Comm.from_env(hc.L)
    box = bench_box(hc, comm, kind, n, args.steps, args.warmup, local, want_parity=not args.no_parity,
                    want_cpu=(world == 1 and not args.no_cpu_baseline), want_refcuda=(world == 1 and not args.no_ref_cuda),
                    cpu_pairs=args.cpu_pairs)
    target = None
    if (world == 8 and not args.no_target) or args.target:
        try:
            t = bench_box(hc, comm, "clustered", args.target_n, steps=2, warmup=1, local=local, want_parity=True, want_cpu=False,
                          want_refcuda=False, want_single_gpu_check=False, sample_clocks=False)
            pairs = t["pc_pairs"] + t["pp_pairs"]
            target = {"workload": workload_name("clustered", args.target_n), "n_gpus": world, "steps": 2,
                      "force_step_ms": t["resident_ms_per_step"], "interactions_per_s": pairs / (t["resident_ms_per_step"] * 1e-3),
                      "e2e_ms": t["e2e_s_per_step"] * 1e3, "e2e_interactions_per_s": pairs / t["e2e_s_per_step"],
                      "pc_pairs": t["pc_pairs"], "pp_pairs": t["pp_pairs"], "rank0_phases_ms": t["rank0_phases_ms"],
                      "rank0": t["rank0"], "max_rank": t["max_rank"], "parity": t.get("parity"),
                      "rank0_hbm_in_use_gb": t.get("rank0_hbm_in_use_gb"), "nodes": t["nodes"], "buckets": t["buckets"],
                      "let_block_level": t.get("let_block_level"), "let_fallback": t.get("let_fallback")}
        except Exception as e:
            target = {"error": repr(e)}

    if rank == 0:
        peaks = measured_peaks()
        pairs = box["pc_pairs"] + box["pp_pairs"]
        ms_step = box["resident_ms_per_step"]
        r0 = box["rank0"]
        pc_tflops = r0["pc_pairs"] * FLOP_PC / (r0["pc_ms"] * 1e-3) / 1e12 if r0["pc_ms"] > 0 else 0.0
        pp_tflops = r0["pp_pairs"] * FLOP_PP / (r0["pp_ms"] * 1e-3) / 1e12 if r0["pp_ms"] > 0 else 0.0
        pc_bytes = r0["list_entries_cell"] * (8 + 128) + (r0["particle_range"][1] - r0["particle_range"][0]) * (32 + 40)
        line = {
            "metric": "gravity_interactions_per_s", "value": pairs / (ms_step * 1e-3), "unit": "interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(kind, n), "particles_total": n, "theta": THETA, "expansion": "hexadecapole",
                       "bucket_size": BUCKET, "pc_pairs": box["pc_pairs"], "pp_pairs": box["pp_pairs"],
                       "nodes": box["nodes"], "buckets": box["buckets"], "tree_levels": box["levels"],
                       "moment_build": (f"locally essential below tree level {box['let_block_level']}: a rank builds the subtrees near its "
                                        f"own buckets, block records exchanged by one integer all-reduce (fallbacks: {box['let_fallback']})")
                       if box.get("let_block_level", -1) >= 0 else "every rank builds every node",
                       "l2": "inputs larger than L2: 40 B x N particle records, > 10 GB of lists per step; nothing is reused between steps",
                       "launch": "cb200_step_run: eager launches on the step's stream, 3 host synchronisations per step",
                       "cpu_affinity": f"{numa} GPU-local cores per rank (NVML)" if numa else None,
                       "parallelism": (f"one process per GPU x{world}: buckets sharded by SFC range (cost-balanced), one NCCL "
                                       f"all-gather of the 40-byte particle records per step, tree topology replicated")
                       if world > 1 else "single GPU"},
            "force_step_ms": ms_step,
            "timing": {"value": "sum of the steps' CUDA-event times (first enqueue to last kernel) on the step's stream, max over ranks; "
                                "K steps bracketed by barrier + synchronize",
                       "resident_wall_ms_per_step": box["resident_wall_ms_per_step"],
                       "e2e_event_ms_per_step": box["e2e_event_ms_per_step"]},
            "phases_ms_rank0": box["rank0_phases_ms"],
            "kernels": {"pc_ms": r0["pc_ms"], "pp_ms": r0["pp_ms"], "ewald_ms": r0["ewald_ms"],
                        "pc_interactions_per_s": r0["pc_pairs"] / (r0["pc_ms"] * 1e-3) if r0["pc_ms"] else None,
                        "pp_interactions_per_s": r0["pp_pairs"] / (r0["pp_ms"] * 1e-3) if r0["pp_ms"] else None,
                        "ewald_particles_per_s": r0["ewald_particles"] / (r0["ewald_ms"] * 1e-3) if r0["ewald_ms"] else None,
                        "pp_tflops": pp_tflops, "pp_frac_of_fp32_peak": pp_tflops / peaks["fp32_tflops"],
                        "note": "rank 0, CUDA events around each launch (separate pass); the Ewald kernel runs on a second stream under the tree walk"},
            "roofline": {"bound": "fp32_fma", "kernel": "cell_list_x2_kernel (p-c hexadecapole, packed f32x2)",
                         "achieved": pc_tflops, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s", "frac": pc_tflops / peaks["fp32_tflops"],
                         "traffic": peaks["pc_traffic"], "traffic_source": peaks["pc_traffic_source"],
                         "flop_per_pair": FLOP_PC, "achieved_ref170": pc_tflops * FLOP_PC_REF / FLOP_PC,
                         "peak_source": peaks["fp32_source"],
                         "hbm": {"algorithmic_bytes": pc_bytes, "achieved_gbs": pc_bytes / (r0["pc_ms"] * 1e-3) / 1e9 if r0["pc_ms"] else None,
                                 "peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["hbm_source"]}},
            "e2e": {"value": pairs / box["e2e_s_per_step"], "unit": "interactions/s", "ms_per_step": box["e2e_s_per_step"] * 1e3,
                    "steps": args.steps, "h2d_bytes_per_step": box["h2d"], "d2h_bytes_per_step": box["d2h"],
                    "timing": "wall clock around cb200_step_run from pinned host records to accelerations in pinned host memory "
                              "(each rank's rows in SFC order with their caller indices, copied back slab by slab under the list "
                              "kernels), max over ranks"},
            "gpu_launches": int(round(box["launches_per_step"] * args.steps * 2)),
            "clocks": box["clocks"],
            "hbm_in_use_gb_rank0": box.get("rank0_hbm_in_use_gb"),
        }
        for key in ("parity", "parity_range", "parity_vs_n1", "cpu_baseline", "ref_cuda"):
            if key in box:
                line[key] = box[key]
        if world > 1:
            line["load_balance"] = {"rank0": r0, "max_rank": box["max_rank"]}
        if target is not None:
            line["target"] = target
        sys.stdout.flush()
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
    comm.barrier()
    comm.destroy()


if __name__ == "__main__":
    main()
