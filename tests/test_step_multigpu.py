"""The in-library multi-GPU step (cb200_comm_* / cb200_step_*).

CPU part: the cost-feedback cut rule (cb200_cost_targets) as a pure host function.
GPU part (needs two devices; `gpurun --gpus 2`): two processes, one per GPU, an NCCL communicator made
from an id handed over in a file, the same box stepped by both -- every rank's rows must equal the
single-GPU result bit for bit, and the ranks' bucket ranges must tile the box."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cost_targets_balance_a_lopsided_box():
    """ranks that measured more cost get narrower ranges; equal costs keep equal counts; targets are
    monotone, start at 0 and end at n; a second application on the predicted costs is a fixed point"""
    from changa_b200.step import cost_targets
    n = 1_000_000
    even = np.linspace(0, n, 5).astype(np.int64)
    t = cost_targets(even, [1.0, 1.0, 1.0, 1.0], n)
    assert list(t) == list(even)
    t = cost_targets(even, [4.0, 1.0, 1.0, 2.0], n)
    assert t[0] == 0 and t[-1] == n and np.all(np.diff(t) > 0)
    # cost density is piecewise constant: the cost inside each new range is total / world
    dens = np.repeat(np.array([4.0, 1.0, 1.0, 2.0]) / 250_000, 250_000)
    loads = np.add.reduceat(dens, t[:-1])
    assert np.allclose(loads, loads.mean(), rtol=1e-4)
    t2 = cost_targets(t, loads, n)
    assert np.abs(t2 - t).max() <= 1
    # a rank that measured nothing (a void) does not break the rule
    t = cost_targets(even, [0.0, 3.0, 0.0, 1.0], n)
    assert t[0] == 0 and t[-1] == n and np.all(np.diff(t) >= 0)


WORKER = r"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, {root!r})
rank, world, idfile, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
from changa_b200.hostcuda import HostCUDA
from changa_b200.step import Comm, NativeStep
from changa_b200.workloads import clustered_box
hc = HostCUDA(double=False, device=rank)
if rank == 0:
    uid = Comm.make_id(hc.L)
    open(idfile + ".tmp", "wb").write(uid)
    os.rename(idfile + ".tmp", idfile)
else:
    for _ in range(600):
        if os.path.exists(idfile):
            break
        time.sleep(0.1)
    uid = open(idfile, "rb").read()
comm = Comm(hc.L, rank, world, uid)
pos, mass, soft = clustered_box(60000, seed=5)
rows = {{}}
for cost in (0, 1):
    st = NativeStep(hc, len(pos), theta=0.7, n_replicas=1, period=1.0, ewald={{}}, comm=comm, cost_cuts=bool(cost))
    st.set_particles(pos, mass, soft)
    for it in range(2 + cost):        # the cost rule needs one step to measure
        res = st.run()
        if cost == 0 and it == 0:
            np.savez(os.path.join(outdir, f"rank{{rank}}_first.npz"), let=np.array([res.letBlockLevel, res.letFallback]))
    np.savez(os.path.join(outdir, f"rank{{rank}}_cost{{cost}}.npz"), idx=st.idx.array[:res.rows].copy(),
             rows=st.out.array[:res.rows].copy(), range=np.array([res.bucketLo, res.bucketHi, res.partLo, res.partHi]),
             cost=res.cost, pairs=np.array([res.pcPairs, res.ppPairs]), let=np.array([res.letBlockLevel, res.letFallback]))
    st.free()
tot = comm.allreduce([1.0, rank], "sum")
assert tot[0] == world and tot[1] == world * (world - 1) / 2
comm.barrier()
comm.destroy()
print("rank", rank, "ok")
"""


@pytest.mark.gpu
@pytest.mark.parametrize("let_blocks", [0, 64, -64])
def test_two_ranks_over_nccl_equal_the_single_gpu_step(tmp_path, let_blocks, monkeypatch):
    """let_blocks = 64: the locally essential moment build (csrc/let_kernels.cuh) is switched on for this small box
    (block level = first level with 64 x world nodes; by default a box needs 32768 x world): every rank builds only
    the subtrees near its own buckets below that level, exchanges the block records, and must still reproduce the
    single-GPU rows bit for bit without falling back to the full build"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    if let_blocks:
        monkeypatch.setenv("CB200_LET_BLOCKS_PER_RANK", str(abs(let_blocks)))
        if let_blocks < 0:  # a halo far too thin: the walk must notice, fall back to the full build and still be exact
            monkeypatch.setenv("CB200_LET_RADIUS_SCALE", "0.02")
    else:
        monkeypatch.setenv("CB200_LET", "0")
    from changa_b200.hostcuda import HostCUDA
    from changa_b200.step import NativeStep
    from changa_b200.workloads import clustered_box
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    hc = HostCUDA(double=False, device=0)
    pos, mass, soft = clustered_box(60000, seed=5)
    st = NativeStep(hc, len(pos), theta=0.7, n_replicas=1, period=1.0, ewald={})
    st.set_particles(pos, mass, soft)
    res = st.run()
    want = st.out.array[: len(pos)].copy()
    nb, pairs1 = res.numBuckets, (res.pcPairs, res.ppPairs)
    st.free()
    for cost in (0, 1):
        got = np.full_like(want, np.nan)
        r = [np.load(tmp_path / f"rank{k}_cost{cost}.npz") for k in range(2)]
        assert r[0]["range"][0] == 0 and r[0]["range"][1] == r[1]["range"][0] and r[1]["range"][1] == nb
        assert r[0]["range"][2] == 0 and r[0]["range"][3] == r[1]["range"][2] and r[1]["range"][3] == len(pos)
        for k in range(2):
            got[r[k]["idx"]] = r[k]["rows"]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))       # bit for bit, every particle once
        for k in range(2):
            level, fallback = (int(x) for x in r[k]["let"])
            if let_blocks >= 0:
                assert fallback == 0
                assert (level >= 0) == bool(let_blocks), (level, let_blocks)
        if let_blocks < 0:  # the first step fell back (both ranks together); later steps build everything
            assert all(int(np.load(tmp_path / f"rank{k}_first.npz")["let"][1]) == 1 for k in range(2))
        assert tuple(r[0]["pairs"] + r[1]["pairs"]) == pairs1                    # same work, split not duplicated
        if cost:  # the measured costs of the two ranks are closer than with equal particle counts
            c = np.array([float(r[0]["cost"]), float(r[1]["cost"])])
            c0 = np.array([float(np.load(tmp_path / f"rank{k}_cost0.npz")["cost"]) for k in range(2)])
            assert abs(c[0] - c[1]) <= abs(c0[0] - c0[1]) + 0.02 * c.sum()
