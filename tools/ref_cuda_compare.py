"""Times the reference's own CUDA path (oracle/_ref/libhostcuda_ref.so: HostCUDA.cu compiled
unmodified for sm_100a) against this library on the same serialized requests, end to end
(upload -> lists -> Ewald -> copy back, wall clock, best of 5).  Writes one JSON line."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from changa_b200.hostcuda import HostCUDA, ForceStep
from changa_b200.workloads import config_workload, interaction_counts
from oracle import ref_cuda

name = sys.argv[1] if len(sys.argv) > 1 else "cube300"
wl = config_workload(name)
wl_ref = dict(wl, softcell=None)          # the reference evaluates softened cells on the host
cnt = interaction_counts(wl)
hc = HostCUDA(double=False, device=0)
fs = ForceStep(hc, wl)
for _ in range(3):
    fs.run()
best = 1e9
for _ in range(10):
    t0 = time.perf_counter(); ours = fs.run().copy(); best = min(best, time.perf_counter() - t0)
fs.free()
ref, tref = ref_cuda.RefCuda().force_step(wl_ref, repeats=8)
pairs = cnt["cell"] + cnt["part"]
print(json.dumps({"workload": wl["name"], "pairs": pairs, "ours_ms": best * 1e3, "reference_cuda_ms": tref * 1e3,
                  "speedup": tref / best, "note": "end to end through the host entry points, pinned host buffers, wall clock"}))
