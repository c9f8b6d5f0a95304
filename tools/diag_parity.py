"""where does the worst-particle error of the periodic tree workload come from?"""
import numpy as np, sys
sys.path.insert(0, ".")
from changa_b200.hostcuda import HostCUDA, ForceStep
from changa_b200.workloads import config_workload
from oracle import oracle as orc
hc = HostCUDA(double=False, device=0)  # CB200_LIB selects another build
wl = config_workload("cube300", n=16 ** 3)
f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).astype(np.float64))
parts, mom = f32(wl["parts"]), f32(wl["moments"])
def gpu(w):
    s = ForceStep(hc, w); out = s.run().copy().astype(np.float64); s.free(); return out
lists_gpu = gpu(dict(wl, ewald=None))
ew_gpu = gpu(dict(wl, cell=None, part=None, softcell=None))
v = np.zeros((len(parts), 5)); orc.cell_list(parts, mom, *wl["cell"], 1.0, v); orc.part_list(parts, parts, *wl["part"], 1.0, v)
ew = wl["ewald"]; e = np.zeros((len(parts), 5))
orc.ewald(parts, None, f32(ew["root"]), f32(ew["momc"]), 1.0, ew["fEwCut"], ew["nReps"], 3, 1.1e-2, f32(ew["ewt"]), e)
tot = v + e
amag = np.linalg.norm(tot[:, :3], axis=1)
for name, g, o in (("lists", lists_gpu, v), ("ewald", ew_gpu, e)):
    d = np.linalg.norm(g[:, :3] - o[:, :3], axis=1)
    i = int(np.argmax(d / amag))
    print(name, "max |da|/|a_total| = %.3g at %d; |a_part|=%.3g |a_tot|=%.3g median %.3g" % ((d / amag).max(), i, np.linalg.norm(o[i, :3]), amag[i], np.median(d / amag)),
          "r_to_root_cm=%.4f" % np.linalg.norm(parts[i, 2:5] - ew["root"][3:6]))
d = np.linalg.norm(ew_gpu[:, :3] - e[:, :3], axis=1) / amag
r = np.linalg.norm(parts[:, 2:5] - ew["root"][3:6], axis=1)
for lo, hi in ((0, .05), (.05, .1), (.1, .12), (.12, .2), (.2, .5), (.5, 1)):
    k = (r >= lo) & (r < hi)
    if k.any(): print("r in [%.2f,%.2f): n=%d max ewald err %.3g median %.3g" % (lo, hi, k.sum(), d[k].max(), np.median(d[k])))
