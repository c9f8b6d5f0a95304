#!/usr/bin/env python
"""The reference's `teststep` test (teststep/Makefile:3-8, test_pg.param) with this repository's force
engine: king_soft.bin (36 000 particles), theta 0.7, hexadecapole, leapfrog (kick-drift-kick) to t = 1.0,
pass criterion |E_final + 32.19| < 0.005 (the golden log teststep/pkdtest.log ends at -32.1921134).
The reference steps particles on individual power-of-two rungs (dDelta 0.1, dEta 0.03); here every
particle takes the same step dDelta / 2^k, which is the same integrator on its finest rung.

  python tools/teststep_energy.py --engine oracle --div 16        CPU: host tree + walk, oracle forces
  python tools/teststep_energy.py --engine gpu --div 16           B200: cb200_step_run (tree, moments, lists, forces on the device)

Velocities come from the reference's file when it is present, else from tests/golden/king_velocities.npz."""
import argparse, json, os, struct, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def king_state():
    from changa_b200.workloads import fixture_particles
    pos, mass, soft = fixture_particles("king")
    path = "/root/reference/teststep/king_soft.bin"
    vfix = os.path.join(ROOT, "tests", "golden", "king_velocities.npz")
    if os.path.exists(path):
        raw = open(path, "rb").read()
        _, nb, _, ns, nd, nst = struct.unpack(">diiiii", raw[:28])
        off = 32 if len(raw) >= 32 + 4 * (12 * ns + 9 * nd + 11 * nst) else 28
        a = np.frombuffer(raw, dtype=">f4", count=nd * 9, offset=off).reshape(nd, 9)
        vel = a[:, 4:7].astype(np.float64)
    else:
        vel = np.load(vfix)["vel"].astype(np.float64)
    return np.asarray(pos, dtype=np.float64), np.asarray(mass, dtype=np.float64), np.asarray(soft, dtype=np.float64), vel


class OracleEngine:
    def forces(self, pos, mass, soft):
        from changa_b200.tree import Tree, tree_workload
        from oracle import oracle as orc
        ext = float(np.abs(pos).max()) * 1.0001
        t = Tree(pos, mass, soft, max_bucket=12, root_lo=(-ext,) * 3, root_hi=(ext,) * 3)
        wl = tree_workload(pos, mass, soft, theta=0.7, n_replicas=0, period=1.0, max_bucket=12, tree=t)
        parts = np.ascontiguousarray(wl["parts"], dtype=np.float64)
        mom = np.ascontiguousarray(wl["moments"], dtype=np.float64)
        v = np.zeros((len(parts), 5))
        orc.cell_list(parts, mom, *wl["cell"], wl["fperiod"], v)
        orc.part_list(parts, parts, *wl["part"], wl["fperiod"], v)
        if wl.get("softcell"):
            orc.part_list(parts, np.ascontiguousarray(wl["softcell"][4]), *wl["softcell"][:4], wl["fperiod"], v)
        out = np.zeros_like(v)
        out[wl["order"]] = v  # back to the caller's order
        wl["tree"].free()
        return out


class GpuEngine:
    def __init__(self):
        from changa_b200.hostcuda import HostCUDA
        self.hc = HostCUDA(double=False, device=0)

        self.st = None

    def forces(self, pos, mass, soft):
        from changa_b200.step import NativeStep
        if self.st is None:  # one step object for the whole run; the root box leaves room for the drift
            ext = float(np.abs(pos).max()) * 1.5
            self.st = NativeStep(self.hc, len(pos), theta=0.7, n_replicas=0, period=1.0, ewald=None,
                                 root_lo=(-ext,) * 3, root_hi=(ext,) * 3)
        self.st.set_particles(pos, mass, soft)
        self.st.run()  # cb200_step_run: tree, moments, lists and forces on the device, rows in the caller's order
        return np.asarray(self.st.out.array[: len(pos)], dtype=np.float64).copy()


def energy(mass, vel, f):
    T = 0.5 * float(np.sum(mass * (vel ** 2).sum(1)))
    U = 0.5 * float(np.sum(mass * f[:, 3]))
    return T + U, T, U


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="oracle", choices=["oracle", "gpu"])
    ap.add_argument("--div", type=int, default=16, help="step = dDelta / div")
    ap.add_argument("--t-end", type=float, default=1.0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    pos, mass, soft, vel = king_state()
    mass = np.broadcast_to(mass, (len(pos),)).astype(np.float64)
    soft = np.broadcast_to(soft, (len(pos),)).astype(np.float64)
    eng = OracleEngine() if a.engine == "oracle" else GpuEngine()
    dt = 0.1 / a.div
    nsteps = int(round(a.t_end / dt))
    t0 = time.time()
    f = eng.forces(pos, mass, soft)
    log = [(0.0,) + energy(mass, vel, f)]
    for k in range(nsteps):
        vel += 0.5 * dt * f[:, :3]
        pos = pos + dt * vel
        f = eng.forces(pos, mass, soft)
        vel += 0.5 * dt * f[:, :3]
        if (k + 1) % a.div == 0 or k + 1 == nsteps:
            log.append((round((k + 1) * dt, 6),) + energy(mass, vel, f))
            print("t = %.2f  E = %.6f  T = %.5f  U = %.5f" % log[-1], flush=True)
    E = log[-1][1]
    res = {"test": "teststep (king_soft.bin, theta 0.7, leapfrog to t = %g)" % a.t_end, "engine": a.engine,
           "step": dt, "force_evaluations": nsteps + 1, "E0": log[0][1], "E_final": E,
           "golden_E0": -32.1919514, "golden_E_final": -32.1921134,
           "criterion": "|E_final + 32.19| < 0.005 (teststep/Makefile:3-8)", "passed": bool(abs(E + 32.19) < 0.005),
           "log": [dict(zip(("t", "E", "T", "U"), r)) for r in log], "seconds": round(time.time() - t0, 1)}
    print(json.dumps({k: v for k, v in res.items() if k != "log"}))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
