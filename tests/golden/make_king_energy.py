"""The t = 0 line of the reference's own golden log for its teststep test (teststep/pkdtest.log: PKDGRAV on
king_soft.bin, theta 0.7; teststep/Makefile:3-8 accepts a ChaNGa run whose total energy stays within 0.005 of
-32.19) together with the kinetic energy of the file's velocities, which the position fixture does not hold.
Run in the container that has the reference:

    python tests/golden/make_king_energy.py        -> tests/golden/king_energy.json, king_velocities.npz
"""
import json
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LOG = "/root/reference/teststep/pkdtest.log"
BIN = "/root/reference/teststep/king_soft.bin"


def main():
    row = [l.split() for l in open(LOG) if not l.startswith("#")][0]
    raw = open(BIN, "rb").read()
    _, nbodies, _, nsph, ndark, nstar = struct.unpack(">diiiii", raw[:28])
    off = 32 if len(raw) >= 32 + 4 * (12 * nsph + 9 * ndark + 11 * nstar) else 28
    a = np.frombuffer(raw, dtype=">f4", count=ndark * 9, offset=off).reshape(ndark, 9).astype(np.float64)
    kinetic = 0.5 * float(np.sum(a[:, 0] * (a[:, 4:7] ** 2).sum(1)))
    out = {"source": "teststep/pkdtest.log, first data line (time, z, E, T, U, ...)",
           "time": float(row[0]), "E": float(row[2]), "T": float(row[3]), "U": float(row[4]),
           "kinetic_from_file": kinetic, "n": int(nbodies),
           "makefile_tolerance_on_E": 0.005}
    # the velocities themselves, for tools/teststep_energy.py on machines without the reference
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "king_velocities.npz"), vel=a[:, 4:7].astype(np.float32))
    dst = os.path.join(ROOT, "tests", "golden", "king_energy.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
