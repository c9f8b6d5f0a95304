"""Particle positions of the reference's three Tipsy fixtures, as float32 (what the files hold),
so that BASELINE.json's configs 1, 2 and 5 run on the reference's own particle sets on machines
without /root/reference (the GPU box).  Masses and softenings are constant in all three files and
are stored as scalars.  Run in the container that has the reference:

    python tests/golden/make_fixture_positions.py        -> tests/golden/fixture_positions.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from changa_b200.workloads import read_tipsy  # noqa: E402

FILES = {
    "cube300": "/root/reference/testcosmo/cube300.tbin",
    "king": "/root/reference/teststep/king_soft.bin",
    "collapse": "/root/reference/testcollapse/adiabtophat_glass_28721.bin",
}


def main():
    out = {}
    for name, path in FILES.items():
        pos, mass, soft = read_tipsy(path)
        assert np.ptp(mass) == 0 and np.ptp(soft) == 0, name
        out[name + "_pos"] = pos.astype(np.float32)
        out[name + "_mass"] = np.float64(mass[0])
        out[name + "_soft"] = np.float64(soft[0])
        print(name, len(pos), "particles, mass", mass[0], "soft", soft[0])
    dst = os.path.join(ROOT, "tests", "golden", "fixture_positions.npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
