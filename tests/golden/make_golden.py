"""Generates tests/golden/moments_kat.npz from the REFERENCE's own moments.c
(oracle/_ref/libmoments_ref.so = /root/reference/moments.c compiled unmodified,
see oracle/Makefile).  Run here, in the container that has /root/reference:

    python tests/golden/make_golden.py

The fixture lets the oracle stay pinned on machines without the reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

FM_N, MC_N = orc.FM_N, orc.MC_N


def main():
    orc.build()
    R = orc.ref_moments()
    assert R is not None, "needs /root/reference (oracle/_ref)"
    rng = np.random.default_rng(20261017)
    n = 64
    out = {}
    # momMakeFmomr
    mk_in = np.column_stack([rng.uniform(0.1, 3, n), rng.uniform(0.05, 2, n),
                             rng.normal(size=(n, 3))])
    mk_out = np.zeros((n, FM_N)); mk_d2 = np.zeros(n)
    for i in range(n):
        mk_d2[i] = R.momMakeFmomr(mk_out[i], *mk_in[i])
    out.update(make_in=mk_in, make_out=mk_out, make_d2=mk_d2)
    # momShiftFmomr on those moments
    sh_in = np.column_stack([rng.uniform(0.05, 2, n), rng.normal(size=(n, 3)) * 0.3])
    sh_out = mk_out.copy()
    for i in range(n):
        R.momShiftFmomr(sh_out[i], *sh_in[i])
    out.update(shift_in=sh_in, shift_out=sh_out)
    # momScaledAddFmomr / momRescaleFmomr / momMulAddFmomr / momAddFmomr
    ur = rng.uniform(0.1, 2, n); ua = rng.uniform(0.1, 2, n); mm = rng.uniform(0.1, 2, n)
    sa = sh_out.copy(); rs = sh_out.copy(); ma = sh_out.copy(); ad = sh_out.copy()
    for i in range(n):
        R.momScaledAddFmomr(sa[i], ur[i], mk_out[(i + 1) % n].copy(), ua[i])
        R.momRescaleFmomr(rs[i], ur[i], ua[i])
        R.momMulAddFmomr(ma[i], ur[i], mm[i], mk_out[(i + 3) % n].copy(), ua[i])
        R.momAddFmomr(ad[i], mk_out[(i + 5) % n].copy())
    out.update(ur=ur, ua=ua, mm=mm, scaled_add_out=sa, rescale_out=rs, mul_add_out=ma, add_out=ad)
    # momFmomr2Momc
    mc = np.zeros((n, MC_N))
    for i in range(n):
        R.momFmomr2Momc(sh_out[i].copy(), mc[i])
    out.update(momc_out=mc)
    # momEvalFmomrcm: target at r from the cell centre
    ev_r = rng.normal(size=(n, 3)) * 3 + np.sign(rng.normal(size=(n, 3))) * 2
    ev_u = rng.uniform(0.2, 1.5, n)
    ev_out = np.zeros((n, 5))
    for i in range(n):
        dirr = 1.0 / np.sqrt((ev_r[i] ** 2).sum())
        p = np.zeros(1); ax = np.zeros(1); ay = np.zeros(1); az = np.zeros(1); mg = np.zeros(1)
        R.momEvalFmomrcm(sh_out[i].copy(), ev_u[i], dirr, *ev_r[i], p, ax, ay, az, mg)
        ev_out[i] = [p[0], ax[0], ay[0], az[0], mg[0]]
    out.update(eval_r=ev_r, eval_u=ev_u, eval_out=ev_out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "moments_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
