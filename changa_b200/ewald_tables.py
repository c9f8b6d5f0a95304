"""Host-side Ewald set-up: the tables TreePiece::EwaldInit builds once per step and
TreePiece::EwaldGPU ships to the device (Ewald.cpp:285-375, 387-517).

Written as symmetric-tensor contractions: with T2, T3, T4 the COMPLETE (not
trace-free-reduced) moment tensors of the root cell and h an integer wave
vector, the reference's QEVAL (Ewald.cpp:9-46) reduces for the potential term to
    cos part:  -(gam0 M + gam2 T2:hh/2 + gam4 T4::hhhh/24)
    sin part:  -(gam3 T3:.hhh/6)
where gam_k = gam0 (2 pi/L)^k with the sign pattern (+,+,-,-,+,+)."""
import itertools
import math

import numpy as np

# order of the 21 reduced components in CudaMultipoleMoments after {radius, soft, mass, cm}
_CELL_NAMES = ["xx", "xy", "xz", "yy", "yz",
               "xxx", "xyy", "xxy", "yyy", "xxz", "yyz", "xyz",
               "xxxx", "xyyy", "xxxy", "yyyy", "xxxz", "yyyz", "xxyy", "xxyz", "xyyz"]
# MomcData field order (EwaldCUDA.h:29-37)
MOMC_NAMES = ["m", "xx", "yy", "xy", "xz", "yz",
              "xxx", "xyy", "xxy", "yyy", "xxz", "yyz", "xyz",
              "xxxx", "xyyy", "xxxy", "yyyy", "xxxz", "yyyz", "xxyy", "xxyz", "xyyz",
              "zz", "xzz", "yzz", "zzz", "xxzz", "xyzz", "xzzz", "yyzz", "yzzz", "zzzz"]
_AX = {"x": 0, "y": 1, "z": 2}


def _canon(name):
    return "".join(sorted(name))


def complete_tensors(root_cell):
    """(M, T2, T3, T4): unscaled complete moment tensors from a 27-value cell record.
    Stored components are scaled by radius^order (FMOMR convention, moments.c:238-267) and
    trace-free, so components with two or more z indices follow from T[..zz] = -(T[..xx]+T[..yy])."""
    radius = float(root_cell[0])
    comp = {}
    for name, v in zip(_CELL_NAMES, root_cell[6:27]):
        comp[_canon(name)] = float(v) * radius ** len(name)
    for order in (2, 3, 4):
        names = sorted({"".join(c) for c in itertools.combinations_with_replacement("xyz", order)},
                       key=lambda s: s.count("z"))
        for nm in names:
            if nm in comp:
                continue
            rest = nm.replace("z", "", 2)  # strip one zz pair
            comp[nm] = -(comp[_canon(rest + "xx")] + comp[_canon(rest + "yy")])
    tens = []
    for order in (2, 3, 4):
        T = np.zeros((3,) * order)
        for idx in itertools.product(range(3), repeat=order):
            T[idx] = comp[_canon("".join("xyz"[i] for i in idx))]
        tens.append(T)
    return float(root_cell[2]), tens[0], tens[1], tens[2], comp


def ewald_tables(root_cell, L, dEwhCut=2.8):
    """momc (32 complete root moments, MomcData order) and ewt rows {hx,hy,hz,hCfac,hSfac}."""
    M, T2, T3, T4, comp = complete_tensors(np.asarray(root_cell, dtype=np.float64))
    momc = np.array([M] + [comp[_canon(n)] for n in MOMC_NAMES[1:]])
    hreps = int(math.ceil(dEwhCut))
    rng = np.arange(-hreps, hreps + 1)
    H = np.array([(a, b, c) for a in rng for b in rng for c in rng], dtype=np.float64)  # hx outermost
    h2 = (H ** 2).sum(1)
    H = H[(h2 > 0) & (h2 <= dEwhCut * dEwhCut)]
    h2 = (H ** 2).sum(1)
    alpha = 2.0 / L
    k4 = math.pi ** 2 / (alpha * alpha * L * L)
    c = 2.0 * math.pi / L
    g0 = np.exp(-k4 * h2) / (math.pi * h2 * L)
    g2, g3, g4 = -c ** 2 * g0, -c ** 3 * g0, c ** 4 * g0
    q2 = np.einsum("ij,ni,nj->n", T2, H, H) / 2.0
    q3 = np.einsum("ijk,ni,nj,nk->n", T3, H, H, H) / 6.0
    q4 = np.einsum("ijkl,ni,nj,nk,nl->n", T4, H, H, H, H) / 24.0
    hC = -(g0 * M + g2 * q2 + g4 * q4)
    hS = -(g3 * q3)
    ewt = np.column_stack([c * H, hC, hS])
    return momc, np.ascontiguousarray(ewt)
