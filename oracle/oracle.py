"""ctypes front-end of oracle/liboracle.so (and oracle/_ref when present).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from changa_b200/.
See the header of gravity_oracle.c for what is restated and how it is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF_MOM = None
_REF_GRAV = None

FM_N, MC_N, CM_N = 22, 32, 27
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """make liboracle.so (and _ref/ when /root/reference exists)."""
    so = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
            os.path.join(HERE, "gravity_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "liboracle.so")
        src = os.path.join(HERE, "gravity_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build()
        L = C.CDLL(so)
        d, i, vp = C.c_double, C.c_int, C.c_void_p
        L.orc_fm_add.argtypes = [dp, dp]
        L.orc_fm_scaled_add.argtypes = [dp, d, dp, d]
        L.orc_fm_rescale.argtypes = [dp, d, d]
        L.orc_fm_mul_add.argtypes = [dp, d, d, dp, d]
        L.orc_fm_make.argtypes = [dp, d, d, d, d, d]
        L.orc_fm_make.restype = d
        L.orc_fm_shift.argtypes = [dp, d, d, d, d]
        L.orc_fm_to_momc.argtypes = [dp, dp]
        L.orc_fm_eval.argtypes = [dp, d, d, d, d, d, dp, dp, dp, dp, dp]
        L.orc_spline.argtypes = [d, d, dp, dp]
        L.orc_cell_list.argtypes = [dp, dp, ip, ip, ip, ip, i, d, dp]
        L.orc_part_list.argtypes = [dp, dp, ip, ip, ip, ip, i, d, dp]
        L.orc_softened_cell.argtypes = [dp, i, i, d, d, dp, dp, dp]
        L.orc_ewald_root_momc.argtypes = [dp, dp]
        L.orc_ewald_init.argtypes = [dp, d, d, dp, i]
        L.orc_ewald_init.restype = i
        L.orc_ewald.argtypes = [dp, vp, i, dp, dp, d, d, i, i, d, dp, i, dp, vp]
        L.orc_build_moments.argtypes = [dp, dp, dp, ip, ip, ip, ip, dp, dp, dp, dp, i, dp]
        L.orc_open_softening.argtypes = [dp, dp, dp, dp, dp]
        L.orc_open_softening.restype = i
        L.orc_open_criterion_node.argtypes = [dp, i, dp, dp, dp, dp, i, d, d]
        L.orc_open_criterion_node.restype = i
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        _LIB = L
    return _LIB


def ref_moments():
    """oracle/_ref/libmoments_ref.so: the reference's moments.c compiled as is.
    Returns None where it has not been built (no /root/reference)."""
    global _REF_MOM
    if _REF_MOM is None:
        so = os.path.join(HERE, "_ref", "libmoments_ref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        d = C.c_double
        L.momMakeFmomr.argtypes = [dp, d, d, d, d, d]
        L.momMakeFmomr.restype = d
        L.momShiftFmomr.argtypes = [dp, d, d, d, d]
        L.momAddFmomr.argtypes = [dp, dp]
        L.momScaledAddFmomr.argtypes = [dp, d, dp, d]
        L.momRescaleFmomr.argtypes = [dp, d, d]
        L.momMulAddFmomr.argtypes = [dp, d, d, dp, d]
        L.momFmomr2Momc.argtypes = [dp, dp]
        L.momEvalFmomrcm.argtypes = [dp, d, d, d, d, d, dp, dp, dp, dp, dp]
        _REF_MOM = L
    return _REF_MOM


def ref_gravity():
    """oracle/_ref/libgravity_ref.so: the reference's own gravity.h (SPLINE, partBucketForce, nodeBucketForce,
    openSoftening, openCriterionBucket / Node) and Ewald.cpp (EwaldInit, BucketEwald) compiled unmodified through
    oracle/gravity_ref.cpp / ewald_ref.cpp, with the reference's moments.c linked in.  Returns None where it has not been built (no /root/reference)."""
    global _REF_GRAV
    if _REF_GRAV is None:
        so = os.path.join(HERE, "_ref", "libgravity_ref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        d, i, vp = C.c_double, C.c_int, C.c_void_p
        L.gref_set_theta.argtypes = [d, d]
        L.gref_spline.argtypes = [d, d, dp, dp]
        L.gref_splineq.argtypes = [d, d, d, dp]
        L.gref_part_bucket_force.argtypes = [dp, dp, dp, i, i, vp, i, dp]
        L.gref_part_bucket_force.restype = i
        L.gref_node_bucket_force.argtypes = [dp, dp, dp, dp, dp, dp, i, i, vp, i, dp]
        L.gref_node_bucket_force.restype = i
        L.gref_open_softening.argtypes = [dp, dp, dp, dp, dp]
        L.gref_open_softening.restype = i
        L.gref_open_criterion_node.argtypes = [dp, i, dp, dp, dp, dp, i]
        L.gref_open_criterion_node.restype = i
        L.gref_open_criterion_bucket.argtypes = [dp, i, dp, dp, dp, dp]
        L.gref_open_criterion_bucket.restype = i
        L.gref_build_moments.argtypes = [dp, dp, dp, ip, ip, ip, ip, dp, dp, dp, dp, i, dp]
        lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
        L.gref_step_create.argtypes = [dp, i, dp, i, ip, i, dp, dp, ip, ip, d, i, d, d, i]
        L.gref_step_create.restype = vp
        L.gref_step_run.argtypes = [vp, i, i, ip, lp, ip, lp, ip, lp, i]
        L.gref_step_vars.argtypes = [vp, i, i, dp]
        L.gref_step_destroy.argtypes = [vp]
        L.eref_init.argtypes = [dp, d, d, dp, dp, i]
        L.eref_init.restype = i
        L.eref_bucket_ewald.argtypes = [dp, d, d, d, i, dp, i, i, vp, i, dp]
        _REF_GRAV = L
    return _REF_GRAV


class ReferenceStep:
    """one force evaluation by the reference's own CPU routines (nodeBucketForce, partBucketForce, BucketEwald of
    gravity.h / Ewald.cpp compiled unmodified: oracle/ewald_ref.cpp::gref_step_*), on a tree given as flat arrays
    (changa_b200.tree.Tree) and the per-bucket lists of its walk.  OpenMP over buckets."""

    def __init__(self, tree, period, ewald=None, n_replicas=0):
        G = ref_gravity()
        assert G is not None, "oracle/_ref/libgravity_ref.so needs /root/reference at build time"
        self.G, self.t = G, tree
        ew = ewald or {}
        self._keep = [as_f64(tree.parts), as_f64(tree.moments), as_i32(tree.bucket_node), as_f64(tree.boxlo).reshape(-1),
                      as_f64(tree.boxhi).reshape(-1), as_i32(tree.first), as_i32(tree.last)]
        k = self._keep
        self.h = G.gref_step_create(k[0].reshape(-1), len(tree.parts), k[1].reshape(-1), len(tree.child0), k[2], len(tree.bucket_node),
                                    k[3], k[4], k[5], k[6], float(period), 1 if ewald is not None else 0,
                                    float(ew.get("dEwhCut", 2.8)), float(ew.get("dEwCut", 2.6)), int(n_replicas))

    def run(self, walk, b0, b1, threads=1):
        """walk: the dict Tree.walk returns (cell / part / soft with markers over all buckets)"""
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
        self.G.gref_step_run(self.h, int(b0), int(b1), i32(walk["cell"]), i64(walk["cell_mark"]), i32(walk["part"]),
                             i64(walk["part_mark"]), i32(walk["soft"]), i64(walk["soft_mark"]), int(threads))

    def vars(self, p0, p1):
        out = np.zeros((p1 - p0, 5))
        self.G.gref_step_vars(self.h, int(p0), int(p1), out.reshape(-1))
        return out

    def free(self):
        if self.h:
            self.G.gref_step_destroy(self.h)
            self.h = None


# ---- array helpers ---------------------------------------------------------
def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def cell_list(part, cells, ilist, markers, starts, sizes, fperiod, vars_):
    """part Nx5 {m,soft,x,y,z}; cells Mx27; ilist Lx2 {index,offsetID};
    accumulates into vars_ Nx5 {ax,ay,az,pot,dtGrav} (float64, in place)."""
    lib().orc_cell_list(part, cells, as_i32(ilist).reshape(-1), as_i32(markers), as_i32(starts),
                        as_i32(sizes), len(starts), float(fperiod), vars_)
    return vars_


def part_list(part, src, ilist, markers, starts, sizes, fperiod, vars_):
    lib().orc_part_list(part, src, as_i32(ilist).reshape(-1), as_i32(markers), as_i32(starts),
                        as_i32(sizes), len(starts), float(fperiod), vars_)
    return vars_


def ewald_tables(root_cell, L, dEwhCut, cap=4096):
    """(momc[32], ewt[n,5]) per Ewald.cpp:285-375."""
    momc = np.zeros(MC_N)
    lib().orc_ewald_root_momc(as_f64(root_cell), momc)
    ewt = np.zeros((cap, 5))
    n = lib().orc_ewald_init(momc, float(L), float(dEwhCut), ewt.reshape(-1), cap)
    assert n <= cap
    return momc, np.ascontiguousarray(ewt[:n])


def ewald(part, active, root_cell, momc, L, fEwCut, nReps, nEwReps, fInner2coef, ewt, vars_):
    act = None if active is None else as_i32(active)
    nact = len(part) if act is None else len(act)
    nreal = C.c_longlong(0)
    lib().orc_ewald(part, None if act is None else act.ctypes.data, nact, as_f64(root_cell),
                    as_f64(momc), float(L), float(fEwCut), int(nReps), int(nEwReps),
                    float(fInner2coef), as_f64(ewt).reshape(-1), len(ewt), vars_,
                    C.addressof(nreal))
    return nreal.value


def build_moments(pos, mass, soft, child0, child1, first, last, geolo, geohi, boxlo, boxhi):
    n = len(child0)
    out = np.zeros((n, CM_N))
    lib().orc_build_moments(as_f64(pos).reshape(-1), as_f64(mass), as_f64(soft), as_i32(child0),
                            as_i32(child1), as_i32(first), as_i32(last),
                            as_f64(geolo).reshape(-1), as_f64(geohi).reshape(-1),
                            as_f64(boxlo).reshape(-1), as_f64(boxhi).reshape(-1), n,
                            out.reshape(-1))
    return out
