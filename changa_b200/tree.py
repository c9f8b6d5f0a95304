"""ctypes view of libcb200_host.so (csrc/treewalk.cpp): the single-TreePiece tree,
its moments and the Stadel double walk that produce the hot path's INPUT in the
shape TreePiece/DataManager hand to the GPU entry points.

Host-side only (no CUDA needed to load).  `tree_workload` returns the workload
dict described in changa_b200.workloads.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libcb200_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built: run __graft_entry__.build()")
        L = C.CDLL(path)
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        L.cb200h_tree_build.restype = vp
        L.cb200h_tree_build.argtypes = [vp, vp, vp, i, i, vp, vp]
        L.cb200h_tree_free.argtypes = [vp]
        L.cb200h_tree_sizes.argtypes = [vp, vp]
        L.cb200h_tree_export.argtypes = [vp] * 16
        L.cb200h_tree_export_links.argtypes = [vp, vp, vp, vp]
        L.cb200h_walk.restype = vp
        L.cb200h_walk.argtypes = [vp, d, i, d, vp, i, i]
        L.cb200h_lists_free.argtypes = [vp]
        L.cb200h_lists_sizes.argtypes = [vp, vp]
        L.cb200h_lists_stats.argtypes = [vp, vp]
        L.cb200h_lists_export.argtypes = [vp] * 7
        L.cb200h_expand_part_list.argtypes = [vp, vp, i, vp, vp]
        L.cb200h_num_threads.restype = i
        L.cb200h_ewald_tables.restype = i
        L.cb200h_ewald_tables.argtypes = [vp, d, d, vp, vp, i]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data


class Tree:
    """Particles sorted along the Morton curve + the binary-oct tree over them.
    Node arrays are in breadth-first order (= nodeArrayIndex, DataManager.cpp:797-828);
    buckets are in particle order (= bucketList)."""

    def __init__(self, pos, mass, soft, max_bucket=12, root_lo=(-0.5, -0.5, -0.5), root_hi=(0.5, 0.5, 0.5)):
        L = lib()
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = len(pos)
        mass = np.ascontiguousarray(np.broadcast_to(mass, (n,)), dtype=np.float64)
        soft = np.ascontiguousarray(np.broadcast_to(soft, (n,)), dtype=np.float64)
        lo = np.ascontiguousarray(root_lo, dtype=np.float64)
        hi = np.ascontiguousarray(root_hi, dtype=np.float64)
        self.h = L.cb200h_tree_build(_p(pos), _p(mass), _p(soft), n, int(max_bucket), _p(lo), _p(hi))
        sz = np.zeros(4, dtype=np.int32)
        L.cb200h_tree_sizes(self.h, _p(sz))
        self.num_nodes, self.num_buckets, self.num_levels, self.n = (int(x) for x in sz)
        nn, nb = self.num_nodes, self.num_buckets
        self.order = np.zeros(n, dtype=np.int32)
        self.parts = np.zeros((n, 5))
        self.moments = np.zeros((nn, 27))
        self.child0, self.child1 = np.zeros(nn, dtype=np.int32), np.zeros(nn, dtype=np.int32)
        self.first, self.last = np.zeros(nn, dtype=np.int32), np.zeros(nn, dtype=np.int32)
        self.level_start = np.zeros(self.num_levels + 1, dtype=np.int32)
        self.geolo, self.geohi = np.zeros((nn, 3)), np.zeros((nn, 3))
        self.boxlo, self.boxhi = np.zeros((nn, 3)), np.zeros((nn, 3))
        self.bucket_node = np.zeros(nb, dtype=np.int32)
        self.bucket_starts, self.bucket_sizes = np.zeros(nb, dtype=np.int32), np.zeros(nb, dtype=np.int32)
        L.cb200h_tree_export(self.h, _p(self.order), _p(self.parts), _p(self.moments), _p(self.child0),
                             _p(self.child1), _p(self.first), _p(self.last), _p(self.level_start),
                             _p(self.geolo), _p(self.geohi), _p(self.boxlo), _p(self.boxhi),
                             _p(self.bucket_node), _p(self.bucket_starts), _p(self.bucket_sizes))

        self.parent = np.zeros(nn, dtype=np.int32)
        self.bucket_first, self.bucket_count = np.zeros(nn, dtype=np.int32), np.zeros(nn, dtype=np.int32)
        L.cb200h_tree_export_links(self.h, _p(self.parent), _p(self.bucket_first), _p(self.bucket_count))

    def walk(self, theta=0.7, n_replicas=0, period=1.0, bucket_active=None, bucket_range=None):
        """Interaction lists of the active buckets in bucket_range (default: all).
        Returns a dict:
          cell  (Lc,2) int32 {nodeArrayIndex, offsetID}, cell_mark (nb+1) int64
          part  (Lp,3) int32 {first particle, offset code, count} (ILPart), part_mark
          soft  (Ls,2) int32 softened cells the reference evaluates as softened
                monopoles on the host (Compute.cpp:1683-1699), soft_mark
        markers cover every bucket of the tree (empty lists included)."""
        L = lib()
        nb = self.num_buckets
        act = None
        if bucket_active is not None:
            act = np.ascontiguousarray(bucket_active, dtype=np.uint8)
            assert len(act) == nb
        b0, b1 = (0, nb) if bucket_range is None else bucket_range
        h = L.cb200h_walk(self.h, float(theta), int(n_replicas), float(period), _p(act), int(b0), int(b1))
        sz = np.zeros(6, dtype=np.int64)
        L.cb200h_lists_sizes(h, _p(sz))
        out = {
            "cell": np.zeros((int(sz[0]), 2), dtype=np.int32), "cell_mark": np.zeros(nb + 1, dtype=np.int64),
            "part": np.zeros((int(sz[1]), 3), dtype=np.int32), "part_mark": np.zeros(nb + 1, dtype=np.int64),
            "soft": np.zeros((int(sz[2]), 2), dtype=np.int32), "soft_mark": np.zeros(nb + 1, dtype=np.int64),
            "expanded_part_entries": int(sz[3]), "mac_tests": int(sz[4]), "mac_opened": int(sz[5]),
        }
        st = np.zeros(8, dtype=np.int64)
        L.cb200h_lists_stats(h, _p(st))
        out["node_stats"] = dict(zip(("visited", "max_chk", "max_clist", "max_lplist", "max_undlist",
                                      "sum_clist", "sum_lplist", "sum_undlist"), (int(x) for x in st)))
        L.cb200h_lists_export(h, _p(out["cell"]), _p(out["cell_mark"]), _p(out["part"]), _p(out["part_mark"]),
                              _p(out["soft"]), _p(out["soft_mark"]))
        L.cb200h_lists_free(h)
        return out

    def expand_part_list(self, part, part_mark):
        """GenericList<ILPart>::serialize (Compute.cpp:1174-1187): one {index, off} per source particle"""
        L = lib()
        nb = len(part_mark) - 1
        part = np.ascontiguousarray(part, dtype=np.int32)
        part_mark = np.ascontiguousarray(part_mark, dtype=np.int64)
        em = np.zeros(nb + 1, dtype=np.int64)
        L.cb200h_expand_part_list(_p(part), _p(part_mark), nb, None, _p(em))
        ex = np.zeros((int(em[-1]), 2), dtype=np.int32)
        L.cb200h_expand_part_list(_p(part), _p(part_mark), nb, _p(ex), _p(em))
        return ex, em

    def free(self):
        if self.h:
            lib().cb200h_tree_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ewald_tables_fast(root_cell, L, dEwhCut=2.8):
    """C twin of changa_b200.ewald_tables.ewald_tables (microseconds instead of a millisecond)"""
    root = np.ascontiguousarray(root_cell, dtype=np.float64)
    momc = np.zeros(32)
    ewt = np.zeros((512, 5))
    n = lib().cb200h_ewald_tables(_p(root), float(L), float(dEwhCut), _p(momc), _p(ewt), 512)
    return momc, np.ascontiguousarray(ewt[:n])


def serialize(ilist, mark, starts, sizes):
    """GenericList<T>::serialize (Compute.cpp:1034-1111): buckets with empty lists are
    dropped, markers rebased to the flat list.  Returns (ilist, markers, starts, sizes, bucket ids)."""
    mark = np.asarray(mark, dtype=np.int64)
    ln = np.diff(mark)
    k = np.nonzero(ln > 0)[0]
    lo = int(mark[k[0]]) if len(k) else 0
    hi = int(mark[k[-1] + 1]) if len(k) else 0
    m = np.concatenate([[0], np.cumsum(ln[k])])
    assert m[-1] == hi - lo  # filled buckets are contiguous in the flat list by construction
    assert m[-1] < 2 ** 31
    return (np.ascontiguousarray(ilist[lo:hi], dtype=np.int32), m.astype(np.int32),
            np.ascontiguousarray(starts[k], dtype=np.int32), np.ascontiguousarray(sizes[k], dtype=np.int32), k)


def tree_workload(pos, mass, soft, theta=0.7, n_replicas=0, period=1.0, max_bucket=12, ewald=None,
                  bucket_active=None, bucket_range=None, name="tree", tree=None):
    """Tree + walk + serialize -> workload dict (see changa_b200.workloads).
    ewald: None | dict(dEwCut=2.6, dEwhCut=2.8) -> adds the Ewald tables of the root cell."""
    from .ewald_tables import ewald_tables
    t = tree or Tree(pos, mass, soft, max_bucket=max_bucket)
    w = t.walk(theta=theta, n_replicas=n_replicas, period=period, bucket_active=bucket_active,
               bucket_range=bucket_range)
    wl = {"parts": t.parts, "moments": t.moments, "fperiod": float(period) if (n_replicas or ewald is not None) else 0.0,
          "order": t.order, "name": name, "tree": t,
          "walk_stats": {k: w[k] for k in ("mac_tests", "mac_opened")}}
    wl["cell"] = serialize(w["cell"], w["cell_mark"], t.bucket_starts, t.bucket_sizes)[:4]
    ex, em = t.expand_part_list(w["part"], w["part_mark"])
    wl["part"] = serialize(ex, em, t.bucket_starts, t.bucket_sizes)[:4]
    wl["part_buckets"] = (w["part"], w["part_mark"])
    wl["softcell"] = None
    if len(w["soft"]):
        # softened cells become ad-hoc source particles {M, soft, cm} shipped with the request
        nodes, inv = np.unique(w["soft"][:, 0], return_inverse=True)
        src = np.column_stack([t.moments[nodes, 2], t.moments[nodes, 1], t.moments[nodes, 3:6]])
        il = np.column_stack([inv.astype(np.int32), w["soft"][:, 1]]).astype(np.int32)
        wl["softcell"] = serialize(il, w["soft_mark"], t.bucket_starts, t.bucket_sizes)[:4] + (src,)
    wl["ewald"] = None
    if ewald is not None:
        momc, ewt = ewald_tables(t.moments[0], period, ewald.get("dEwhCut", 2.8))
        act = None
        if bucket_active is not None or bucket_range is not None:
            live = np.zeros(t.num_buckets, dtype=bool)
            b0, b1 = (0, t.num_buckets) if bucket_range is None else bucket_range
            live[b0:b1] = True
            if bucket_active is not None:
                live &= np.asarray(bucket_active, dtype=bool)
            act = np.concatenate([np.arange(s, s + z) for s, z in
                                  zip(t.bucket_starts[live], t.bucket_sizes[live])] or
                                 [np.zeros(0, dtype=np.int64)]).astype(np.int32)
        wl["ewald"] = {"root": t.moments[0].copy(), "momc": momc, "ewt": ewt, "L": float(period),
                       "fEwCut": float(ewald.get("dEwCut", 2.6)), "nReps": int(n_replicas), "active": act}
    return wl
