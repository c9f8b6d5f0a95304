"""ctypes shim over the in-library force step (include/changa_b200_api.h: cb200_comm_* / cb200_step_*).

Everything on the timed path is C++/CUDA inside libchanga_b200.so: the record upload, the NCCL
all-gather, tree, moments, walk, forces, Ewald, the copy back.  Python only holds the pinned host
buffers and hands pointers over -- what a one-process-per-device Charm++ host (DataManager.h:329-337)
would do from C++.  No torch on this path (torch's TCPStore is used once, at start-up, to hand the
128-byte NCCL id from rank 0 to the other ranks when the launcher is torchrun)."""
import ctypes as C
import os

import numpy as np

from . import lib as _lib

PHASES = ["h2d", "gather", "tree", "moments", "ewald", "walk", "pc", "pp", "finish", "total"]


class StepConfig(C.Structure):
    _fields_ = [("numParticles", C.c_longlong), ("maxBucket", C.c_int), ("nReplicas", C.c_int), ("ewald", C.c_int),
                ("activeRung", C.c_int), ("overlapEwald", C.c_int), ("costCuts", C.c_int),
                ("theta", C.c_double), ("period", C.c_double), ("dEwCut", C.c_double), ("dEwhCut", C.c_double),
                ("rootlo", C.c_double * 3), ("roothi", C.c_double * 3)]


class StepResult(C.Structure):
    _fields_ = [("error", C.c_int), ("numNodes", C.c_int), ("numBuckets", C.c_int), ("numLevels", C.c_int),
                ("bucketLo", C.c_int), ("bucketHi", C.c_int), ("partLo", C.c_int), ("partHi", C.c_int),
                ("activeBuckets", C.c_int), ("activeParticles", C.c_int), ("rows", C.c_int),
                ("nCell", C.c_longlong), ("nSoft", C.c_longlong), ("nPart", C.c_longlong),
                ("pcPairs", C.c_longlong), ("ppPairs", C.c_longlong),
                ("h2dBytes", C.c_longlong), ("d2hBytes", C.c_longlong), ("cost", C.c_double),
                ("ms", C.c_float * len(PHASES)), ("letBlockLevel", C.c_int), ("letFallback", C.c_int),
                ("treeRebuilt", C.c_int), ("walkRepeated", C.c_int)]


def _bind(L):
    if getattr(L, "_step_bound", False):
        return L
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.cb200_comm_id_bytes.restype = sz
    L.cb200_comm_unique_id.argtypes = [vp]
    L.cb200_comm_init.argtypes = [i, i, vp]
    L.cb200_comm_init.restype = vp
    L.cb200_comm_destroy.argtypes = [vp]
    L.cb200_comm_rank.argtypes = [vp]
    L.cb200_comm_world.argtypes = [vp]
    L.cb200_comm_nccl_version.restype = i
    L.cb200_comm_allreduce_f64.argtypes = [vp, vp, i, i, vp]
    L.cb200_comm_barrier.argtypes = [vp, vp]
    L.cb200_comm_allgather.argtypes = [vp, vp, vp, sz, vp]
    L.cb200_step_create.argtypes = [vp, C.POINTER(StepConfig)]
    L.cb200_step_create.restype = vp
    L.cb200_step_destroy.argtypes = [vp]
    L.cb200_step_chunk_rows.argtypes = [vp]
    L.cb200_step_out_capacity.argtypes = [vp]
    L.cb200_step_stream.argtypes = [vp]
    L.cb200_step_stream.restype = vp
    L.cb200_step_device_records.argtypes = [vp]
    L.cb200_step_device_records.restype = vp
    L.cb200_step_device_rungs.argtypes = [vp]
    L.cb200_step_device_rungs.restype = vp
    L.cb200_step_run.argtypes = [vp, vp, vp, vp, vp, i, i, C.POINTER(StepResult)]
    T = L.types
    L.cb200_step_tree.argtypes = [vp]
    L.cb200_step_tree.restype = C.POINTER(T.DevTree)
    L.cb200_step_lists.argtypes = [vp]
    L.cb200_step_lists.restype = C.POINTER(T.Lists)
    for name in ("cb200_step_moments_f64", "cb200_step_packed_moments", "cb200_step_vars", "cb200_step_markers"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = vp
    L.cb200_cost_targets.argtypes = [vp, vp, i, C.c_longlong, vp]
    L._step_bound = True
    return L


def cost_targets(prev_cut, prev_cost, n):
    """particle targets of the rank boundaries from last step's cuts and measured costs (host-side; no GPU)"""
    L = _bind(_lib.load(False))
    world = len(prev_cost)
    cut = np.ascontiguousarray(prev_cut, dtype=np.int64)
    cost = np.ascontiguousarray(prev_cost, dtype=np.float64)
    out = np.zeros(world + 1, dtype=np.int32)
    L.cb200_cost_targets(cut.ctypes.data, cost.ctypes.data, world, int(n), out.ctypes.data)
    return out


class Comm:
    """one NCCL communicator per process (cb200_comm_init).  world == 1: no NCCL, handle stays NULL."""

    def __init__(self, L, rank, world, unique_id=None):
        self.L, self.rank, self.world = _bind(L), int(rank), int(world)
        self.handle = None
        if self.world > 1:
            assert unique_id is not None and len(unique_id) == self.L.cb200_comm_id_bytes()
            buf = C.create_string_buffer(bytes(unique_id), len(unique_id))
            self.handle = self.L.cb200_comm_init(self.rank, self.world, buf)

    @staticmethod
    def make_id(L):
        L = _bind(L)
        buf = C.create_string_buffer(L.cb200_comm_id_bytes())
        L.cb200_comm_unique_id(buf)
        return buf.raw

    @classmethod
    def from_env(cls, L, key="cb200_nccl_id"):
        """torchrun-style environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT): rank 0 makes the id and
        publishes it in the launcher's TCP store, the others read it.  cb200_set_device must have been called."""
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if world == 1:
            return cls(L, 0, 1)
        from datetime import timedelta
        from torch.distributed import TCPStore
        agent = os.environ.get("TORCHELASTIC_USE_AGENT_STORE", "False") == "True"
        store = TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ["MASTER_PORT"]), world,
                         is_master=(rank == 0 and not agent), timeout=timedelta(seconds=300), wait_for_workers=False)
        key = f"{key}/{os.environ.get('TORCHELASTIC_RUN_ID', 'run')}/{os.environ.get('TORCHELASTIC_RESTART_COUNT', '0')}"
        if rank == 0:
            store.set(key, cls.make_id(L))
        uid = store.get(key)
        comm = cls(L, rank, world, uid)
        comm._store = store  # rank 0 hosts it outside torchrun: keep it alive
        return comm

    def allreduce(self, values, op="sum", stream=None):
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        if self.world > 1:
            self.L.cb200_comm_allreduce_f64(self.handle, v.ctypes.data, v.size, {"sum": 0, "max": 1, "min": 2}[op], stream)
        return v

    def barrier(self, stream=None):
        if self.world > 1:
            self.L.cb200_comm_barrier(self.handle, stream)

    def destroy(self):
        if self.handle:
            self.L.cb200_comm_destroy(self.handle)
            self.handle = None


class NativeStep:
    """cb200_step_*: the force step of one box shared by comm.world GPUs, from unsorted particles.

    pos (n,3), mass, soft: the WHOLE box in the caller's order (every rank passes the same arrays or only its
    own rows via `rows=(lo, hi)` semantics: only rows [rank*chunk, (rank+1)*chunk) are read).  ewald: None or
    dict(dEwCut, dEwhCut).  rung + active_rung: multistep step."""

    def __init__(self, hc, n, theta=0.7, n_replicas=0, period=1.0, ewald=None, max_bucket=12,
                 root_lo=(-0.5, -0.5, -0.5), root_hi=(0.5, 0.5, 0.5), comm=None, active_rung=0,
                 overlap_ewald=True, cost_cuts=True):
        assert hc.L.cb200_real_bytes() == 4, "the device step is a float pipeline: use the float build"
        self.hc, self.L = hc, _bind(hc.L)
        self.comm = comm
        self.rank = comm.rank if comm else 0
        self.world = comm.world if comm else 1
        self.n = int(n)
        cfg = StepConfig()
        cfg.numParticles, cfg.maxBucket, cfg.nReplicas = self.n, int(max_bucket), int(n_replicas)
        cfg.ewald = 1 if ewald is not None else 0
        cfg.activeRung = int(active_rung)
        cfg.overlapEwald = 1 if overlap_ewald else 0
        cfg.costCuts = 1 if cost_cuts else 0
        cfg.theta, cfg.period = float(theta), float(period)
        cfg.dEwCut = float((ewald or {}).get("dEwCut", 2.6))
        cfg.dEwhCut = float((ewald or {}).get("dEwhCut", 2.8))
        for d in range(3):
            cfg.rootlo[d], cfg.roothi[d] = float(root_lo[d]), float(root_hi[d])
        self.cfg = cfg
        self.handle = self.L.cb200_step_create(comm.handle if comm and comm.handle else None, C.byref(cfg))
        self.chunk = self.L.cb200_step_chunk_rows(self.handle)
        self.capacity = self.L.cb200_step_out_capacity(self.handle)
        self.stream = self.L.cb200_step_stream(self.handle)
        # pinned host buffers, as a host program would hold them
        self.rec = hc.allocatePinnedHostMemory((self.chunk, 5), np.float64)
        self.rec.array[:] = 0
        self.rec.array[:, 4] = 1.0
        self.rung = None
        if active_rung > 0:
            self.rung = hc.allocatePinnedHostMemory((self.chunk,), np.uint8)
            self.rung.array[:] = 0
        self.out = hc.allocatePinnedHostMemory((self.capacity, 5), np.float32)
        self.idx = hc.allocatePinnedHostMemory((self.capacity,), np.int32)
        self.result = StepResult()

    def my_rows(self):
        lo = self.rank * self.chunk
        return lo, min(self.n, lo + self.chunk)

    def set_particles(self, pos, mass, soft, rung=None):
        """fill this rank's slice from whole-box arrays (pad rows keep mass 0 / soft 1 and never reach the tree:
        the gathered array is cut back to n rows)"""
        lo, hi = self.my_rows()
        k = max(0, hi - lo)
        r = self.rec.array
        r[:k, :3] = np.asarray(pos)[lo:hi]
        r[:k, 3] = np.broadcast_to(mass, (self.n,))[lo:hi]
        r[:k, 4] = np.broadcast_to(soft, (self.n,))[lo:hi]
        if rung is not None and self.rung is not None:
            self.rung.array[:k] = np.asarray(rung, dtype=np.uint8)[lo:hi]

    def set_rows(self, pos_rows, mass, soft, rung_rows=None):
        """the same from this rank's OWN rows (my_rows()): what a host that holds only its share passes"""
        lo, hi = self.my_rows()
        k = max(0, hi - lo)
        assert len(pos_rows) == k
        r = self.rec.array
        r[:k, :3] = pos_rows
        r[:k, 3] = mass
        r[:k, 4] = soft
        if rung_rows is not None and self.rung is not None:
            self.rung.array[:k] = np.asarray(rung_rows, dtype=np.uint8)

    def upload(self):
        """records to the device once (then run(resident=True) times the step without the host copies)"""
        d = self.L.cb200_step_device_records(self.handle)
        self.L.cb200_copy_device(d, self.rec.array.ctypes.data, self.rec.array.nbytes, self.stream)
        if self.rung is not None:
            self.L.cb200_copy_device(self.L.cb200_step_device_rungs(self.handle), self.rung.array.ctypes.data,
                                     self.rung.array.nbytes, self.stream)
        self.hc.stream_synchronize(self.stream)

    def run(self, resident=False, keep_lists=False, sfc_order=False):
        """one step.  resident: records already on the device (upload()), results stay there.  Returns the
        StepResult.  world > 1, or sfc_order on one GPU: self.out.array[:rows] holds this rank's rows in SFC (tree)
        order and self.idx.array[:rows] their caller indices -- the form that is copied back slab by slab under the
        list kernels; one GPU without sfc_order: all rows in the caller's order (one scatter + one copy at the end)."""
        res = self.result
        rec = None if resident else self.rec.array.ctypes.data
        rung = None if (resident or self.rung is None) else self.rung.array.ctypes.data
        out = None if resident else self.out.array.ctypes.data
        idx = None if (resident or not (self.world > 1 or sfc_order)) else self.idx.array.ctypes.data
        self.L.cb200_step_run(self.handle, rec, rung, out, idx, self.capacity, 1 if keep_lists else 0, C.byref(res))
        if res.error:
            raise RuntimeError(f"cb200_step_run: error {res.error} (11 node capacity, 2x walk capacity, 30 result buffer)")
        return res

    def phases(self):
        return {name: float(self.result.ms[k]) for k, name in enumerate(PHASES)}

    # -- products of the last run, for tests and parity sampling ------------------------------
    def _download(self, ptr, count, dtype):
        a = np.empty(count, dtype=dtype)
        if count:
            self.L.cb200_copy_device(a.ctypes.data, ptr, a.nbytes, self.stream)
            self.hc.stream_synchronize(self.stream)
        return a

    def tree(self):
        tr = self.L.cb200_step_tree(self.handle).contents
        n, nn, nb = tr.numParticles, tr.numNodes, tr.numBuckets
        i32, f64 = np.int32, np.float64
        out = {"order": self._download(tr.d_order, n, i32), "pos": self._download(tr.d_pos, 3 * n, f64).reshape(n, 3),
               "mass": self._download(tr.d_mass, n, f64), "soft": self._download(tr.d_soft, n, f64),
               "level_start": np.array(tr.levelStart[:tr.numLevels + 1], dtype=np.int32)}
        for name, ptr in (("child0", tr.d_child0), ("child1", tr.d_child1), ("parent", tr.d_parent),
                          ("first", tr.d_first), ("last", tr.d_last), ("bucket_first", tr.d_bucketFirst),
                          ("bucket_count", tr.d_bucketCount)):
            out[name] = self._download(ptr, nn, i32)
        for name, ptr in (("geolo", tr.d_geolo), ("geohi", tr.d_geohi), ("boxlo", tr.d_boxlo), ("boxhi", tr.d_boxhi)):
            out[name] = self._download(ptr, 3 * nn, f64).reshape(nn, 3)
        for name, ptr in (("bucket_node", tr.d_bucketNode), ("bucket_starts", tr.d_bucketStarts),
                          ("bucket_sizes", tr.d_bucketSizes)):
            out[name] = self._download(ptr, nb, i32)
        return out

    def moments(self):
        nn = self.result.numNodes
        return self._download(self.L.cb200_step_moments_f64(self.handle), 27 * nn, np.float64).reshape(nn, 27)

    def vars(self):
        """accumulators of the last run in TREE order (all n rows; rows outside this rank's range are zero)"""
        return self._download(self.L.cb200_step_vars(self.handle), 5 * self.n, np.float32).reshape(self.n, 5)

    def bucket_lists(self, b0, b1):
        """interaction lists of buckets [b0, b1) of the last run (needs run(keep_lists=True)): dict with the
        three flat lists rebased to the range and their markers, plus the buckets' starts / sizes"""
        li = self.L.cb200_step_lists(self.handle).contents
        nb = self.result.numBuckets
        out = {}
        for key, lp, mp in (("cell", li.d_cell, li.d_cellMarkers), ("part", li.d_part, li.d_partMarkers),
                            ("soft", li.d_soft, li.d_softMarkers)):
            m = self._download(mp + 4 * b0, b1 - b0 + 1, np.int32).astype(np.int64)
            lo, hi = int(m[0]), int(m[-1])
            il = self._download(lp + 8 * lo, 2 * (hi - lo), np.int32).reshape(-1, 2)
            out[key] = il
            out[key + "_mark"] = (m - lo).astype(np.int32)
        out["starts"] = self._download(li.d_starts + 4 * b0, b1 - b0, np.int32)
        out["sizes"] = self._download(li.d_sizes + 4 * b0, b1 - b0, np.int32)
        assert li.numBuckets == nb
        return out

    def free(self):
        if self.handle:
            self.L.cb200_step_destroy(self.handle)
            self.handle = None
        for b in (self.rec, self.rung, self.out, self.idx):
            if b is not None:
                self.hc.freePinnedHostMemory(b)
        self.rec = self.rung = self.out = self.idx = None
