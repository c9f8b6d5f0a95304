"""Host-side mirror of the reference's GPU gravity interface (HostCUDA.h:99-126,
EwaldCUDA.h:59-62), bound to libchanga_b200.so through its C ABI.

Function names, argument order and meaning are the reference's; buffers are
numpy arrays over pinned host memory obtained from allocatePinnedHostMemory,
device arrays are opaque integer handles exactly as ChaNGa's DataManager
treats them.  `ForceStep` plays the role of DataManager + TreePiece for one
force evaluation: upload -> cell/particle list requests -> Ewald -> results
back (call sequence of SURVEY.md section 3.1).

Nothing here computes forces on the host; if the CUDA library is missing the
constructor raises.
"""
import ctypes as C
import threading

import numpy as np

from . import lib as _lib

# request-size thresholds of the reference (cuda_typedef.h:18-25): a TreePiece
# flushes its lists to the GPU once this many interactions are pending
NODE_INTERACTIONS_PER_REQUEST = 1_000_000
PART_INTERACTIONS_PER_REQUEST = 1_000_000


def encode_offset(x, y, z, bucket=0):
    """TreePiece.cpp:3631-3644: replica (x,y,z) in [-3,3] -> offsetID"""
    return bucket | (((x + 3) | ((y + 3) << 3) | ((z + 3) << 6)) << 22)


class Pinned:
    """numpy array over memory from allocatePinnedHostMemory"""

    def __init__(self, hc, shape, dtype):
        self.hc = hc
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self.nbytes = n
        p = C.c_void_p()
        hc.L.cb200_allocatePinnedHostMemory(C.byref(p), max(n, 1))
        self.ptr = p.value
        buf = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.hc.L.cb200_freePinnedHostMemory(self.ptr)
            self.ptr = None


class HostCUDA:
    def __init__(self, double=False, device=None):
        self.L = _lib.load(double)
        self.T = self.L.types
        self.double = double
        self.np_real = self.T.np_real
        if device is not None:
            self.L.cb200_set_device(int(device))
        self._fired = {}
        self._lock = threading.Lock()
        self._handler = _lib.CALLBACK_FN(self._on_callback)  # keep alive
        self.L.cb200_set_callback_handler(self._handler)
        self._next_token = 1

    # -- completion tokens (stand-in for heap CkCallback*) --------------------
    def _on_callback(self, token):
        with self._lock:
            self._fired[token] = self._fired.get(token, 0) + 1

    def new_callback(self):
        with self._lock:
            t = self._next_token
            self._next_token += 1
        return t

    def callback_count(self, token):
        with self._lock:
            return self._fired.get(token, 0)

    # -- streams ----------------------------------------------------------------
    def stream_create(self):
        return self.L.cb200_stream_create()

    def stream_destroy(self, s):
        self.L.cb200_stream_destroy(s)

    def stream_synchronize(self, s):
        self.L.cb200_stream_synchronize(s)

    def device_synchronize(self):
        self.L.cb200_device_synchronize()

    # -- the reference surface -----------------------------------------------------
    def allocatePinnedHostMemory(self, shape, dtype):
        return Pinned(self, shape, dtype)

    def freePinnedHostMemory(self, buf):
        buf.free()

    def DataManagerTransferLocalTree(self, moments, compactParts, varParts, stream, numParticles, callback=None):
        """moments (Nn,27), compactParts (Np,5) {mass,soft,x,y,z}, varParts (Np,5): arrays of cudatype.
        Returns (d_localMoments, d_compactParts, d_varParts)."""
        dm, dp, dv = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.L.cb200_DataManagerTransferLocalTree(
            moments.ctypes.data, moments.nbytes, compactParts.ctypes.data, compactParts.nbytes,
            varParts.ctypes.data, varParts.nbytes, C.byref(dm), C.byref(dp), C.byref(dv), stream,
            int(numParticles), callback)
        return dm.value, dp.value, dv.value

    def DataManagerTransferRemoteChunk(self, moments, remoteParts, stream, callback=None):
        dm, dp = C.c_void_p(), C.c_void_p()
        self.L.cb200_DataManagerTransferRemoteChunk(
            moments.ctypes.data, moments.nbytes, remoteParts.ctypes.data, remoteParts.nbytes,
            C.byref(dm), C.byref(dp), stream, callback)
        return dm.value, dp.value

    def TransferParticleVarsBack(self, hostBuffer, d_varParts, stream, cb=None):
        self.L.cb200_TransferParticleVarsBack(hostBuffer.ctypes.data, hostBuffer.nbytes, d_varParts, stream, cb)

    def device_free(self, dptr):
        """what DataManager does with the arrays after the copy-back (DataManager.cpp:992-996)"""
        if dptr:
            self.L.cb200_device_free(dptr)

    def make_request(self, stream, d_localMoments, d_localParts, d_localVars, ilist, markers, starts, sizes,
                     fperiod, cb=None, d_remoteMoments=None, d_remoteParts=None, missedNodes=None,
                     missedParts=None, node=True, remote=False):
        """CudaRequest as ListCompute::send{Node,Part}InteractionsToGpu fills it (Compute.cpp:2064-2260).
        ilist (Li,2) int32 {index, offsetID}; markers (nb+1), starts (nb), sizes (nb) int32."""
        r = self.T.CudaRequest()
        r.stream = stream
        r.d_localMoments, r.d_localParts, r.d_localVars = d_localMoments, d_localParts, d_localVars
        r.d_remoteMoments, r.d_remoteParts = d_remoteMoments, d_remoteParts
        r.list = ilist.ctypes.data
        r.bucketMarkers, r.bucketStarts, r.bucketSizes = markers.ctypes.data, starts.ctypes.data, sizes.ctypes.data
        r.numInteractions = int(markers[-1] - markers[0]) if len(markers) else 0
        r.numBucketsPlusOne = len(markers)
        if missedNodes is not None:
            r.missedNodes, r.sMissed = missedNodes.ctypes.data, missedNodes.nbytes
        if missedParts is not None:
            r.missedParts, r.sMissed = missedParts.ctypes.data, missedParts.nbytes
        r.cb = cb
        r.fperiod = float(fperiod)
        r.node, r.remote = node, remote
        r._keep = (ilist, markers, starts, sizes, missedNodes, missedParts)
        return r

    def TreePieceCellListDataTransferLocal(self, req):
        self.L.cb200_TreePieceCellListDataTransferLocal(C.byref(req))

    def TreePieceCellListDataTransferRemote(self, req):
        self.L.cb200_TreePieceCellListDataTransferRemote(C.byref(req))

    def TreePieceCellListDataTransferRemoteResume(self, req):
        self.L.cb200_TreePieceCellListDataTransferRemoteResume(C.byref(req))

    def TreePiecePartListDataTransferLocal(self, req):
        self.L.cb200_TreePiecePartListDataTransferLocal(C.byref(req))

    def TreePiecePartListDataTransferLocalSmallPhase(self, req, parts):
        self.L.cb200_TreePiecePartListDataTransferLocalSmallPhase(C.byref(req), parts.ctypes.data, len(parts))

    def TreePiecePartListDataTransferRemote(self, req):
        self.L.cb200_TreePiecePartListDataTransferRemote(C.byref(req))

    def TreePiecePartListDataTransferRemoteResume(self, req):
        self.L.cb200_TreePiecePartListDataTransferRemoteResume(C.byref(req))

    # -- Ewald ----------------------------------------------------------------------
    def EwaldHostMemorySetup(self, nParticles, nEwhLoop, largephase):
        e = self.T.EwaldData()
        self.L.cb200_EwaldHostMemorySetup(C.byref(e), int(nParticles), int(nEwhLoop), int(largephase))
        return e

    def EwaldHostMemoryFree(self, e, largephase):
        self.L.cb200_EwaldHostMemoryFree(C.byref(e), int(largephase))

    def EwaldHost(self, d_localParts, d_localVars, h_idata, stream, cb=None, myIndex=0, largephase=1):
        self.L.cb200_EwaldHost(d_localParts, d_localVars, C.byref(h_idata), stream, cb, myIndex, int(largephase))

    def fill_ewald(self, e, root_cell, momc, ewt, L, fEwCut, nReps, active=None, first=0, last=None,
                   fInner2coef=1.1e-2):
        """what TreePiece::EwaldGPU writes into the pinned buffers (Ewald.cpp:387-517).
        root_cell: 27 values (mass, cm used); momc: 32 complete root moments; ewt (nh,5)."""
        ro = e.cachedData.contents
        ro.mm.totalMass, ro.mm.cmx, ro.mm.cmy, ro.mm.cmz = (float(root_cell[2]), float(root_cell[3]),
                                                          float(root_cell[4]), float(root_cell[5]))
        for (name, _), v in zip(self.T.MomcData._fields_, momc):
            setattr(ro.momcRoot, name, float(v))
        nh = len(ewt)
        for i in range(nh):
            e.ewt[i].hx, e.ewt[i].hy, e.ewt[i].hz, e.ewt[i].hCfac, e.ewt[i].hSfac = (float(x) for x in ewt[i])
        alpha = 2.0 / L
        if active is not None:
            n = len(active)
            C.memmove(e.EwaldMarkers, np.ascontiguousarray(active, dtype=np.int32).ctypes.data, 4 * n)
            e.EwaldRange[0], e.EwaldRange[1] = 0, n - 1
        else:
            n = last - first + 1
            e.EwaldRange[0], e.EwaldRange[1] = first, last
        ro.n, ro.nReps, ro.nEwReps, ro.nEwhLoop = n, int(nReps), int(np.ceil(fEwCut)), nh
        ro.L, ro.fEwCut, ro.alpha, ro.alpha2 = L, fEwCut, alpha, alpha * alpha
        ro.k1 = np.pi / (alpha * alpha * L * L * L)
        ro.ka = 2.0 * alpha / np.sqrt(np.pi)
        ro.fEwCut2 = fEwCut * fEwCut * L * L
        ro.fInner2 = fInner2coef * L * L
        return e

    # -- timing taps ----------------------------------------------------------------
    def timing(self, on=True):
        self.L.cb200_timing_reset()
        self.L.cb200_timing_enable(1 if on else 0)

    def timing_read(self):
        out = (C.c_double * 6)()
        self.L.cb200_timing_read(C.byref(out))
        return {"cell_ms": out[0], "part_ms": out[1], "ewald_ms": out[2],
                "cell_launches": int(out[3]), "part_launches": int(out[4]), "ewald_launches": int(out[5])}

    def kernel_launches(self):
        return int(self.L.cb200_kernel_launches())


def split_requests(markers, max_interactions):
    """bucket ranges [b0,b1) whose lists hold about max_interactions entries --
    where ListCompute::stateReady flushes (Compute.cpp:1737-1741)."""
    nb = len(markers) - 1
    out, b0 = [], 0
    while b0 < nb:
        b1 = int(np.searchsorted(markers, markers[b0] + max_interactions, side="left"))
        b1 = min(max(b1, b0 + 1), nb)
        out.append((b0, b1))
        b0 = b1
    return out


class StagedLists:
    """One kind of interaction list (cell or particle) of one TreePiece, cut into
    requests and copied into pinned buffers the way GenericList<T>::serialize
    leaves them (Compute.cpp:1034-1218): filled buckets only, markers rebased."""

    def __init__(self, hc, ilist, markers, starts, sizes, max_interactions):
        self.hc = hc
        self.chunks = []
        markers = np.asarray(markers, dtype=np.int64)
        for b0, b1 in split_requests(markers, max_interactions):
            lo, hi = int(markers[b0]), int(markers[b1])
            pl = hc.allocatePinnedHostMemory((max(hi - lo, 1), 2), np.int32)
            pm = hc.allocatePinnedHostMemory((b1 - b0 + 1,), np.int32)
            ps = hc.allocatePinnedHostMemory((b1 - b0,), np.int32)
            pz = hc.allocatePinnedHostMemory((b1 - b0,), np.int32)
            pl.array[: hi - lo] = ilist[lo:hi]
            pm.array[:] = markers[b0:b1 + 1] - lo
            ps.array[:] = starts[b0:b1]
            pz.array[:] = sizes[b0:b1]
            self.chunks.append((pl, pm, ps, pz, hi - lo))
        self.num_interactions = int(markers[-1] - markers[0]) if len(markers) else 0

    @property
    def h2d_bytes(self):
        return sum(n * 8 + pm.nbytes + ps.nbytes + pz.nbytes for (_, pm, ps, pz, n) in self.chunks)

    def free(self):
        for pl, pm, ps, pz, _ in self.chunks:
            for b in (pl, pm, ps, pz):
                b.free()
        self.chunks = []


class ForceStep:
    """DataManager + TreePiece for one force evaluation of one workload (a dict, see
    changa_b200.workloads): every call below is one the reference's host code makes."""

    def __init__(self, hc, wl):
        """All requests of the step go on ONE stream, as every request of a TreePiece does in the
        reference (TreePiece.cpp:5380): the list kernels accumulate into the particle rows with plain
        read-modify-write, so requests that touch the same buckets must be ordered."""
        self.hc, self.wl = hc, wl
        rt = hc.np_real
        self.np_ = len(wl["parts"])
        self.moments = hc.allocatePinnedHostMemory(wl["moments"].shape, rt)
        self.parts = hc.allocatePinnedHostMemory(wl["parts"].shape, rt)
        self.vars_in = hc.allocatePinnedHostMemory((self.np_, 5), rt)
        self.vars_out = hc.allocatePinnedHostMemory((self.np_, 5), rt)
        self.moments.array[:] = wl["moments"]
        self.parts.array[:] = wl["parts"]
        self.vars_in.array[:] = 0
        self.cell = StagedLists(hc, *wl["cell"], NODE_INTERACTIONS_PER_REQUEST) if wl.get("cell") else None
        self.part = StagedLists(hc, *wl["part"], PART_INTERACTIONS_PER_REQUEST) if wl.get("part") else None
        # softened cells (Compute.cpp:1683-1699): a p-p request whose sources travel with it
        self.soft, self.soft_src = None, None
        if wl.get("softcell"):
            self.soft = StagedLists(hc, *wl["softcell"][:4], PART_INTERACTIONS_PER_REQUEST)
            src = wl["softcell"][4]
            self.soft_src = hc.allocatePinnedHostMemory(src.shape, rt)
            self.soft_src.array[:] = src
        self.streams = [hc.stream_create()]
        self.ewald = None
        ew = wl.get("ewald")
        if ew:
            active = ew.get("active")
            n = len(active) if active is not None else self.np_
            self.ewald = hc.EwaldHostMemorySetup(n, len(ew["ewt"]), 1)
            if active is None:
                active = np.arange(self.np_, dtype=np.int32)
            hc.fill_ewald(self.ewald, ew["root"], ew["momc"], ew["ewt"], ew["L"], ew["fEwCut"], ew["nReps"],
                          active=active, fInner2coef=ew.get("fInner2coef", 1.1e-2))
        self.fperiod = float(wl.get("fperiod", 0.0))

    @property
    def h2d_bytes(self):
        n = self.moments.nbytes + self.parts.nbytes
        for s in (self.cell, self.part, self.soft):
            if s:
                n += s.h2d_bytes
        if self.soft_src is not None:
            n += self.soft_src.nbytes * len(self.soft.chunks)
        if self.ewald:
            n += 4 * self.ewald.cachedData.contents.n
        return n

    @property
    def d2h_bytes(self):
        return self.vars_out.nbytes

    def _requests(self):
        """CudaRequests of the step, built once (a TreePiece builds them in C++ as its lists fill up);
        only the device handles change from step to step"""
        if getattr(self, "_reqs", None) is None:
            hc, reqs, k = self.hc, [], 0
            for staged, call, node, extra in (
                    (self.cell, hc.TreePieceCellListDataTransferLocal, True, None),
                    (self.part, hc.TreePiecePartListDataTransferLocal, False, None),
                    (self.soft, hc.TreePiecePartListDataTransferLocalSmallPhase, False, self.soft_src)):
                if not staged:
                    continue
                for pl, pm, ps, pz, n in staged.chunks:
                    st = self.streams[k % len(self.streams)]
                    k += 1
                    req = hc.make_request(st, None, None, None, pl.array, pm.array, ps.array, pz.array,
                                          self.fperiod, node=node)
                    reqs.append((req, call, extra))
            self._reqs = reqs
        return self._reqs

    def run(self, sync=True):
        """enqueue the whole step; returns the (pinned) result array after the copy-back.
        Order on the stream: upload, Ewald (needs only the particles, so its kernel runs while the
        first lists are still crossing PCIe), cell lists, particle lists, softened cells, copy back."""
        hc = self.hc
        s0 = self.streams[0]
        dm, dp, dv = hc.DataManagerTransferLocalTree(self.moments.array, self.parts.array, self.vars_in.array,
                                                     s0, self.np_)
        if len(self.streams) > 1:
            hc.stream_synchronize(s0)  # DataManager waits for the upload callback before the walks start
        if self.ewald:
            hc.EwaldHost(dp, dv, self.ewald, s0)
        for req, call, extra in self._requests():
            req.d_localMoments, req.d_localParts, req.d_localVars = dm, dp, dv
            if extra is not None:
                call(req, extra.array)
            else:
                call(req)
        if len(self.streams) > 1:
            for st in self.streams[1:]:
                hc.stream_synchronize(st)
        hc.TransferParticleVarsBack(self.vars_out.array, dv, s0)
        if sync:
            hc.stream_synchronize(s0)
        for p in (dm, dp, dv):
            hc.device_free(p)
        return self.vars_out.array

    def free(self):
        for b in (self.moments, self.parts, self.vars_in, self.vars_out):
            b.free()
        for s in (self.cell, self.part, self.soft):
            if s:
                s.free()
        if self.soft_src is not None:
            self.soft_src.free()
        if self.ewald:
            self.hc.EwaldHostMemoryFree(self.ewald, 1)
        for s in self.streams:
            self.hc.stream_destroy(s)


class ShardedForceStep(ForceStep):
    """ForceStep of ONE rank of a multi-GPU step (one process per GPU): instead of every
    rank pushing the whole particle and moment arrays through its own PCIe link
    (DataManagerTransferLocalTree), a rank uploads only ITS slice of the two arrays, packs it, and one
    all-gather each over NVLink (torch.distributed / NCCL) replicates the packed arrays; the list
    requests are the reference's (TreePiece*ListDataTransferLocal, EwaldHost) against those arrays, and
    only the rank's own rows of the accelerations come back.  Same kernels, same packed inputs: the rows
    equal ForceStep's bit for bit (bench.py checks it once per run)."""

    def __init__(self, hc, wl, torch, dist, rank, world):
        super().__init__(hc, wl)
        from .multigpu import shard_rows
        self.torch, self.dist = torch, dist
        rt, L = hc.np_real, hc.L
        mine_p, self.pc = shard_rows(np.ascontiguousarray(wl["parts"], dtype=rt), rank, world)
        mine_m, self.mc = shard_rows(np.ascontiguousarray(wl["moments"], dtype=rt), rank, world)
        for b in (self.moments, self.parts):  # the full-size pinned copies are not used
            b.free()
        self.parts = hc.allocatePinnedHostMemory(mine_p.shape, rt)
        self.moments = hc.allocatePinnedHostMemory(mine_m.shape, rt)
        self.parts.array[:] = mine_p
        self.moments.array[:] = mine_m
        tdt = torch.float64 if np.dtype(rt) == np.float64 else torch.float32
        dev = torch.device("cuda", torch.cuda.current_device())
        pb, mb = L.cb200_packed_particle_bytes(), L.cb200_packed_moment_bytes()
        self.d_parts = torch.empty(mine_p.shape, dtype=tdt, device=dev)
        self.d_mom = torch.empty(mine_m.shape, dtype=tdt, device=dev)
        self.send_p = torch.empty(self.pc * pb, dtype=torch.uint8, device=dev)
        self.send_m = torch.empty(self.mc * mb, dtype=torch.uint8, device=dev)
        self.pk_parts = torch.empty(self.pc * world * pb, dtype=torch.uint8, device=dev)
        self.pk_mom = torch.empty(self.mc * world * mb, dtype=torch.uint8, device=dev)
        self.d_vars = torch.empty((self.np_, 5), dtype=tdt, device=dev)
        self.ext = torch.cuda.ExternalStream(self.streams[0])
        # my rows: the particles of my buckets (contiguous in tree order)
        lo, hi = [], []
        for key in ("cell", "part", "softcell"):
            if wl.get(key) and len(wl[key][2]):
                st, sz = np.asarray(wl[key][2]), np.asarray(wl[key][3])
                lo.append(int(st.min())); hi.append(int((st + sz).max()))
        ew = wl.get("ewald")
        if ew and ew.get("active") is not None and len(ew["active"]):
            lo.append(int(np.min(ew["active"]))); hi.append(int(np.max(ew["active"])) + 1)
        self.p0, self.p1 = (min(lo), max(hi)) if lo else (0, self.np_)
        self.row_bytes = 5 * np.dtype(rt).itemsize

    @property
    def d2h_bytes(self):
        return (self.p1 - self.p0) * self.row_bytes

    def run(self, sync=True):
        hc, L, s0 = self.hc, self.hc.L, self.streams[0]
        L.cb200_copy_device(self.d_parts.data_ptr(), self.parts.ptr, self.parts.nbytes, s0)
        L.cb200_copy_device(self.d_mom.data_ptr(), self.moments.ptr, self.moments.nbytes, s0)
        L.cb200_pack_particles_device(self.d_parts.data_ptr(), self.send_p.data_ptr(), self.pc, s0)
        L.cb200_pack_moments_device(self.d_mom.data_ptr(), self.send_m.data_ptr(), self.mc, s0)
        with self.torch.cuda.stream(self.ext):
            self.dist.all_gather_into_tensor(self.pk_parts, self.send_p)
            self.dist.all_gather_into_tensor(self.pk_mom, self.send_m)
        dm, dp, dv = self.pk_mom.data_ptr(), self.pk_parts.data_ptr(), self.d_vars.data_ptr()
        L.cb200_zero_vars_device(dv, self.np_, s0)
        if self.ewald:
            hc.EwaldHost(dp, dv, self.ewald, s0)
        for req, call, extra in self._requests():
            req.d_localMoments, req.d_localParts, req.d_localVars = dm, dp, dv
            if extra is not None:
                call(req, extra.array)
            else:
                call(req)
        off = self.p0 * self.row_bytes
        L.cb200_TransferParticleVarsBack(self.vars_out.ptr + off, (self.p1 - self.p0) * self.row_bytes, dv + off, s0, None)
        if sync:
            hc.stream_synchronize(s0)
        return self.vars_out.array

