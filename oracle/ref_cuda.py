"""The reference's own GPU path -- /root/reference/HostCUDA.cu + CUDAMoments.cu compiled
UNMODIFIED for sm_100a into oracle/_ref/libhostcuda_ref.so (oracle/Makefile) -- driven through
the same call sequence as changa_b200.hostcuda.ForceStep.  TEST INFRASTRUCTURE ONLY: a second
parity target ("the reference's CUDA on this box") and the kernel-time baseline the product has
to beat.  Needs a GPU; absent where oracle/_ref was never built."""
import ctypes as C
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libhostcuda_ref.so")


def available():
    return os.path.exists(PATH)


class RefCuda:
    def __init__(self):
        from changa_b200 import lib as _lib
        self.T = _lib.make_types(False)
        L = C.CDLL(PATH)
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        pvp = C.POINTER(C.c_void_p)
        L.refshim_stream_create.restype = vp
        L.refshim_stream_sync.argtypes = [vp]
        L.refshim_free.argtypes = [vp]
        L.ref_allocatePinnedHostMemory.argtypes = [pvp, sz]
        L.ref_freePinnedHostMemory.argtypes = [vp]
        L.ref_DataManagerTransferLocalTree.argtypes = [vp, sz, vp, sz, vp, sz, pvp, pvp, pvp, vp, i, vp]
        L.ref_TransferParticleVarsBack.argtypes = [vp, sz, vp, vp, vp]
        req = C.POINTER(self.T.CudaRequest)
        L.ref_TreePieceCellListDataTransferLocal.argtypes = [req]
        L.ref_TreePiecePartListDataTransferLocal.argtypes = [req]
        ew = C.POINTER(self.T.EwaldData)
        L.ref_EwaldHostMemorySetup.argtypes = [ew, i, i, i]
        L.ref_EwaldHostMemoryFree.argtypes = [ew, i]
        L.ref_EwaldHost.argtypes = [vp, vp, ew, vp, vp, i, i]
        self.L = L
        self.stream = L.refshim_stream_create()

    def pinned(self, arr, dtype):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        p = C.c_void_p()
        self.L.ref_allocatePinnedHostMemory(C.byref(p), max(arr.nbytes, 1))
        buf = (C.c_char * max(arr.nbytes, 1)).from_address(p.value)
        out = np.frombuffer(buf, dtype=dtype, count=arr.size).reshape(arr.shape)
        out[...] = arr
        return out, p.value

    def force_step(self, wl, repeats=1, max_interactions=1_000_000):
        """returns (vars (N,5) float32, seconds per step: upload -> lists -> Ewald -> copy back)"""
        from changa_b200.hostcuda import split_requests
        L, T, s = self.L, self.T, self.stream
        n = len(wl["parts"])
        mom, _ = self.pinned(wl["moments"], np.float32)
        par, _ = self.pinned(wl["parts"], np.float32)
        var, _ = self.pinned(np.zeros((n, 5)), np.float32)
        out, _ = self.pinned(np.zeros((n, 5)), np.float32)
        staged = []
        for key in ("cell", "part"):
            if not wl.get(key):
                continue
            il, m, st, sz = wl[key]
            m64 = np.asarray(m, dtype=np.int64)
            for b0, b1 in split_requests(m64, max_interactions):
                lo, hi = int(m64[b0]), int(m64[b1])
                staged.append((key, self.pinned(il[lo:hi], np.int32)[0], self.pinned(m64[b0:b1 + 1] - lo, np.int32)[0],
                               self.pinned(st[b0:b1], np.int32)[0], self.pinned(sz[b0:b1], np.int32)[0]))
        ew = wl.get("ewald")
        e = None
        if ew:
            from changa_b200.hostcuda import HostCUDA
            act = ew["active"] if ew["active"] is not None else np.arange(n, dtype=np.int32)
            e = T.EwaldData()
            L.ref_EwaldHostMemorySetup(C.byref(e), len(act), len(ew["ewt"]), 1)
            HostCUDA.fill_ewald(type("X", (), {"T": T})(), e, ew["root"], ew["momc"], ew["ewt"], ew["L"], ew["fEwCut"],
                                ew["nReps"], active=act)
        best = None
        for _ in range(repeats):
            L.refshim_stream_sync(s)
            t0 = time.perf_counter()
            dm, dp, dv = C.c_void_p(), C.c_void_p(), C.c_void_p()
            L.ref_DataManagerTransferLocalTree(mom.ctypes.data, mom.nbytes, par.ctypes.data, par.nbytes,
                                               var.ctypes.data, var.nbytes, C.byref(dm), C.byref(dp), C.byref(dv),
                                               s, n, None)
            for key, il, m, st, sz in staged:
                r = T.CudaRequest()
                r.stream = s
                r.d_localMoments, r.d_localParts, r.d_localVars = dm, dp, dv
                r.sMoments, r.sCompactParts, r.sVarParts = mom.nbytes, par.nbytes, var.nbytes
                r.list = il.ctypes.data
                r.bucketMarkers, r.bucketStarts, r.bucketSizes = m.ctypes.data, st.ctypes.data, sz.ctypes.data
                r.numInteractions, r.numBucketsPlusOne = int(m[-1]), len(m)
                r.fperiod = float(wl["fperiod"])
                r.node = key == "cell"
                (L.ref_TreePieceCellListDataTransferLocal if key == "cell" else L.ref_TreePiecePartListDataTransferLocal)(C.byref(r))
            if e is not None:
                L.ref_EwaldHost(dp, dv, C.byref(e), s, None, 0, 1)
            L.ref_TransferParticleVarsBack(out.ctypes.data, out.nbytes, dv, s, None)
            L.refshim_stream_sync(s)
            for p in (dm, dp, dv):
                L.refshim_free(p)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        if e is not None:
            L.ref_EwaldHostMemoryFree(C.byref(e), 1)
        return out.copy(), best
