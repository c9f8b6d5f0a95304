"""ctypes view of libchanga_b200.so -- the C ABI declared in include/changa_b200_api.h.

The records (CudaRequest, EwaldData ...) mirror include/changa_b200_types.h,
which in turn is byte-compatible with the reference's HostCUDA.h:31-97 and
EwaldCUDA.h:11-57.  Loading fails loudly when the library has not been built:
there is no Python or CPU fallback for any entry point.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NEWH = 80

_LIBS = {}


class LibraryMissing(RuntimeError):
    pass


def _real(double):
    return C.c_double if double else C.c_float


def make_types(double=False):
    """ctypes Structures for one precision (cudatype = float | double)."""
    real = _real(double)

    class CudaRequest(C.Structure):
        _fields_ = [
            ("stream", C.c_void_p),
            ("d_localMoments", C.c_void_p), ("d_remoteMoments", C.c_void_p),
            ("d_localParts", C.c_void_p), ("d_remoteParts", C.c_void_p),
            ("d_localVars", C.c_void_p),
            ("sMoments", C.c_size_t), ("sCompactParts", C.c_size_t), ("sVarParts", C.c_size_t),
            ("list", C.c_void_p),
            ("bucketMarkers", C.c_void_p), ("bucketStarts", C.c_void_p), ("bucketSizes", C.c_void_p),
            ("numInteractions", C.c_int), ("numBucketsPlusOne", C.c_int),
            ("tp", C.c_void_p),
            ("missedNodes", C.c_void_p), ("missedParts", C.c_void_p), ("sMissed", C.c_size_t),
            ("affectedBuckets", C.c_void_p), ("cb", C.c_void_p), ("state", C.c_void_p),
            ("fperiod", real),
            ("node", C.c_bool), ("remote", C.c_bool),
        ]

    class EwtData(C.Structure):
        _fields_ = [(n, real) for n in ("hx", "hy", "hz", "hCfac", "hSfac")]

    class MultipoleMomentsData(C.Structure):
        _fields_ = [(n, real) for n in ("totalMass", "cmx", "cmy", "cmz")]

    class MomcData(C.Structure):
        _fields_ = [(n, real) for n in (
            "m", "xx", "yy", "xy", "xz", "yz",
            "xxx", "xyy", "xxy", "yyy", "xxz", "yyz", "xyz",
            "xxxx", "xyyy", "xxxy", "yyyy", "xxxz", "yyyz", "xxyy", "xxyz", "xyyz",
            "zz", "xzz", "yzz", "zzz", "xxzz", "xyzz", "xzzz", "yyzz", "yzzz", "zzzz")]

    class EwaldReadOnlyData(C.Structure):
        _fields_ = [("mm", MultipoleMomentsData), ("momcRoot", MomcData),
                    ("n", C.c_int), ("nReps", C.c_int), ("nEwReps", C.c_int), ("nEwhLoop", C.c_int)] + \
                   [(n, real) for n in ("L", "fEwCut", "alpha", "alpha2", "k1", "ka", "fEwCut2", "fInner2")]

    class EwaldData(C.Structure):
        _fields_ = [("EwaldRange", C.c_int * 2), ("EwaldMarkers", C.POINTER(C.c_int)),
                    ("ewt", C.POINTER(EwtData)), ("cachedData", C.POINTER(EwaldReadOnlyData))]

    class CudaDevPtr(C.Structure):
        _fields_ = [("d_list", C.c_void_p), ("d_bucketMarkers", C.c_void_p),
                    ("d_bucketStarts", C.c_void_p), ("d_bucketSizes", C.c_void_p)]

    class Lists(C.Structure):
        _fields_ = [("d_cell", C.c_void_p), ("d_soft", C.c_void_p), ("d_part", C.c_void_p),
                    ("d_cellMarkers", C.c_void_p), ("d_softMarkers", C.c_void_p), ("d_partMarkers", C.c_void_p),
                    ("d_starts", C.c_void_p), ("d_sizes", C.c_void_p), ("d_nodeParticles", C.c_void_p),
                    ("nCell", C.c_longlong), ("nSoft", C.c_longlong), ("nPart", C.c_longlong),
                    ("numBuckets", C.c_int), ("error", C.c_int)]

    class DevTree(C.Structure):
        _fields_ = [(n, C.c_void_p) for n in (
            "d_pos", "d_mass", "d_soft", "d_packedParts", "d_order", "d_child0", "d_child1", "d_parent",
            "d_first", "d_last", "d_geolo", "d_geohi", "d_boxlo", "d_boxhi", "d_bucketNode",
            "d_bucketFirst", "d_bucketCount", "d_bucketStarts", "d_bucketSizes")] + [
            ("numParticles", C.c_int), ("numNodes", C.c_int), ("numBuckets", C.c_int), ("numLevels", C.c_int),
            ("levelStart", C.c_int * 66), ("error", C.c_int)]

    class T:
        pass

    T.Lists = Lists
    T.DevTree = DevTree

    T.real = real
    T.np_real = np.float64 if double else np.float32
    T.CudaRequest, T.EwtData, T.MomcData = CudaRequest, EwtData, MomcData
    T.MultipoleMomentsData, T.EwaldReadOnlyData = MultipoleMomentsData, EwaldReadOnlyData
    T.EwaldData, T.CudaDevPtr = EwaldData, CudaDevPtr
    if not double:
        assert C.sizeof(CudaRequest) == 176 and C.sizeof(EwaldReadOnlyData) == 192
    assert C.sizeof(EwaldData) == 32 and C.sizeof(CudaDevPtr) == 32
    return T


# every symbol include/changa_b200_api.h PART 2 declares (tests check the export list)
C_ABI_SYMBOLS = [
    "cb200_abi_version", "cb200_real_bytes", "cb200_build_info", "cb200_set_callback_handler",
    "cb200_stream_create", "cb200_stream_destroy", "cb200_stream_synchronize",
    "cb200_device_synchronize", "cb200_set_device", "cb200_device_free",
    "cb200_allocatePinnedHostMemory", "cb200_freePinnedHostMemory",
    "cb200_DataManagerTransferLocalTree", "cb200_DataManagerTransferRemoteChunk",
    "cb200_TransferParticleVarsBack",
    "cb200_TreePieceCellListDataTransferLocal", "cb200_TreePieceCellListDataTransferRemote",
    "cb200_TreePieceCellListDataTransferRemoteResume",
    "cb200_TreePiecePartListDataTransferLocal", "cb200_TreePiecePartListDataTransferLocalSmallPhase",
    "cb200_TreePiecePartListDataTransferRemote", "cb200_TreePiecePartListDataTransferRemoteResume",
    "cb200_EwaldHostMemorySetup", "cb200_EwaldHostMemoryFree", "cb200_EwaldHost",
    "cb200_cell_list_device", "cb200_part_list_device", "cb200_cell_list_device_ex",
    "cb200_part_list_device_ex", "cb200_ewald_device",
    "cb200_packed_moment_bytes", "cb200_packed_particle_bytes",
    "cb200_pack_moments_device", "cb200_pack_particles_device", "cb200_zero_vars_device", "cb200_copy_device",
    "cb200_timing_enable", "cb200_timing_reset", "cb200_timing_read", "cb200_kernel_launches",
    "cb200_build_moments", "cb200_partition_buckets", "cb200_walk_device", "cb200_lists_free",
    "cb200_walk_device_active", "cb200_active_sets_device",
    "cb200_build_tree", "cb200_tree_free",
    "cb200_comm_id_bytes", "cb200_comm_unique_id", "cb200_comm_init", "cb200_comm_destroy", "cb200_comm_rank",
    "cb200_comm_world", "cb200_comm_nccl_version", "cb200_comm_allreduce_f64", "cb200_comm_barrier",
    "cb200_comm_allgather", "cb200_cost_targets",
    "cb200_step_create", "cb200_step_destroy", "cb200_step_chunk_rows", "cb200_step_out_capacity",
    "cb200_step_stream", "cb200_step_device_records", "cb200_step_device_rungs", "cb200_step_run",
    "cb200_step_tree", "cb200_step_lists", "cb200_step_moments_f64", "cb200_step_packed_moments",
    "cb200_step_vars", "cb200_step_markers",
]

CALLBACK_FN = C.CFUNCTYPE(None, C.c_void_p)


def library_path(double=False):
    """in-tree build; CB200_LIB / CB200_LIB_F64 point at another build of the same ABI (A/B runs)"""
    override = os.environ.get("CB200_LIB_F64" if double else "CB200_LIB")
    if override:
        return os.path.abspath(override)
    return os.path.join(HERE, "libchanga_b200_f64.so" if double else "libchanga_b200.so")


def load(double=False):
    """Load (once) and type the C ABI.  Raises LibraryMissing if not built."""
    if double in _LIBS:
        return _LIBS[double]
    path = library_path(double)
    if not os.path.exists(path):
        raise LibraryMissing(
            f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    T = make_types(double)
    vp, i, sz, real = C.c_void_p, C.c_int, C.c_size_t, T.real
    pvp = C.POINTER(C.c_void_p)
    req = C.POINTER(T.CudaRequest)
    L.cb200_abi_version.restype = i
    L.cb200_real_bytes.restype = i
    L.cb200_build_info.restype = C.c_char_p
    L.cb200_set_callback_handler.argtypes = [CALLBACK_FN]
    L.cb200_stream_create.restype = vp
    L.cb200_stream_destroy.argtypes = [vp]
    L.cb200_stream_synchronize.argtypes = [vp]
    L.cb200_set_device.argtypes = [i]
    L.cb200_device_free.argtypes = [vp]
    L.cb200_allocatePinnedHostMemory.argtypes = [pvp, sz]
    L.cb200_freePinnedHostMemory.argtypes = [vp]
    L.cb200_DataManagerTransferLocalTree.argtypes = [vp, sz, vp, sz, vp, sz, pvp, pvp, pvp, vp, i, vp]
    L.cb200_DataManagerTransferRemoteChunk.argtypes = [vp, sz, vp, sz, pvp, pvp, vp, vp]
    L.cb200_TransferParticleVarsBack.argtypes = [vp, sz, vp, vp, vp]
    for name in ("cb200_TreePieceCellListDataTransferLocal", "cb200_TreePieceCellListDataTransferRemote",
                 "cb200_TreePieceCellListDataTransferRemoteResume", "cb200_TreePiecePartListDataTransferLocal",
                 "cb200_TreePiecePartListDataTransferRemote", "cb200_TreePiecePartListDataTransferRemoteResume"):
        getattr(L, name).argtypes = [req]
    L.cb200_TreePiecePartListDataTransferLocalSmallPhase.argtypes = [req, vp, i]
    ew = C.POINTER(T.EwaldData)
    L.cb200_EwaldHostMemorySetup.argtypes = [ew, i, i, i]
    L.cb200_EwaldHostMemoryFree.argtypes = [ew, i]
    L.cb200_EwaldHost.argtypes = [vp, vp, ew, vp, vp, i, i]
    L.cb200_cell_list_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, real, vp]
    L.cb200_part_list_device.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, real, vp]
    L.cb200_cell_list_device_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, real, i, vp]
    L.cb200_part_list_device_ex.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, real, i, vp]
    L.cb200_ewald_device.argtypes = [vp, vp, vp, i, C.POINTER(T.EwaldReadOnlyData), C.POINTER(T.EwtData), vp]
    L.cb200_packed_moment_bytes.restype = sz
    L.cb200_packed_particle_bytes.restype = sz
    L.cb200_pack_moments_device.argtypes = [vp, vp, i, vp]
    L.cb200_pack_particles_device.argtypes = [vp, vp, i, vp]
    L.cb200_zero_vars_device.argtypes = [vp, i, vp]
    L.cb200_copy_device.argtypes = [vp, vp, sz, vp]
    L.cb200_timing_enable.argtypes = [i]
    L.cb200_timing_read.argtypes = [C.POINTER(C.c_double * 6)]
    L.cb200_kernel_launches.restype = C.c_longlong
    L.cb200_build_moments.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, vp, vp, vp]
    L.cb200_partition_buckets.argtypes = [vp, i, i, vp]
    L.cb200_walk_device.argtypes = [i, i, i, vp] + [vp] * 11 + [C.c_double, i, C.c_double, i, i,
                                                                 C.POINTER(T.Lists), vp]
    L.cb200_walk_device_active.argtypes = [i, i, i, vp] + [vp] * 11 + [C.c_double, i, C.c_double, i, i, vp,
                                                                        C.POINTER(T.Lists), vp]
    L.cb200_active_sets_device.argtypes = [vp, vp, i, vp, vp, i, i, vp, vp, vp, vp]
    L.cb200_lists_free.argtypes = [C.POINTER(T.Lists), vp]
    L.cb200_build_tree.argtypes = [vp, vp, vp, i, i, vp, vp, C.POINTER(T.DevTree), vp]
    L.cb200_tree_free.argtypes = [C.POINTER(T.DevTree), vp]
    L.types = T
    L.path = path
    _LIBS[double] = L
    return L


def partition_buckets(cost, n_ranks, double=False):
    """cuts[0..n_ranks] over SFC-ordered buckets (host-side; no GPU needed)."""
    L = load(double)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    cuts = np.zeros(n_ranks + 1, dtype=np.int32)
    L.cb200_partition_buckets(cost.ctypes.data, len(cost), n_ranks, cuts.ctypes.data)
    return cuts
