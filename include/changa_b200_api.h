/* changa_b200_api.h -- entry points of the B200 gravity library.
 *
 * PART 1 declares, with C++ linkage and the reference's exact signatures, the
 * free functions ChaNGa's Charm++ host code calls (so TreePiece / DataManager
 * / Compute / Ewald link against this library unchanged).  Each one names the
 * reference declaration it replaces and its call sites.
 *
 * PART 2 is the same surface as a plain C ABI (`extern "C"`, cb200_ prefix,
 * pointers and sizes only) for FFI users (ctypes, cgo, JNI ...), plus the
 * pieces that have no reference analogue: device moment build, SFC bucket
 * partitioner, library/runtime introspection.
 *
 * Error convention (reference: HostCUDA.cu:39-47): every function returns
 * void; a CUDA failure prints "Fatal CUDA Error ..." to stderr and abort()s.
 * Completion is signalled through hapiAddCallback(stream, cb) exactly where
 * the reference does it; standalone builds route that to
 * cb200_set_callback_handler().
 */
#ifndef CHANGA_B200_API_H
#define CHANGA_B200_API_H

#include "changa_b200_types.h"

/* ======================= PART 1: reference-compatible ==================== */
#ifdef __cplusplus

/* HostCUDA.h:99-100; callers Compute.cpp:1061-1064,1153-1156,2045-2054,
 * DataManager.cpp:513-515,847-848,970.  size==0 asserts, like the reference. */
void allocatePinnedHostMemory(void **ptr, size_t size);
void freePinnedHostMemory(void *ptr);

/* HostCUDA.h:102-107; caller DataManager.cpp:903.  Uploads the node-level
 * tree, returns three device arrays the CALLER later cudaFree()s
 * (DataManager.cpp:992-996), leaves d_varParts zeroed for numParticles rows. */
void DataManagerTransferLocalTree(void *moments, size_t sMoments,
                                  void *compactParts, size_t sCompactParts,
                                  void *varParts, size_t sVarParts,
                                  void **d_localMoments, void **d_compactParts,
                                  void **d_varParts, cudaStream_t stream,
                                  int numParticles, void *callback);

/* HostCUDA.h:108-112; caller DataManager.cpp:590. */
void DataManagerTransferRemoteChunk(void *moments, size_t sMoments,
                                    void *compactParts, size_t sCompactParts,
                                    void **d_remoteMoments, void **d_remoteParts,
                                    cudaStream_t stream, void *callback);

/* HostCUDA.h:114; caller DataManager.cpp:986. */
void TransferParticleVarsBack(VariablePartData *hostBuffer, size_t size,
                              void *d_varParts, cudaStream_t stream, void *cb);

/* HostCUDA.h:116-118; callers Compute.cpp:2133,2139,2151 (particle-cell). */
void TreePieceCellListDataTransferLocal(CudaRequest *data);
void TreePieceCellListDataTransferRemote(CudaRequest *data);
void TreePieceCellListDataTransferRemoteResume(CudaRequest *data);

/* HostCUDA.h:121-124; callers Compute.cpp:2228,2233,2241,2253 (particle-particle). */
void TreePiecePartListDataTransferLocal(CudaRequest *data);
void TreePiecePartListDataTransferLocalSmallPhase(CudaRequest *data,
                                                  CompactPartData *parts, int len);
void TreePiecePartListDataTransferRemote(CudaRequest *data);
void TreePiecePartListDataTransferRemoteResume(CudaRequest *data);

/* CudaFunctions.h:7-8.  Kept for ABI completeness; the library itself no
 * longer allocates per request (see DESIGN.md "device arena"). */
void TreePieceDataTransferBasic(CudaRequest *data, CudaDevPtr *ptr);
void TreePieceDataTransferBasicCleanup(CudaDevPtr *ptr);

/* EwaldCUDA.h:59-62; callers Ewald.cpp:403,527,539. */
void EwaldHostMemorySetup(EwaldData *h_idata, int size, int nEwhLoop, int largephase);
void EwaldHostMemoryFree(EwaldData *h_idata, int largephase);
void EwaldHost(CompactPartData *d_localParts, VariablePartData *d_localVars,
               EwaldData *h_idata, cudaStream_t stream, void *cb, int myIndex,
               int largephase);

extern "C" {
#endif /* __cplusplus */

/* ============================ PART 2: C ABI ============================== */

/* --- introspection ------------------------------------------------------- */
int cb200_abi_version(void);          /* bumps when a signature changes      */
int cb200_real_bytes(void);           /* sizeof(cudatype): 4, or 8 in FP64   */
const char *cb200_build_info(void);   /* arch, flags, kernel variants        */

/* completion tokens: standalone stand-in for Charm++ HAPI.  handler(cb) runs
 * on a CUDA host-callback thread once the stream reaches the signal point. */
typedef void (*cb200_callback_fn)(void *cb);
void cb200_set_callback_handler(cb200_callback_fn handler);

/* streams for FFI users that cannot create a cudaStream_t themselves */
void *cb200_stream_create(void);
void cb200_stream_destroy(void *stream);
void cb200_stream_synchronize(void *stream);
void cb200_device_synchronize(void);
void cb200_set_device(int ordinal);
void cb200_device_free(void *dptr);   /* what DataManager's cudaFree() does  */

/* --- the reference surface, one-to-one ------------------------------------ */
void cb200_allocatePinnedHostMemory(void **ptr, size_t size);
void cb200_freePinnedHostMemory(void *ptr);
void cb200_DataManagerTransferLocalTree(void *moments, size_t sMoments,
                                        void *compactParts, size_t sCompactParts,
                                        void *varParts, size_t sVarParts,
                                        void **d_localMoments, void **d_compactParts,
                                        void **d_varParts, void *stream,
                                        int numParticles, void *callback);
void cb200_DataManagerTransferRemoteChunk(void *moments, size_t sMoments,
                                          void *compactParts, size_t sCompactParts,
                                          void **d_remoteMoments, void **d_remoteParts,
                                          void *stream, void *callback);
void cb200_TransferParticleVarsBack(void *hostBuffer, size_t size, void *d_varParts,
                                    void *stream, void *cb);
void cb200_TreePieceCellListDataTransferLocal(CudaRequest *data);
void cb200_TreePieceCellListDataTransferRemote(CudaRequest *data);
void cb200_TreePieceCellListDataTransferRemoteResume(CudaRequest *data);
void cb200_TreePiecePartListDataTransferLocal(CudaRequest *data);
void cb200_TreePiecePartListDataTransferLocalSmallPhase(CudaRequest *data,
                                                        CompactPartData *parts, int len);
void cb200_TreePiecePartListDataTransferRemote(CudaRequest *data);
void cb200_TreePiecePartListDataTransferRemoteResume(CudaRequest *data);
void cb200_EwaldHostMemorySetup(EwaldData *h_idata, int size, int nEwhLoop, int largephase);
void cb200_EwaldHostMemoryFree(EwaldData *h_idata, int largephase);
void cb200_EwaldHost(void *d_localParts, void *d_localVars, EwaldData *h_idata,
                     void *stream, void *cb, int myIndex, int largephase);

/* --- device-resident variants (inputs already in HBM) ---------------------
 * Same kernels as the *ListDataTransfer* entry points, but the flattened
 * list / markers / starts / sizes are device pointers: no H2D copy is issued.
 * Used by the multi-GPU driver (lists built or staged on the device) and by
 * bench.py's kernel-only timing.  `moments`/`sources` choose the gather array
 * (local, remote chunk, or a missed-data buffer). */
void cb200_cell_list_device(void *d_parts, void *d_vars, void *d_moments,
                            const ILCell *d_list, const int *d_markers,
                            const int *d_starts, const int *d_sizes,
                            int numBuckets, cudatype fperiod, void *stream);
void cb200_part_list_device(void *d_parts, void *d_vars, void *d_sources,
                            const ILCell *d_list, const int *d_markers,
                            const int *d_starts, const int *d_sizes,
                            int numBuckets, cudatype fperiod, void *stream);
/* same, with the largest bucket size of the request stated by the caller (the
 * host entry points scan CudaRequest::bucketSizes for it); it selects how many
 * target particles a warp keeps in registers per pass.  0 = unknown. */
void cb200_cell_list_device_ex(void *d_parts, void *d_vars, void *d_moments,
                               const ILCell *d_list, const int *d_markers,
                               const int *d_starts, const int *d_sizes,
                               int numBuckets, cudatype fperiod, int maxBucketSize,
                               void *stream);
void cb200_part_list_device_ex(void *d_parts, void *d_vars, void *d_sources,
                               const ILCell *d_list, const int *d_markers,
                               const int *d_starts, const int *d_sizes,
                               int numBuckets, cudatype fperiod, int maxBucketSize,
                               void *stream);
void cb200_ewald_device(void *d_parts, void *d_vars, const int *d_markers,
                        int nActive, const EwaldReadOnlyData *h_ro,
                        const EwtData *h_ewt, void *stream);

/* device layout (DESIGN.md "data layout in HBM"): the d_localMoments /
 * d_localParts handles returned by the transfer functions address arrays of
 * packed rows, not the caller's AoS records.  Callers that fill device arrays
 * themselves (multi-GPU driver: records arrive by NCCL all-gather) convert
 * with these; d_raw holds CudaMultipoleMoments / CompactPartData records. */
size_t cb200_packed_moment_bytes(void);   /* bytes per cell row      */
size_t cb200_packed_particle_bytes(void); /* bytes per particle row  */
void cb200_pack_moments_device(const void *d_raw, void *d_packed, int n, void *stream);
void cb200_pack_particles_device(const void *d_raw, void *d_packed, int n, void *stream);
/* stream-ordered copy between any two of {device, pinned host} (cudaMemcpyDefault) */
void cb200_copy_device(void *dst, const void *src, size_t bytes, void *stream);
/* what ZeroVars does (HostCUDA.cu:2195-2205): clear n VariablePartData rows */
void cb200_zero_vars_device(void *d_vars, int n, void *stream);

/* --- timing taps (CUDA events recorded around every kernel we launch) ------ */
void cb200_timing_enable(int on);
void cb200_timing_reset(void);
/* out[0..2] = ms in p-c, p-p, Ewald kernels; out[3..5] = launches of each */
void cb200_timing_read(double out[6]);
long long cb200_kernel_launches(void);

/* --- new: tree-moment build on the device (oracle = moments.c) -- */
/* Leaves: particles added in index order about the running centre of mass,
 * radius = farthest particle.  Internal nodes: children combined bottom-up in
 * order 0,1, radius = farthest corner of the node's bounding box.  Topology
 * arrays are int32 device arrays, nodes in BFS order (child index > parent
 * index, levels contiguous), child -1 = absent.  geo* = the box a bucket got
 * from its parent's split (initial scale), box* = tight bounding boxes.
 * h_levelStart is a HOST array of numLevels+1 node offsets.  All math FP64;
 * d_moments_out receives CudaMultipoleMoments records (cudatype),
 * d_moments_f64_out (optional) the same 27 values in double. */
void cb200_build_moments(const double *d_pos_xyz, const double *d_mass,
                         const double *d_soft, int numParticles,
                         const int *d_child0, const int *d_child1,
                         const int *d_firstPart, const int *d_lastPart,
                         const double *d_geolo_xyz, const double *d_geohi_xyz,
                         const double *d_boxlo_xyz, const double *d_boxhi_xyz,
                         const int *h_levelStart, int numLevels, int numNodes,
                         void *d_moments_out, double *d_moments_f64_out,
                         void *stream);

/* --- new: tree topology built on the device --------------------- */
/* The single-TreePiece tree of csrc/treewalk.cpp (GenericTreeNode.h:473-598,
 * Compute.cpp:2476-2477, DataManager.cpp:797-828) from UNSORTED device arrays:
 * Morton keys in [rootlo, roothi), stable sort, binary splits on key bits, nodes
 * breadth first, buckets in particle order, tight boxes.  Every array equals the
 * host build bit for bit.  The result feeds cb200_build_moments and
 * cb200_walk_device directly; d_order[i] = caller index of sorted particle i.
 * Arrays live in the stream-ordered pool (cb200_tree_free).  error != 0: the node
 * capacity (1.5 n + 4096) was exceeded. */
typedef struct cb200_tree {
  double *d_pos, *d_mass, *d_soft; /* sorted particles */
  void *d_packedParts;             /* the same as packed particles (layout of the force kernels) */
  int *d_order;
  int *d_child0, *d_child1, *d_parent, *d_first, *d_last;
  double *d_geolo, *d_geohi, *d_boxlo, *d_boxhi;
  int *d_bucketNode, *d_bucketFirst, *d_bucketCount, *d_bucketStarts, *d_bucketSizes;
  int numParticles, numNodes, numBuckets, numLevels;
  int levelStart[66]; /* HOST: node offsets of the levels, numLevels + 1 entries */
  int error;
} cb200_tree;

void cb200_build_tree(const double *d_pos_xyz, const double *d_mass, const double *d_soft,
                      int numParticles, int maxBucket, const double *rootlo,
                      const double *roothi, cb200_tree *out, void *stream);
void cb200_tree_free(cb200_tree *tree, void *stream);

/* --- new: interaction lists built on the device ----------------- */
/* The double walk of TreeWalk.cpp:308-397 / Compute.cpp:690-884,1608-1863 on the
 * GPU: per bucket exactly the entries, order and offsetID bits of ChaNGa's host
 * walk, already in the layout the *_device_ex entry points consume, so the lists
 * never cross PCIe.  Tree arrays are DEVICE pointers, nodes in breadth-first
 * (nodeArrayIndex) order, child -1 = absent; h_levelStart is a HOST array of
 * numLevels+1 node offsets; d_moments_f64 = 27 doubles per node as written by
 * cb200_build_moments; boxes are the tight bounding boxes.  Only buckets
 * [bucketLo, bucketHi) get lists (a rank's SFC share; the whole tree is the source).
 * Every array of the result lives in the stream-ordered pool; release it with
 * cb200_lists_free.  error 1/2: a per-node or pool capacity was exceeded; error 3: a list of the
 * range has more than 2^31-1 entries (the markers are int, like the ABI's): walk a narrower
 * bucket range.  Nothing is usable when error != 0. */
typedef struct cb200_lists {
  ILCell *d_cell, *d_soft, *d_part; /* cells | softened cells (index = node) | particles, expanded */
  int *d_cellMarkers, *d_softMarkers, *d_partMarkers; /* numBuckets+1 each; empty lists outside the range */
  int *d_starts, *d_sizes;          /* first target particle / particle count of every bucket */
  void *d_nodeParticles;            /* every node as a packed source particle {cm, M | soft}: sources of d_soft (NULL when nSoft == 0) */
  long long nCell, nSoft, nPart;
  int numBuckets;
  int error;
} cb200_lists;
void cb200_walk_device(int numNodes, int numBuckets, int numLevels, const int *h_levelStart,
                       const int *d_child0, const int *d_child1, const int *d_parent,
                       const int *d_firstPart, const int *d_lastPart,
                       const int *d_bucketFirst, const int *d_bucketCount, const int *d_bucketNode,
                       const double *d_boxlo_xyz, const double *d_boxhi_xyz, const double *d_moments_f64,
                       double theta, int nReplicas, double period, int bucketLo, int bucketHi,
                       cb200_lists *out, void *stream);
/* Multistep (Compute.cpp:1278,1574: only buckets with rungs >= activeRung get lists):
 * the same walk restricted to the buckets whose byte in d_bucketActive (numBuckets bytes on the
 * device) is non-zero; NULL = every bucket = cb200_walk_device.  Entries, order and offsetID bits
 * equal the host walk's with the same mask (cb200h_walk). */
void cb200_walk_device_active(int numNodes, int numBuckets, int numLevels, const int *h_levelStart,
                              const int *d_child0, const int *d_child1, const int *d_parent,
                              const int *d_firstPart, const int *d_lastPart,
                              const int *d_bucketFirst, const int *d_bucketCount, const int *d_bucketNode,
                              const double *d_boxlo_xyz, const double *d_boxhi_xyz, const double *d_moments_f64,
                              double theta, int nReplicas, double period, int bucketLo, int bucketHi,
                              const unsigned char *d_bucketActive, cb200_lists *out, void *stream);
/* The active sets of a multistep force step from per-particle rungs (one byte per particle, indexed
 * through d_order when it is not NULL, i.e. in the caller's order): d_bucketActive[b] = some particle
 * of bucket b has rung >= activeRung (GenericTreeNode::rungs, Compute.cpp:1278); d_ewaldMarkers =
 * ascending tree-order indices of the particles with rung >= activeRung (the marker array of
 * EwaldHost's large phase, Ewald.cpp:416-437; capacity numParticles ints).  h_counts[0] = active
 * buckets, h_counts[1] = active particles; synchronises the stream. */
void cb200_active_sets_device(const unsigned char *d_rung, const int *d_order, int numParticles,
                              const int *d_bucketStarts, const int *d_bucketSizes, int numBuckets, int activeRung,
                              unsigned char *d_bucketActive, int *d_ewaldMarkers, int *h_counts, void *stream);
void cb200_lists_free(cb200_lists *lists, void *stream);

/* --- new: per-GPU bucket partitioner --------------------------- */
/* Cuts numBuckets SFC-ordered buckets into nRanks contiguous ranges of equal
 * summed cost; cuts[0]=0 .. cuts[nRanks]=numBuckets.  Host-side, O(n). */
void cb200_partition_buckets(const double *cost, int numBuckets, int nRanks,
                             int *cuts);

/* --- new: one communicator per process, the whole force step in the library -------------- */
/* One process per GPU (one logical node per device, DataManager.h:329-337).  The host program makes a
 * 128-byte id on one rank (cb200_comm_unique_id), hands it to the others by its own means (a Charm++
 * broadcast, MPI, a file) and every rank calls cb200_comm_init after cb200_set_device.  NCCL is loaded
 * with dlopen("libnccl.so.2") on first use: single-GPU hosts do not need it. */
typedef struct cb200_comm cb200_comm;
size_t cb200_comm_id_bytes(void);                 /* sizeof(ncclUniqueId) = 128 */
void cb200_comm_unique_id(void *id);              /* ncclGetUniqueId */
cb200_comm *cb200_comm_init(int rank, int world, const void *id);
void cb200_comm_destroy(cb200_comm *comm);
int cb200_comm_rank(const cb200_comm *comm);      /* 0 for NULL */
int cb200_comm_world(const cb200_comm *comm);     /* 1 for NULL */
int cb200_comm_nccl_version(void);
/* element-wise reduction of n HOST doubles over the ranks (op 0 sum, 1 max, 2 min); blocking */
void cb200_comm_allreduce_f64(cb200_comm *comm, double *h_values, int n, int op, void *stream);
void cb200_comm_barrier(cb200_comm *comm, void *stream);
/* ncclAllGather of equal byte slices: d_recv gets world x sendBytes in rank order */
void cb200_comm_allgather(cb200_comm *comm, const void *d_send, void *d_recv, size_t sendBytes, void *stream);

/* The force step from UNSORTED particles (SURVEY f1 + f2 + 8e).  Every rank holds rows
 * [rank * chunk, (rank + 1) * chunk) of the box's {x, y, z, mass, soft} records (doubles, caller order;
 * chunk = ceil(numParticles / world), the tail of the last rank is padding).  Per step: upload of the
 * slice, ONE ncclAllGather of the records, then on every rank keys, sort, tree, boxes, moments (all
 * replicated), and for the rank's own contiguous SFC range of buckets the interaction lists, p-c, p-p,
 * softened cells and the Ewald sum; the rank's rows come back with the caller indices they belong to.
 * Bucket ranges: equal particle counts, or (costCuts) equal cost as measured by the previous step
 * (198 x p-c pairs + 30 x p-p pairs; never splits a bucket).  Multistep: see
 * cb200_active_sets_device; ranks are then cut by active-particle count. */
typedef struct cb200_step_config {
  long long numParticles;   /* the whole box, all ranks */
  int maxBucket;            /* 12 */
  int nReplicas;            /* periodic replicas of the walk (0: isolated) */
  int ewald;                /* 1: add the Ewald correction (periodic boxes) */
  int activeRung;           /* > 0: multistep step, rung bytes are uploaded with the records */
  int overlapEwald;         /* 1: the Ewald kernel runs on a second stream under the tree walk */
  int costCuts;             /* 1: cut ranks by last step's measured cost */
  double theta, period, dEwCut, dEwhCut;
  double rootlo[3], roothi[3];
} cb200_step_config;

enum { CB200_PH_H2D = 0, CB200_PH_GATHER, CB200_PH_TREE, CB200_PH_MOMENTS, CB200_PH_EWALD, CB200_PH_WALK,
       CB200_PH_PC, CB200_PH_PP, CB200_PH_FINISH, CB200_PH_TOTAL, CB200_PH_COUNT };

typedef struct cb200_step_result {
  int error;                /* 0 ok; 11: node capacity; 21-23: walk (see cb200_lists); 30: result buffer too small */
  int numNodes, numBuckets, numLevels;
  int bucketLo, bucketHi, partLo, partHi; /* this rank's share (tree order) */
  int activeBuckets, activeParticles;     /* multistep */
  int rows;                               /* result rows of this rank */
  long long nCell, nSoft, nPart;          /* list entries of this rank */
  long long pcPairs, ppPairs;             /* pair interactions of this rank (Compute.cpp:1643-1651) */
  long long h2dBytes, d2hBytes;
  double cost;                            /* 198 pcPairs + 30 ppPairs */
  float ms[CB200_PH_COUNT];               /* CUDA events on the step's stream; with overlapEwald the Ewald
                                             kernel's time lies inside the walk phase */
  int letBlockLevel;                      /* multi-GPU: tree level below which this rank built only the subtrees near its
                                             own buckets (locally essential moment build); -1: everything built */
  int letFallback;                        /* 1: this step's walk left that part and was repeated on the full build */
  int treeRebuilt, walkRepeated;          /* 1: the node arrays / the walk's pools, sized from the last step, were too small
                                             and the phase was repeated with worst-case sizes */
} cb200_step_result;

/* the costCuts rule on the host: last step's boundaries prevCut[0..world] (particle indices) and the cost
 * each rank measured between them -> targets[0..world] for the next step (cost taken as uniform inside
 * each old range; the device snaps every target to the next bucket start) */
void cb200_cost_targets(const long long *prevCut, const double *prevCost, int world, long long numParticles, int *targets);

typedef struct cb200_step cb200_step;
cb200_step *cb200_step_create(cb200_comm *comm /* NULL: one GPU */, const cb200_step_config *cfg);
void cb200_step_destroy(cb200_step *step);
int cb200_step_chunk_rows(const cb200_step *step);
int cb200_step_out_capacity(const cb200_step *step);
void *cb200_step_stream(const cb200_step *step);
double *cb200_step_device_records(const cb200_step *step);
unsigned char *cb200_step_device_rungs(const cb200_step *step);
/* h_records NULL: the records are already in cb200_step_device_records().  h_out NULL: results stay on the
 * device.  world > 1, or world == 1 with h_index != NULL: h_out gets result->rows VariablePartData rows of this rank's
 * SFC range (tree order, as TransferParticleVarsBack returns a TreePiece's rows) and h_index their caller indices;
 * the rows are copied back in up to eight slabs (about a million rows each), each as soon as its list kernels are done, under the next slab's
 * kernels.  world == 1 with h_index == NULL: numParticles rows in the caller's order (one scatter on the device and
 * one copy at the end).  Blocking. */
void cb200_step_run(cb200_step *step, const double *h_records, const unsigned char *h_rungs, void *h_out, int *h_index,
                    int outCapacityRows, int keepLists, cb200_step_result *result);
/* products of the last run, valid until the next run / destroy (tests, parity sampling) */
const cb200_tree *cb200_step_tree(const cb200_step *step);
const cb200_lists *cb200_step_lists(const cb200_step *step);     /* only after keepLists = 1 */
const double *cb200_step_moments_f64(const cb200_step *step);    /* 27 doubles per node */
const void *cb200_step_packed_moments(const cb200_step *step);
const void *cb200_step_vars(const cb200_step *step);             /* VariablePartData, tree order */
const int *cb200_step_markers(const cb200_step *step);           /* multistep: active particles, tree order */

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* CHANGA_B200_API_H */
