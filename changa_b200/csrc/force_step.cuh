/* force_step.cuh -- the whole force step inside the library, one process per GPU (included by hostcuda.cu).
 *
 *   cb200_comm_*   one NCCL communicator per process (ncclCommInitRank from a 128-byte id the host
 *                  program distributes however it likes: Charm++ broadcast, MPI, a file, a TCP store).
 *                  NCCL is loaded with dlopen("libnccl.so.2") the first time a communicator is made, so
 *                  a single-GPU host needs no NCCL at all.
 *   cb200_step_*   the step: every rank holds 1/N of the {x, y, z, mass, soft} records of the box
 *                  (rows [rank*chunk, (rank+1)*chunk) in the caller's order); ONE ncclAllGather over
 *                  NVLink replicates them; every rank then builds the same tree and the same
 *                  moments, walks, evaluates and returns only its own contiguous SFC range of buckets
 *                  (owner computes: accelerations never leave the GPU that made them).  Keys, sort,
 *                  topology, boxes, moments, interaction lists, p-c, p-p, softened cells and the Ewald
 *                  sum all run on the device; the Ewald h-table is built on the device from the root
 *                  moments (no root-moment round trip to the host).
 *
 * Reference analogues: DataManager::serializeLocalTree / transferLocalTree (DataManager.cpp:797-903),
 * TreePiece::startGravity -> ListCompute::stateReady -> sendNodeInteractionsToGpu (Compute.cpp:1608-2253),
 * TreePiece::EwaldGPU (Ewald.cpp:387-517), DataManager::transferParticleVarsBack (DataManager.cpp:962-996);
 * one logical node per device as in DataManager.h:329-337.  The reference has no counterpart of the
 * all-gather (it ships needed remote data through CkCache).
 *
 * Host synchronisations per step: one in the tree build (node / bucket / level counts), one in the walk
 * (list lengths size the lists), one at the end; a multistep step adds one (active counts). */
#ifndef CB200_FORCE_STEP_CUH
#define CB200_FORCE_STEP_CUH

#include <dlfcn.h>
#include <nccl.h>

/* ------------------------------------------------------------------ NCCL */
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};
static NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    /* an NCCL already in the process (a host that links one, torch's bundled copy) is found by its soname */
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) {
      fprintf(stderr, "changa_b200: multi-GPU requested but libnccl.so.2 cannot be loaded (%s)\n", dlerror());
      abort();
    }
    auto sym = [&](const char *name) {
      void *p = dlsym(api.handle, name);
      if (!p) { fprintf(stderr, "changa_b200: %s missing from libnccl\n", name); abort(); }
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  });
  return &api;
}
#define ncclChk(call)                                                                               \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess) {                                                                        \
      fprintf(stderr, "Fatal NCCL Error %s at %s:%d\n", nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
      abort();                                                                                      \
    }                                                                                               \
  } while (0)

struct cb200_comm {
  ncclComm_t comm;
  int rank, world;
};

/* ----------------------------------------------------------------- kernels */
/* EwaldInit on the device (Ewald.cpp:285-375; ewald_setup.cuh is shared with the host build):
 * thread 0 completes the root moments, then one thread per h-vector fills its table row. */
__global__ void ewald_setup_kernel(const double *__restrict__ mom64, double L, double fEwCut, int nReps, int nEwh,
                                   const int *__restrict__ hxyz, int first, int last, EwaldParams *__restrict__ out) {
  __shared__ double comp[125];
  __shared__ double momc[32];
  if (threadIdx.x == 0) {
    ewald_complete_moments(mom64, comp, momc);
    EwaldReadOnlyData &ro = out->ro;
    ro.mm.totalMass = (real)mom64[2];
    ro.mm.cmx = (real)mom64[3]; ro.mm.cmy = (real)mom64[4]; ro.mm.cmz = (real)mom64[5];
    real *q = reinterpret_cast<real *>(&ro.momcRoot);
    for (int i = 0; i < 32; ++i) q[i] = (real)momc[i];
    const double alpha = 2.0 / L;
    ro.n = last - first + 1; ro.nReps = nReps; ro.nEwReps = (int)ceil(fEwCut); ro.nEwhLoop = nEwh;
    ro.L = (real)L; ro.fEwCut = (real)fEwCut; ro.alpha = (real)alpha; ro.alpha2 = (real)(alpha * alpha);
    ro.k1 = (real)(3.14159265358979323846 / (alpha * alpha * L * L * L));
    ro.ka = (real)(2.0 * alpha / sqrt(3.14159265358979323846));
    ro.fEwCut2 = (real)(fEwCut * fEwCut * L * L);
    ro.fInner2 = (real)(1.1e-2 * L * L); /* Ewald.cpp:516 (not used by ewald_kernel) */
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nEwh; i += blockDim.x) {
    double row[5];
    ewald_h_row(comp, mom64[2], L, hxyz[3 * i], hxyz[3 * i + 1], hxyz[3 * i + 2], row);
    EwtData &e = out->ewt[i];
    e.hx = (real)row[0]; e.hy = (real)row[1]; e.hz = (real)row[2]; e.hCfac = (real)row[3]; e.hSfac = (real)row[4];
  }
}

/* pair interactions of buckets [b0, b1), counted like Compute.cpp:1643-1651 (list length x bucket
 * size): out[0] = p-c, out[1] = p-p (+ softened cells) */
__global__ void step_pairs_kernel(const int *__restrict__ cellMarkers, const int *__restrict__ partMarkers,
                                  const int *__restrict__ softMarkers, const int *__restrict__ sizes, int b0, int b1,
                                  unsigned long long *__restrict__ out) {
  unsigned long long pc = 0, pp = 0;
  for (int b = b0 + blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += gridDim.x * blockDim.x) {
    const unsigned long long z = (unsigned long long)sizes[b];
    pc += z * (unsigned long long)(cellMarkers[b + 1] - cellMarkers[b]);
    pp += z * (unsigned long long)((partMarkers[b + 1] - partMarkers[b]) + (softMarkers[b + 1] - softMarkers[b]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    pc += __shfl_xor_sync(0xffffffffu, pc, o);
    pp += __shfl_xor_sync(0xffffffffu, pp, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out, pc); atomicAdd(out + 1, pp); }
}

/* results back to the caller's particle order: out[order[i]] = vars[i] (single GPU), or the rank's own
 * rows [p0, p1) with the caller index each belongs to */
__global__ void step_scatter_kernel(const VariablePartData *__restrict__ vars, const int *__restrict__ order, int n,
                                    VariablePartData *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[order[i]] = vars[i];
}

/* multistep rank boundaries: boundary r = first bucket starting at or after the particle that holds
 * active marker at[r] (changa_b200.multigpu.bucket_range_by_active) */
__global__ void step_active_cuts_kernel(const int *__restrict__ markers, const int *__restrict__ at, int nCuts,
                                        const int *__restrict__ bucketStarts, int nb, int n, int *__restrict__ cuts) {
  const int r = threadIdx.x;
  if (r >= nCuts) return;
  const int want = markers[at[r]];
  int a = 0, b = nb;
  while (a < b) {
    const int mid = a + ((b - a) >> 1);
    if (bucketStarts[mid] < want) a = mid + 1; else b = mid;
  }
  cuts[2 * r] = a;
  cuts[2 * r + 1] = a < nb ? bucketStarts[a] : n;
}
/* out[k] = number of markers below p0 (k = 0) / p1 (k = 1): the slice of the ascending marker array in [p0, p1) */
__global__ void step_marker_range_kernel(const int *__restrict__ markers, int nAct, int p0, int p1, int *__restrict__ out) {
  const int want = threadIdx.x == 0 ? p0 : p1;
  int a = 0, b = nAct;
  while (a < b) {
    const int mid = a + ((b - a) >> 1);
    if (markers[mid] < want) a = mid + 1; else b = mid;
  }
  out[threadIdx.x] = a;
}

/* ------------------------------------------------------------------- step */
constexpr int kOutSlabsMax = 8;
enum { PH_H2D = 0, PH_GATHER, PH_TREE, PH_MOMENTS, PH_EWALD, PH_WALK, PH_PC, PH_PP, PH_FINISH, PH_TOTAL, PH_N };

struct cb200_step {
  cb200_comm *comm = nullptr;
  cb200_step_config cfg;
  int rank = 0, world = 1, device = 0;
  long long n = 0;
  int chunk = 0;
  cudaStream_t stream = nullptr, aux = nullptr, outStream = nullptr;
  cudaEvent_t ev[PH_N + 2], evFork = nullptr, evJoin = nullptr, evTree = nullptr, evOut = nullptr;
  cudaEvent_t evSlab[kOutSlabsMax][3]; /* start / after p-c / after p-p of every output slab */
  int ewaldSlot = 0;
  int nEwh = 0;
  int *d_hxyz = nullptr;
  EwaldParams *d_ewald = nullptr;
  double *d_rec = nullptr, *d_all = nullptr;       /* my slice / the gathered box */
  unsigned char *d_rung = nullptr, *d_rungAll = nullptr;
  VariablePartData *d_vars = nullptr, *d_out = nullptr;
  unsigned long long *d_counts = nullptr;
  /* node-sized arrays follow the tree */
  int nodeCap = 0;
  double *d_mom64 = nullptr;
  PackedCell *d_pkMom = nullptr;
  real *d_mom32 = nullptr;
  int bucketCap = 0;
  unsigned char *d_bucketActive = nullptr;
  int *d_markers = nullptr;
  /* what the last run left for inspection (tests; cb200_step_tree / _lists) */
  cb200_tree tree;
  cb200_lists lists;
  bool haveTree = false, haveLists = false;
  /* sizes learned from the last step: node arrays (tree build) and the walk's pools, instead of their worst cases */
  double treeCapFactor = 0.0;              /* 0: not known yet */
  unsigned long long poolHint[3] = {0, 0, 0}; /* capacities of the walk's pools for the next step (0: from the tree) */
  long long stepsRun = 0;
  /* locally essential moment build (let_kernels.cuh) */
  bool letOff = false;
  unsigned char *d_letFlag = nullptr;
  LetBox *d_letCover = nullptr;
  double *d_letScalars = nullptr; /* [0] largest particle softening, [1] (int) cover count */
  int letLevel = -1;
  /* cost feedback: last step's particle cuts and the cost every rank measured between them */
  std::vector<double> prevCost;
  std::vector<long long> prevCut;
};

static std::atomic<int> g_nextEwaldSlot{0};

static void step_release_products(cb200_step *st) {
  cudaStream_t s = st->stream;
  if (st->haveLists) { cb200_lists_free(&st->lists, s); st->haveLists = false; }
  if (st->haveTree) { cb200_tree_free(&st->tree, s); st->haveTree = false; }
}

/* New particle targets of the rank boundaries from last step's measured costs: the cost is taken as
 * uniform inside each old range, the cumulative cost is inverted at k/N of its total (never splits
 * a bucket: the device snaps each target to the next bucket start).  Pure host arithmetic on values
 * every rank holds identically, so all ranks cut alike. */
static void cost_targets(const std::vector<long long> &cut, const std::vector<double> &cost, int world, long long n,
                         int *targets) {
  double total = 0.0;
  for (double c : cost) total += c;
  int seg = 0;
  double before = 0.0;
  for (int r = 1; r < world; ++r) {
    const double want = total * r / world;
    while (seg < world - 1 && before + cost[seg] < want) { before += cost[seg]; ++seg; }
    const double span = (double)(cut[seg + 1] - cut[seg]);
    const double frac = cost[seg] > 0.0 ? (want - before) / cost[seg] : 0.0;
    long long p = cut[seg] + (long long)(frac * span + 0.5);
    if (p < 0) p = 0;
    if (p > n) p = n;
    targets[r] = (int)p;
  }
  targets[0] = 0;
  targets[world] = (int)n;
}

extern "C" {

/* ---- communicator ---- */
size_t cb200_comm_id_bytes(void) { return sizeof(ncclUniqueId); }
void cb200_comm_unique_id(void *id) {
  ncclUniqueId u;
  ncclChk(nccl_api()->GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
}
cb200_comm *cb200_comm_init(int rank, int world, const void *id) {
  device_info();
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  cb200_comm *c = new cb200_comm;
  c->rank = rank; c->world = world;
  ncclChk(nccl_api()->CommInitRank(&c->comm, world, u, rank));
  return c;
}
void cb200_comm_destroy(cb200_comm *c) {
  if (!c) return;
  ncclChk(nccl_api()->CommDestroy(c->comm));
  delete c;
}
int cb200_comm_rank(const cb200_comm *c) { return c ? c->rank : 0; }
int cb200_comm_world(const cb200_comm *c) { return c ? c->world : 1; }
int cb200_comm_nccl_version(void) {
  int v = 0;
  ncclChk(nccl_api()->GetVersion(&v));
  return v;
}
/* element-wise reduction of n host doubles over the ranks (op: 0 sum, 1 max, 2 min); blocking.
 * With one rank (comm NULL) the values stay as they are. */
void cb200_comm_allreduce_f64(cb200_comm *c, double *h_values, int n, int op, void *stream) {
  if (!c || c->world == 1 || n <= 0) return;
  cudaStream_t s = (cudaStream_t)stream;
  double *d = (double *)pool_alloc((size_t)n * sizeof(double), s);
  cudaChk(cudaMemcpyAsync(d, h_values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  ncclChk(nccl_api()->AllReduce(d, d, (size_t)n, ncclDouble, op == 1 ? ncclMax : (op == 2 ? ncclMin : ncclSum), c->comm, s));
  cudaChk(cudaMemcpyAsync(h_values, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  cudaChk(cudaStreamSynchronize(s));
  pool_free(d, s);
}
void cb200_comm_barrier(cb200_comm *c, void *stream) {
  double one = 1.0;
  cb200_comm_allreduce_f64(c, &one, 1, 0, stream);
  cudaChk(cudaStreamSynchronize((cudaStream_t)stream));
}
/* replicate equal slices: every rank contributes sendBytes at d_send, d_recv gets world x sendBytes in rank order */
void cb200_comm_allgather(cb200_comm *c, const void *d_send, void *d_recv, size_t sendBytes, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!c || c->world == 1) {
    if (d_send != d_recv && sendBytes) cudaChk(cudaMemcpyAsync(d_recv, d_send, sendBytes, cudaMemcpyDeviceToDevice, s));
    return;
  }
  ncclChk(nccl_api()->AllGather(d_send, d_recv, sendBytes, ncclChar, c->comm, s));
}

/* the cut rule of costCuts as a plain host function (prevCut: world + 1 particle indices, prevCost: world
 * costs; targets: world + 1 particle indices out) */
void cb200_cost_targets(const long long *prevCut, const double *prevCost, int world, long long n, int *targets) {
  cost_targets(std::vector<long long>(prevCut, prevCut + world + 1), std::vector<double>(prevCost, prevCost + world), world, n,
               targets);
}

/* ---- step ---- */
cb200_step *cb200_step_create(cb200_comm *comm, const cb200_step_config *cfg) {
  device_info();
  cb200_step *st = new cb200_step;
  st->comm = comm;
  st->cfg = *cfg;
  st->rank = comm ? comm->rank : 0;
  st->world = comm ? comm->world : 1;
  cudaChk(cudaGetDevice(&st->device));
  st->n = cfg->numParticles;
  if (st->n <= 0 || st->n > 0x7fffffffLL || st->world > 17) {
    fprintf(stderr, "cb200_step_create: %lld particles on %d ranks is outside what one step handles (int32 particle "
                    "indices as in the reference's ABI; at most 17 ranks)\n", st->n, st->world);
    abort();
  }
  st->chunk = (int)((st->n + st->world - 1) / st->world);
  cudaChk(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
  cudaChk(cudaStreamCreateWithFlags(&st->aux, cudaStreamNonBlocking));
  cudaChk(cudaStreamCreateWithFlags(&st->outStream, cudaStreamNonBlocking));
  cudaChk(cudaEventCreateWithFlags(&st->evTree, cudaEventDisableTiming));
  cudaChk(cudaEventCreateWithFlags(&st->evOut, cudaEventDisableTiming));
  for (auto &e3 : st->evSlab) for (cudaEvent_t &e : e3) cudaChk(cudaEventCreate(&e));
  for (cudaEvent_t &e : st->ev) cudaChk(cudaEventCreate(&e));
  cudaChk(cudaEventCreateWithFlags(&st->evFork, cudaEventDisableTiming));
  cudaChk(cudaEventCreateWithFlags(&st->evJoin, cudaEventDisableTiming));
  const size_t n = (size_t)st->n, chunk = (size_t)st->chunk;
  cudaChk(cudaMalloc((void **)&st->d_rec, chunk * 5 * sizeof(double)));
  if (st->world > 1) cudaChk(cudaMalloc((void **)&st->d_all, chunk * st->world * 5 * sizeof(double)));
  cudaChk(cudaMalloc((void **)&st->d_vars, n * sizeof(VariablePartData)));
  if (st->world == 1) cudaChk(cudaMalloc((void **)&st->d_out, n * sizeof(VariablePartData)));
  cudaChk(cudaMalloc((void **)&st->d_counts, 64));
  if (cfg->activeRung > 0) {
    cudaChk(cudaMalloc((void **)&st->d_rung, chunk));
    if (st->world > 1) cudaChk(cudaMalloc((void **)&st->d_rungAll, chunk * st->world));
    cudaChk(cudaMalloc((void **)&st->d_markers, n * sizeof(int)));
  }
  if (cfg->ewald) {
    std::vector<int> h(3 * 4096);
    st->nEwh = ewald_h_vectors(cfg->dEwhCut, h.data(), 4096);
    if (st->nEwh > NEWH) { /* HostCUDA.cu:1923 */
      fprintf(stderr, "cb200_step_create: dEwhCut = %g gives %d h-vectors, the table holds %d\n", cfg->dEwhCut, st->nEwh, NEWH);
      abort();
    }
    cudaChk(cudaMalloc((void **)&st->d_hxyz, (size_t)(st->nEwh > 0 ? st->nEwh : 1) * 3 * sizeof(int)));
    cudaChk(cudaMemcpy(st->d_hxyz, h.data(), (size_t)st->nEwh * 3 * sizeof(int), cudaMemcpyHostToDevice));
    cudaChk(cudaMalloc((void **)&st->d_ewald, sizeof(EwaldParams)));
    cudaChk(cudaMemset(st->d_ewald, 0, sizeof(EwaldParams)));
    st->ewaldSlot = g_nextEwaldSlot.fetch_add(1) % kEwaldSlots;
  }
  memset(&st->tree, 0, sizeof st->tree);
  memset(&st->lists, 0, sizeof st->lists);
  stream_counter(st->stream); /* made here, not inside a launch */
  return st;
}

void cb200_step_destroy(cb200_step *st) {
  if (!st) return;
  cudaChk(cudaStreamSynchronize(st->stream));
  step_release_products(st);
  cudaChk(cudaStreamSynchronize(st->stream));
  for (void *p : {(void *)st->d_rec, (void *)st->d_all, (void *)st->d_vars, (void *)st->d_out, (void *)st->d_counts,
                  (void *)st->d_rung, (void *)st->d_rungAll, (void *)st->d_markers, (void *)st->d_hxyz, (void *)st->d_ewald,
                  (void *)st->d_mom64, (void *)st->d_pkMom, (void *)st->d_mom32, (void *)st->d_bucketActive, (void *)st->d_letFlag,
                  (void *)st->d_letCover, (void *)st->d_letScalars})
    if (p) cudaChk(cudaFree(p));
  for (cudaEvent_t &e : st->ev) cudaEventDestroy(e);
  cudaEventDestroy(st->evFork); cudaEventDestroy(st->evJoin); cudaEventDestroy(st->evTree); cudaEventDestroy(st->evOut);
  for (auto &e3 : st->evSlab) for (cudaEvent_t &e : e3) cudaEventDestroy(e);
  cudaStreamDestroy(st->outStream);
  pool_forget_stream(st->stream); pool_forget_stream(st->aux);
  drop_companion(st->stream); drop_stream_counter(st->stream);
  drop_companion(st->aux); drop_stream_counter(st->aux);
  cudaStreamDestroy(st->stream); cudaStreamDestroy(st->aux);
  delete st;
}

int cb200_step_chunk_rows(const cb200_step *st) { return st->chunk; }
void *cb200_step_stream(const cb200_step *st) { return (void *)st->stream; }
/* device buffer of this rank's chunk_rows x {x, y, z, mass, soft} doubles (and rung bytes): a caller that
 * keeps its particles on the device writes them here and calls cb200_step_run with h_records = NULL */
double *cb200_step_device_records(const cb200_step *st) { return st->d_rec; }
unsigned char *cb200_step_device_rungs(const cb200_step *st) { return st->d_rung; }
/* the tree / lists / moments / accumulators of the last run (valid until the next run or destroy) */
const cb200_tree *cb200_step_tree(const cb200_step *st) { return st->haveTree ? &st->tree : nullptr; }
const cb200_lists *cb200_step_lists(const cb200_step *st) { return st->haveLists ? &st->lists : nullptr; }
const double *cb200_step_moments_f64(const cb200_step *st) { return st->d_mom64; }
const void *cb200_step_packed_moments(const cb200_step *st) { return st->d_pkMom; }
const void *cb200_step_vars(const cb200_step *st) { return st->d_vars; }
const int *cb200_step_markers(const cb200_step *st) { return st->d_markers; }

/* One force step.  h_records: this rank's chunk_rows x 5 doubles in (pinned) host memory, or NULL when the
 * caller already wrote cb200_step_device_records().  h_rungs: chunk_rows bytes (multistep steps only).
 * h_out: accelerations, potential and dtGrav as VariablePartData rows -- world == 1: numParticles rows in the
 * CALLER'S particle order; world > 1: this rank's result->rows rows (its SFC range) with the caller index of
 * each row in h_index; outCapacityRows = rows h_out / h_index can hold (cb200_step_out_capacity()).  h_out NULL: results stay on the device
 * (cb200_step_vars, tree order).  Blocking: returns when the results are in h_out.  keepLists: leave the
 * interaction lists on the device for inspection (cb200_step_lists). */
void cb200_step_run(cb200_step *st, const double *h_records, const unsigned char *h_rungs, void *h_out, int *h_index,
                    int outCapacityRows, int keepLists, cb200_step_result *res) {
  cudaChk(cudaSetDevice(st->device));
  cudaStream_t s = st->stream;
  const cb200_step_config &cfg = st->cfg;
  const int n = (int)st->n, world = st->world, rank = st->rank, chunk = st->chunk;
  const bool multistep = cfg.activeRung > 0;
  memset(res, 0, sizeof *res);
  res->letBlockLevel = -1;
  step_release_products(st);
  nvtx_push("cb200_step_run");

  cudaChk(cudaEventRecord(st->ev[0], s));
  nvtx_push("CUDA_XFER_LOCAL");
  if (h_records) {
    cudaChk(cudaMemcpyAsync(st->d_rec, h_records, (size_t)chunk * 5 * sizeof(double), cudaMemcpyHostToDevice, s));
    res->h2dBytes += (long long)chunk * 5 * (long long)sizeof(double);
    if (multistep && h_rungs) {
      cudaChk(cudaMemcpyAsync(st->d_rung, h_rungs, (size_t)chunk, cudaMemcpyHostToDevice, s));
      res->h2dBytes += chunk;
    }
  }
  nvtx_pop();
  cudaChk(cudaEventRecord(st->ev[PH_H2D + 1], s));

  /* the one exchange of the step: every rank's slice of the 40-byte records */
  nvtx_push("all-gather");
  const double *box = st->d_rec;
  const unsigned char *rungs = st->d_rung;
  /* Distributed sort: every rank keys and sorts its OWN slice on the second stream while the records travel,
   * the sorted runs (key + caller index, 12 bytes per particle) are all-gathered and merged pairwise
   * (cub::DeviceMerge, stable: runs are in rank = caller order) -- log2(N) passes over the box instead of the
   * eight passes of a radix sort of all of it on every rank. */
  static const int dsortEnv = getenv("CB200_DSORT") ? atoi(getenv("CB200_DSORT")) : 1;
  const bool dsort = world > 1 && dsortEnv != 0;
  unsigned long long *runKeys[2] = {nullptr, nullptr};
  int *runIdx[2] = {nullptr, nullptr};
  unsigned long long *myKeys = nullptr;
  int *myIdx = nullptr;
  if (dsort) {
    const size_t all = (size_t)chunk * world;
    for (int k = 0; k < 2; ++k) {
      runKeys[k] = (unsigned long long *)pool_alloc(all * 8, s);
      runIdx[k] = (int *)pool_alloc(all * 4, s);
    }
    myKeys = (unsigned long long *)pool_alloc((size_t)chunk * 8, s);
    myIdx = (int *)pool_alloc((size_t)chunk * 4, s);
    cudaChk(cudaEventRecord(st->evFork, s));
    cudaChk(cudaStreamWaitEvent(st->aux, st->evFork, 0));
    const long long base = (long long)rank * chunk;
    const int myRows = (int)(base >= n ? 0 : (n - base < chunk ? n - base : chunk));
    TreeInput mine;
    mine.pos = st->d_rec; mine.mass = st->d_rec + 3; mine.soft = st->d_rec + 4; mine.posStride = 5; mine.attrStride = 5;
    tree_sorted_keys(mine, myRows, cfg.rootlo, cfg.roothi, (int)base, myKeys, myIdx, st->aux);
    cudaChk(cudaEventRecord(st->evJoin, st->aux));
  }
  if (world > 1) {
    ncclChk(nccl_api()->AllGather(st->d_rec, st->d_all, (size_t)chunk * 5, ncclDouble, st->comm->comm, s));
    box = st->d_all;
    if (multistep) {
      ncclChk(nccl_api()->AllGather(st->d_rung, st->d_rungAll, (size_t)chunk, ncclChar, st->comm->comm, s));
      rungs = st->d_rungAll;
    }
  }
  if (dsort) {
    cudaChk(cudaStreamWaitEvent(s, st->evJoin, 0));
    ncclChk(nccl_api()->AllGather(myKeys, runKeys[0], (size_t)chunk, ncclUint64, st->comm->comm, s));
    ncclChk(nccl_api()->AllGather(myIdx, runIdx[0], (size_t)chunk, ncclInt32, st->comm->comm, s));
  }
  nvtx_pop();
  cudaChk(cudaEventRecord(st->ev[PH_GATHER + 1], s));

  /* tree; the rank boundaries ride back with its one synchronisation */
  nvtx_push("CUDA_SER_TREE");
  int targets[18], cuts[2 * kTreeMaxCuts], fine[kTreeMaxCuts];
  const bool byCost = cfg.costCuts && !multistep && (int)st->prevCost.size() == world && world > 1;
  /* profiling aid: CB200_EMULATE_RANK="r/N" makes a single-GPU step do the work of rank r of N (its bucket
   * range only; tree and moments whole, as on every rank), so that ncu can look at one rank's share */
  int cutWorld = world, cutRank = rank;
  if (world == 1 && !multistep)
    if (const char *em = getenv("CB200_EMULATE_RANK")) {
      int r = 0, w = 1;
      if (sscanf(em, "%d/%d", &r, &w) == 2 && w >= 1 && w <= 17 && r >= 0 && r < w) { cutWorld = w; cutRank = r; }
    }
  if (byCost) cost_targets(st->prevCut, st->prevCost, world, n, targets);
  else for (int r = 0; r <= cutWorld; ++r) targets[r] = (int)((long long)r * n / cutWorld);
  TreeInput in;
  in.pos = box; in.mass = box + 3; in.soft = box + 4; in.posStride = 5; in.attrStride = 5;
  unsigned long long *preKeys = nullptr;
  int *preOrder = nullptr;
  if (dsort) {
    std::vector<long long> off;
    std::vector<int> len;
    for (int r = 0; r < world; ++r) {
      const long long base = (long long)r * chunk;
      off.push_back(base);
      len.push_back((int)(base >= n ? 0 : (n - base < chunk ? n - base : chunk)));
    }
    const int at = merge_sorted_runs(runKeys, runIdx, off, len, s);
    preKeys = runKeys[at]; preOrder = runIdx[at]; /* owned by the tree from here */
    pool_free(runKeys[at ^ 1], s); pool_free(runIdx[at ^ 1], s);
    pool_free(myKeys, s); pool_free(myIdx, s);
  }
  /* Output slabs: rows that come back in SFC order (several ranks, or one GPU with an index array) are copied
   * out slab by slab while the next slab's list kernels run, so the slab boundaries are looked up with the rank
   * boundaries: fine[r * S + j] = target r plus j/S of the way to target r + 1 */
  static const int slabsEnv = getenv("CB200_OUT_SLABS") ? atoi(getenv("CB200_OUT_SLABS")) : 0;
  const bool sfcOut = h_out && (world > 1 || h_index);
  /* about a million rows per slab, at most eight: the last slab's copy is the only one nothing hides, and a list
   * launch over fewer buckets loses to its own tail (a multistep step re-cuts by active particles: one slab) */
  int S = sfcOut && !multistep ? (slabsEnv > 0 ? slabsEnv : n / cutWorld / (1 << 20)) : 1;
  if (S < 1) S = 1;
  if (S > kOutSlabsMax) S = kOutSlabsMax;
  for (int r = 0; r < cutWorld; ++r)
    for (int j = 0; j < S; ++j)
      fine[r * S + j] = targets[r] + (int)(((long long)(targets[r + 1] - targets[r]) * j) / S);
  fine[cutWorld * S] = targets[cutWorld];
  /* Node arrays: 1.5 n nodes can never be exceeded by a tree of particles at distinct places, but a real tree has
   * about n / 4; the first step of a step object asks for 0.75 n, later ones for 1.3 x what the last tree had.  The
   * build reports an overflow; it is then repeated with the full size (rare: every rank sees the same tree, so all
   * ranks repeat together; the repeat sorts the whole box itself). */
  static const double capEnv = getenv("CB200_TREE_CAP_FACTOR") ? atof(getenv("CB200_TREE_CAP_FACTOR")) : 0.0;
  /* CB200_LEARN_SIZES=1 switches the learned sizes on.  Off by default: they cut the memory in use (clustered 512^3
   * on eight ranks 134 -> 97 GB per rank, 256^3 on one GPU 34 -> 27 GB) at the same resident step time, but on the
   * 512^3 box the end-to-end steps right after the sizes settle showed 280 ms stalls in the walk phase (the pools'
   * reserved totals jitter with the order the warps draw their chunks, a capacity that grows is a fresh multi-GB
   * block from the system once the first step's reservation has been returned): not measured enough to be default. */
  static const bool learnSizes = getenv("CB200_LEARN_SIZES") && atoi(getenv("CB200_LEARN_SIZES")) != 0;
  double capFactor = !learnSizes ? 1.5 : (st->treeCapFactor > 0.0 ? st->treeCapFactor : 0.75);
  if (capEnv > 0.0 && st->stepsRun == 0) capFactor = capEnv; /* tests: force the overflow path on the first step */
  build_tree_impl(in, n, cfg.maxBucket, cfg.rootlo, cfg.roothi, &st->tree, fine, cutWorld * S + 1, cuts, s, preKeys, preOrder,
                  capFactor);
  if (st->tree.error == 1 && capFactor < 1.5) {
    cb200_tree_free(&st->tree, s);
    build_tree_impl(in, n, cfg.maxBucket, cfg.rootlo, cfg.roothi, &st->tree, fine, cutWorld * S + 1, cuts, s, nullptr, nullptr, 1.5);
    res->treeRebuilt = 1;
  }
  if (!st->tree.error) {
    double f = 1.3 * (double)st->tree.numNodes / (double)n + 0.01;
    if (f > 1.5) f = 1.5;
    st->treeCapFactor = (st->stepsRun >= 2 && st->treeCapFactor > f) ? st->treeCapFactor : f; /* only grows once settled */
  }
  st->haveTree = true;
  cb200_tree &tr = st->tree;
  nvtx_pop();
  cudaChk(cudaEventRecord(st->ev[PH_TREE + 1], s));
  res->numNodes = tr.numNodes; res->numBuckets = tr.numBuckets; res->numLevels = tr.numLevels;
  if (tr.error) { res->error = 10 + tr.error; nvtx_pop(); return; }
  const int nn = tr.numNodes, nb = tr.numBuckets;

  /* my share: buckets [b0, b1) = particles [p0, p1), in S slabs */
  int slabB[kOutSlabsMax + 1], slabP[kOutSlabsMax + 1];
  for (int j = 0; j <= S; ++j) { slabB[j] = cuts[2 * (cutRank * S + j)]; slabP[j] = cuts[2 * (cutRank * S + j) + 1]; }
  if (cutRank == 0) { slabB[0] = 0; slabP[0] = 0; }
  if (cutRank == cutWorld - 1) { slabB[S] = nb; slabP[S] = n; }
  for (int j = 1; j <= S; ++j) { if (slabB[j] < slabB[j - 1]) slabB[j] = slabB[j - 1]; if (slabP[j] < slabP[j - 1]) slabP[j] = slabP[j - 1]; }
  int b0 = slabB[0], p0 = slabP[0], b1 = slabB[S], p1 = slabP[S];

  /* moments */
  nvtx_push("CUDA_SER_TREE moments");
  if (nn > st->nodeCap) {
    for (void *p : {(void *)st->d_mom64, (void *)st->d_pkMom, (void *)st->d_letFlag}) if (p) cudaChk(cudaFree(p));
    st->nodeCap = nn + nn / 16 + 1024;
    cudaChk(cudaMalloc((void **)&st->d_mom64, (size_t)st->nodeCap * 27 * sizeof(double)));
    cudaChk(cudaMalloc((void **)&st->d_pkMom, (size_t)st->nodeCap * sizeof(PackedCell)));
    st->d_letFlag = nullptr;
  }
  /* Locally essential build (let_kernels.cuh): below a block level only the subtrees near my buckets.  The block
   * level is the first with blocksPerRank x world nodes (blocks much smaller than a rank's domain keep the halo
   * thin) that still has three levels below it; the cover level the first with 64 x world nodes. */
  static const int letEnv = getenv("CB200_LET") ? atoi(getenv("CB200_LET")) : 1;
  static const int letBlocksPerRank = getenv("CB200_LET_BLOCKS_PER_RANK") ? atoi(getenv("CB200_LET_BLOCKS_PER_RANK")) : 32768;
  st->letLevel = -1;
  if (world > 1 && !multistep && letEnv && !st->letOff) {
    int Lb = -1, Ld = -1;
    for (int l = 0; l < tr.numLevels; ++l) {
      const long long cnt = tr.levelStart[l + 1] - tr.levelStart[l];
      if (Ld < 0 && cnt >= 64LL * world) Ld = l;
      if (Lb < 0 && cnt >= (long long)letBlocksPerRank * world) Lb = l;
    }
    /* not worth its set-up on a small tree (clustered 4 M particles on two ranks: 1.47 ms against 1.35 for the full
     * build); CB200_LET_BLOCKS_PER_RANK set = the caller wants it regardless (tests) */
    const bool bigEnough = nn >= 2000000 || getenv("CB200_LET_BLOCKS_PER_RANK") != nullptr;
    if (Lb >= 0 && Lb + 3 < tr.numLevels && bigEnough) {
      if (Ld < 0 || Ld > Lb) Ld = Lb;
      st->letLevel = Lb;
      if (!st->d_letFlag) cudaChk(cudaMalloc((void **)&st->d_letFlag, (size_t)st->nodeCap));
      if (!st->d_letCover) cudaChk(cudaMalloc((void **)&st->d_letCover, kLetCoverCap * sizeof(LetBox)));
      if (!st->d_letScalars) cudaChk(cudaMalloc((void **)&st->d_letScalars, 2 * sizeof(double)));
      cudaChk(cudaMemsetAsync(st->d_letScalars, 0, 2 * sizeof(double), s));
      int *nCover = reinterpret_cast<int *>(st->d_letScalars + 1);
      { /* largest particle softening */
        size_t bytes = 0;
        cudaChk(cub::DeviceReduce::Max(nullptr, bytes, tr.d_soft, st->d_letScalars, n, s));
        void *tmp = pool_alloc(bytes, s);
        cudaChk(cub::DeviceReduce::Max(tmp, bytes, tr.d_soft, st->d_letScalars, n, s));
        pool_free(tmp, s);
      }
      const int coverEnd = tr.levelStart[Ld + 1];
      let_cover_kernel<<<(coverEnd + 255) / 256, 256, 0, s>>>(tr.d_child0, tr.d_child1, tr.d_bucketFirst, tr.d_bucketCount,
                                                             tr.d_boxlo, tr.d_boxhi, tr.levelStart[Ld], coverEnd, b0, b1,
                                                             st->d_letCover, nCover);
      cudaChk(cudaPeekAtLastError());
      const int lo = tr.levelStart[Lb], cnt = tr.levelStart[Lb + 1] - lo;
      /* CB200_LET_RADIUS_SCALE (tests): shrinks the halo radius below its bound, so that the walk DOES leave the
       * built part and the fallback is exercised */
      static const double radiusScale = getenv("CB200_LET_RADIUS_SCALE") ? atof(getenv("CB200_LET_RADIUS_SCALE")) : 1.0;
      const double geom = radiusScale * 2.0 / sqrt(3.0) / cfg.theta;
      const int images = (cfg.nReplicas || cfg.ewald) ? (cfg.nReplicas > 0 ? cfg.nReplicas : 1) : 0;
      let_block_flags_kernel<<<(cnt + 127) / 128, 128, 0, s>>>(tr.d_bucketFirst, tr.d_bucketCount, tr.d_boxlo, tr.d_boxhi, lo, cnt, b0,
                                                              b1, st->d_letCover, nCover, radiusScale < 1.0 ? geom : (geom > 1.0 ? geom : 1.0),
                                                              st->d_letScalars, cfg.period, images, st->d_letFlag);
      cudaChk(cudaPeekAtLastError());
      for (int l = Lb; l + 1 < tr.numLevels; ++l) {
        const int llo = tr.levelStart[l], ln = tr.levelStart[l + 1] - llo;
        if (ln <= 0) continue;
        let_propagate_kernel<<<(ln + 255) / 256, 256, 0, s>>>(tr.d_child0, tr.d_child1, llo, ln, st->d_letFlag);
      }
      cudaChk(cudaPeekAtLastError());
      g_launches.fetch_add(3 + tr.numLevels - Lb);
    }
  }
  MomentLet let;
  if (st->letLevel >= 0) {
    let.blockLevel = st->letLevel;
    let.flag = st->d_letFlag;
    let.exchange = [&](double *work, size_t numNodes, int words, int lo, int cnt) {
      const size_t total = (size_t)words * cnt;
      unsigned long long *buf = (unsigned long long *)pool_alloc(total * sizeof(unsigned long long), s);
      let_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(work, numNodes, words, lo, cnt, tr.d_first, p0, p1, buf);
      cudaChk(cudaPeekAtLastError());
      ncclChk(nccl_api()->AllReduce(buf, buf, total, ncclUint64, ncclSum, st->comm->comm, s));
      let_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(work, numNodes, words, lo, cnt, buf);
      cudaChk(cudaPeekAtLastError());
      pool_free(buf, s);
      g_launches.fetch_add(2);
    };
  }
  /* the double records (walk, Ewald set-up) and the packed rows of the force kernels straight from the
   * build: no float AoS copy, no repack pass */
  build_moments_impl(tr.d_pos, tr.d_mass, tr.d_soft, tr.d_child0, tr.d_child1, tr.d_first, tr.d_last, tr.d_geolo,
                     tr.d_geohi, tr.d_boxlo, tr.d_boxhi, tr.levelStart, tr.numLevels, nn, nullptr, st->d_mom64, st->d_pkMom, s,
                     st->letLevel >= 0 ? &let : nullptr);
  cudaChk(cudaMemsetAsync(st->d_vars, 0, (size_t)n * sizeof(VariablePartData), s));
  nvtx_pop();
  cudaChk(cudaEventRecord(st->ev[PH_MOMENTS + 1], s));

  const unsigned char *bucketActive = nullptr;
  int nAct = n;
  if (multistep) {
    if (nb > st->bucketCap) {
      if (st->d_bucketActive) cudaChk(cudaFree(st->d_bucketActive));
      st->bucketCap = nb + nb / 16 + 1024;
      cudaChk(cudaMalloc((void **)&st->d_bucketActive, (size_t)st->bucketCap));
    }
    int counts[2] = {0, 0};
    cb200_active_sets_device(rungs, tr.d_order, n, tr.d_bucketStarts, tr.d_bucketSizes, nb, cfg.activeRung,
                             st->d_bucketActive, st->d_markers, counts, s);
    bucketActive = st->d_bucketActive;
    nAct = counts[1];
    res->activeBuckets = counts[0]; res->activeParticles = counts[1];
    if (world > 1 && nAct > 0) {
      /* equal ACTIVE particle counts (changa_b200.multigpu.bucket_range_by_active): the boundaries are the
       * particles holding markers r * nAct / world, snapped to bucket starts */
      int at[18];
      for (int r = 0; r <= world; ++r) { long long a = (long long)r * nAct / world; at[r] = (int)(a < nAct ? a : nAct - 1); }
      int *d_at = (int *)pool_alloc(18 * sizeof(int), s);
      int *d_cut = (int *)pool_alloc(36 * sizeof(int), s);
      cudaChk(cudaMemcpyAsync(d_at, at, (size_t)(world + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
      step_active_cuts_kernel<<<1, 32, 0, s>>>(st->d_markers, d_at, world + 1, tr.d_bucketStarts, nb, n, d_cut);
      cudaChk(cudaMemcpyAsync(cuts, d_cut, (size_t)(2 * world + 2) * sizeof(int), cudaMemcpyDeviceToHost, s));
      cudaChk(cudaStreamSynchronize(s));
      pool_free(d_at, s); pool_free(d_cut, s);
      b0 = rank == 0 ? 0 : cuts[2 * rank]; p0 = rank == 0 ? 0 : cuts[2 * rank + 1];
      b1 = rank == world - 1 ? nb : cuts[2 * rank + 2]; p1 = rank == world - 1 ? n : cuts[2 * rank + 3];
      slabB[0] = b0; slabB[1] = b1; slabP[0] = p0; slabP[1] = p1; /* S == 1 here */
    }
  }
  res->bucketLo = b0; res->bucketHi = b1; res->partLo = p0; res->partHi = p1;

  /* Ewald needs the particles and the root moments only: it runs on a second stream under the walk
   * (the walk is latency-bound, the Ewald kernel issue-bound; both leave room for the other) */
  const real fper = (real)((cfg.nReplicas || cfg.ewald) ? cfg.period : 0.0);
  cudaStream_t es = cfg.overlapEwald ? st->aux : s;
  if (cfg.ewald && p1 > p0) {
    nvtx_push("CUDA_EWALD");
    if (cfg.overlapEwald) {
      cudaChk(cudaEventRecord(st->evFork, s));
      cudaChk(cudaStreamWaitEvent(es, st->evFork, 0));
    }
    ewald_setup_kernel<<<1, 128, 0, es>>>(st->d_mom64, cfg.period, cfg.dEwCut, cfg.nReplicas, st->nEwh, st->d_hxyz, p0, p1 - 1,
                                          st->d_ewald);
    cudaChk(cudaPeekAtLastError());
    cudaChk(cudaMemcpyToSymbolAsync(c_ewaldSlot, st->d_ewald, sizeof(EwaldParams), (size_t)st->ewaldSlot * sizeof(EwaldParams),
                                    cudaMemcpyDeviceToDevice, es));
    const int *mk = nullptr;
    int first = p0, last = p1 - 1, count = p1 - p0;
    if (multistep) { /* large-phase form: the markers of my particle range (ascending: a binary search on the host copy is not needed) */
      int range[2] = {0, 0};
      int *d_range = (int *)pool_alloc(2 * sizeof(int), es);
      step_marker_range_kernel<<<1, 2, 0, es>>>(st->d_markers, nAct, p0, p1, d_range);
      cudaChk(cudaMemcpyAsync(range, d_range, sizeof range, cudaMemcpyDeviceToHost, es));
      cudaChk(cudaStreamSynchronize(es));
      pool_free(d_range, es);
      mk = st->d_markers + range[0];
      first = 0; count = range[1] - range[0]; last = count - 1;
    }
    if (count > 0) {
      TapScope tap(TAP_EWALD, es);
      ewald_slot_kernel<<<(count + kEwaldThreads - 1) / kEwaldThreads, kEwaldThreads, 0, es>>>(
          (const PackedPart *)tr.d_packedParts, st->d_vars, mk, first, last, st->ewaldSlot);
      cudaChk(cudaPeekAtLastError());
    }
    if (cfg.overlapEwald) cudaChk(cudaEventRecord(st->evJoin, es));
    nvtx_pop();
  }
  /* with the overlap the Ewald kernel's time is inside the walk phase (which ends after the join) */
  cudaChk(cudaEventRecord(st->ev[PH_EWALD + 1], s));

  /* interaction lists of my buckets */
  nvtx_push("CUDA_SER_LIST");
  WalkExtras extras;
  static const double hintScale = getenv("CB200_POOL_HINT_SCALE") ? atof(getenv("CB200_POOL_HINT_SCALE")) : 1.0; /* tests */
  for (int k = 0; k < 3; ++k) extras.poolHint[k] = learnSizes ? (unsigned long long)((double)st->poolHint[k] * hintScale) : 0ull;
  tl_walkExtras = &extras;
  if (st->letLevel >= 0) {
    extras.built = st->d_letFlag;
    extras.builtAlways = tr.levelStart[st->letLevel + 1]; /* every node down to the block level has a record */
    extras.markEnd = tr.levelStart[st->letLevel + 2 <= tr.numLevels ? st->letLevel + 2 : tr.numLevels];
    extras.reduceSoftMax = [&](unsigned long long *d_bits, cudaStream_t ws) {
      /* the largest node softening over the whole tree (a non-negative double orders like its bit pattern): every
       * node is built by the rank that owns it, so the maximum over the ranks is the single-GPU value */
      ncclChk(nccl_api()->AllReduce(d_bits, d_bits, 1, ncclUint64, ncclMax, st->comm->comm, ws));
    };
  }
  auto walk = [&]() {
    cb200_walk_device_active(nn, nb, tr.numLevels, tr.levelStart, tr.d_child0, tr.d_child1, tr.d_parent, tr.d_first, tr.d_last,
                             tr.d_bucketFirst, tr.d_bucketCount, tr.d_bucketNode, tr.d_boxlo, tr.d_boxhi, st->d_mom64, cfg.theta,
                             cfg.nReplicas, cfg.period, b0, b1, bucketActive, &st->lists, s);
  };
  walk();
  /* Two things can make a walk worth repeating, and with several ranks both are decided TOGETHER (the walk of a
   * locally essential step contains a collective, and the next step's exchange needs every rank): the pools,
   * sized from the last step, were too small on some rank -> once more with worst-case sizes; the walk met a node
   * outside the part a rank built (let_kernels.cuh: not expected) -> moments and walk once more on the full build,
   * and the locally essential build stays off for this step object. */
  auto agree2 = [&](bool a, bool b, bool &ra, bool &rb) { /* one blocking all-reduce (max) of two flags */
    double v[2] = {a ? 1.0 : 0.0, b ? 1.0 : 0.0};
    if (world > 1) cb200_comm_allreduce_f64(st->comm, v, 2, 1, s);
    ra = v[0] > 0.0; rb = v[1] > 0.0;
  };
  /* next step's capacities: what this walk reserved + 30 %, rounded coarsely; after the second step (when the
   * first step's worst-case reservation has been returned) they only grow, so that steady-state steps keep asking
   * for the same block sizes */
  auto learn_pools = [&]() {
    for (int k = 0; k < 3; ++k) {
      const unsigned long long need = round_up_coarse(extras.poolHint[k] + extras.poolHint[k] * 3 / 10 + (1u << 16));
      st->poolHint[k] = (st->stepsRun >= 2 && st->poolHint[k] > need) ? st->poolHint[k] : need;
    }
  };
  const bool hinted = learnSizes && st->poolHint[0] != 0;
  bool small = hinted && st->lists.error == 2, outside = st->letLevel >= 0 && st->lists.error == kWalkNotBuilt;
  if (world > 1 && (hinted || st->letLevel >= 0)) agree2(small, outside, small, outside);
  if (small) {
    cb200_lists_free(&st->lists, s);
    for (int k = 0; k < 3; ++k) extras.poolHint[k] = 0;
    walk();
    res->walkRepeated = 1;
    outside = st->letLevel >= 0 && st->lists.error == kWalkNotBuilt;
    if (world > 1 && st->letLevel >= 0) { bool dummy = false; agree2(outside, false, outside, dummy); }
  }
  learn_pools();
  if (outside) {
    fprintf(stderr, "changa_b200: rank %d: the walk left the locally essential tree (block level %d); full moment build from here on\n",
            rank, st->letLevel);
    st->letOff = true;
    st->letLevel = -1;
    res->letFallback = 1;
    cb200_lists_free(&st->lists, s);
    build_moments_impl(tr.d_pos, tr.d_mass, tr.d_soft, tr.d_child0, tr.d_child1, tr.d_first, tr.d_last, tr.d_geolo,
                       tr.d_geohi, tr.d_boxlo, tr.d_boxhi, tr.levelStart, tr.numLevels, nn, nullptr, st->d_mom64, st->d_pkMom, s);
    extras.built = nullptr; extras.reduceSoftMax = nullptr;
    for (int k = 0; k < 3; ++k) extras.poolHint[k] = 0; /* the full build visits no more nodes, but be generous */
    walk();
    learn_pools();
  }
  tl_walkExtras = nullptr;
  res->letBlockLevel = st->letLevel;
  st->haveLists = true;
  cb200_lists &li = st->lists;

  nvtx_pop();
  if (cfg.overlapEwald && cfg.ewald && p1 > p0) cudaChk(cudaStreamWaitEvent(s, st->evJoin, 0));
  cudaChk(cudaEventRecord(st->ev[PH_WALK + 1], s));
  res->nCell = li.nCell; res->nSoft = li.nSoft; res->nPart = li.nPart;
  if (li.error) { res->error = 20 + li.error; nvtx_pop(); return; }

  /* forces */
  const PackedPart *P = (const PackedPart *)tr.d_packedParts;
  unsigned *counter = stream_counter(s);
  /* the list kernels see only this rank's buckets [b0, b1): markers, starts and sizes are offset, the
   * marker VALUES still index the whole lists (a launch over all buckets made every warp draw and skip
   * the other ranks' empty buckets: 0.5 ms of a 3.7 ms p-p launch at two ranks) */
  const bool streamOut = sfcOut && p1 - p0 <= outCapacityRows;
  if (streamOut && p1 > p0) { /* the caller indices of my rows travel under the first slab's kernels */
    cudaChk(cudaEventRecord(st->evTree, s));
    cudaChk(cudaStreamWaitEvent(st->outStream, st->evTree, 0));
    cudaChk(cudaMemcpyAsync(h_index, tr.d_order + p0, (size_t)(p1 - p0) * sizeof(int), cudaMemcpyDeviceToHost, st->outStream));
    res->d2hBytes += (long long)(p1 - p0) * (long long)sizeof(int);
  }
  nvtx_push("CUDA_GRAV_LOCAL / CUDA_PART_GRAV_LOCAL");
  for (int j = 0; j < S; ++j) {
    const int bs = slabB[j], nS = slabB[j + 1] - slabB[j];
    cudaChk(cudaEventRecord(st->evSlab[j][0], s));
    if (nS > 0)
      dispatch_cell_list(cfg.maxBucket, P, st->d_vars, st->d_pkMom, li.d_cell, li.d_cellMarkers + bs, li.d_starts + bs,
                         li.d_sizes + bs, nS, fper, counter, s);
    cudaChk(cudaEventRecord(st->evSlab[j][1], s));
    if (nS > 0) {
      dispatch_part_list(cfg.maxBucket, P, st->d_vars, P, li.d_part, li.d_partMarkers + bs, li.d_starts + bs, li.d_sizes + bs,
                         nS, fper, counter, s);
      if (li.nSoft)
        dispatch_part_list(cfg.maxBucket, P, st->d_vars, (const PackedPart *)li.d_nodeParticles, li.d_soft, li.d_softMarkers + bs,
                           li.d_starts + bs, li.d_sizes + bs, nS, fper, counter, s);
    }
    cudaChk(cudaEventRecord(st->evSlab[j][2], s));
    if (streamOut && slabP[j + 1] > slabP[j]) { /* this slab's rows are final: out they go, under the next slab's kernels */
      cudaChk(cudaStreamWaitEvent(st->outStream, st->evSlab[j][2], 0));
      cudaChk(cudaMemcpyAsync((VariablePartData *)h_out + (slabP[j] - p0), st->d_vars + slabP[j],
                              (size_t)(slabP[j + 1] - slabP[j]) * sizeof(VariablePartData), cudaMemcpyDeviceToHost, st->outStream));
      res->d2hBytes += (long long)(slabP[j + 1] - slabP[j]) * (long long)sizeof(VariablePartData);
    }
  }
  nvtx_pop();
  cudaChk(cudaEventRecord(st->ev[PH_PC + 1], s));
  cudaChk(cudaEventRecord(st->ev[PH_PP + 1], s));

  /* pair counts (the metric) and results */
  nvtx_push("CUDA_XFER_BACK");
  cudaChk(cudaMemsetAsync(st->d_counts, 0, 64, s));
  if (b1 > b0) {
    step_pairs_kernel<<<device_info().sms, 256, 0, s>>>(li.d_cellMarkers, li.d_partMarkers, li.d_softMarkers, li.d_sizes, b0, b1,
                                                       st->d_counts);
    cudaChk(cudaPeekAtLastError());
  }
  unsigned long long pairs[2] = {0, 0};
  cudaChk(cudaMemcpyAsync(pairs, st->d_counts, sizeof pairs, cudaMemcpyDeviceToHost, s));
  if (sfcOut) {
    res->rows = p1 - p0;
    if (!streamOut) res->error = 30; /* the rank's share outgrew the caller's result buffer */
    cudaChk(cudaEventRecord(st->evOut, st->outStream));
    cudaChk(cudaStreamWaitEvent(s, st->evOut, 0));
  } else if (world == 1) {
    res->rows = n;
    if (h_out && n > outCapacityRows) {
      res->error = 30;
    } else if (h_out) { /* caller's order: one scatter on the device, one copy */
      step_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(st->d_vars, tr.d_order, n, st->d_out);
      cudaChk(cudaPeekAtLastError());
      cudaChk(cudaMemcpyAsync(h_out, st->d_out, (size_t)n * sizeof(VariablePartData), cudaMemcpyDeviceToHost, s));
      res->d2hBytes += (long long)n * (long long)sizeof(VariablePartData);
    }
  } else {
    res->rows = p1 - p0;
  }
  g_launches.fetch_add(3);
  nvtx_pop();
  cudaChk(cudaEventRecord(st->ev[PH_FINISH + 1], s));
  if (!keepLists) { cb200_lists_free(&st->lists, s); st->haveLists = false; }
  cudaChk(cudaStreamSynchronize(s));
  res->pcPairs = (long long)pairs[0]; res->ppPairs = (long long)pairs[1];
  res->cost = 198.0 * (double)pairs[0] + 30.0 * (double)pairs[1];
  for (int k = 0; k < PH_FINISH + 1; ++k) cudaChk(cudaEventElapsedTime(&res->ms[k], st->ev[k], st->ev[k + 1]));
  res->ms[PH_PC] = res->ms[PH_PP] = 0.0f;
  for (int j = 0; j < S; ++j) { /* the slabs' list kernels alternate: their times are summed per kind */
    float a = 0.0f, b = 0.0f;
    cudaChk(cudaEventElapsedTime(&a, st->evSlab[j][0], st->evSlab[j][1]));
    cudaChk(cudaEventElapsedTime(&b, st->evSlab[j][1], st->evSlab[j][2]));
    res->ms[PH_PC] += a; res->ms[PH_PP] += b;
  }
  cudaChk(cudaEventElapsedTime(&res->ms[PH_TOTAL], st->ev[0], st->ev[PH_FINISH + 1]));

  st->stepsRun += 1;
  /* the second step ran with sizes learned from the first: what the first one reserved beyond that goes back */
  if (learnSizes && st->stepsRun == 2) pool_trim_device();
  /* cost feedback for the next step's cuts: every rank learns every rank's cost and range */
  if (world > 1 && cfg.costCuts && !multistep) {
    std::vector<double> v(2 * (size_t)world, 0.0);
    v[rank] = res->cost; v[world + rank] = (double)p0;
    cb200_comm_allreduce_f64(st->comm, v.data(), 2 * world, 0, s);
    st->prevCost.assign(v.begin(), v.begin() + world);
    st->prevCut.resize((size_t)world + 1);
    for (int r = 0; r < world; ++r) st->prevCut[r] = (long long)v[world + r];
    st->prevCut[world] = n;
  }
  nvtx_pop();
}

/* rows a caller should provide for h_out / h_index: the whole box on one GPU; with several, three equal
 * shares (cost-balanced cuts give a rank in a void more particles than n / world) */
int cb200_step_out_capacity(const cb200_step *st) {
  const long long want = st->world == 1 ? st->n : 3LL * st->chunk + 64;
  return (int)(want < st->n ? want : st->n);
}

} /* extern "C" */
#endif
