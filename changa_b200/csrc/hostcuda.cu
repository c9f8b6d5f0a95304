/* hostcuda.cu -- the host side of the B200 gravity library: the entry points
 * ChaNGa's Charm++ code calls (include/changa_b200_api.h PART 1, same C++
 * signatures as the reference's HostCUDA.h:99-126 / EwaldCUDA.h:59-62), their
 * C-ABI twins (PART 2), and the runtime under them.
 *
 * What differs from the reference's HostCUDA.cu host code:
 *   - no cudaMalloc / cudaFree per request (HostCUDA.cu:578-613 does 4 + 4,
 *     cudaFree being device-synchronising): every transient buffer comes from
 *     the device's stream-ordered memory pool (release threshold = keep
 *     everything), one sub-allocated block per request, freed in stream order;
 *   - moments and particles are re-laid out on the device at upload time
 *     (device_layout.cuh) so the kernels can use 128-bit accesses;
 *   - Ewald constants travel as a kernel argument, not through process-global
 *     __constant__ symbols: the entry points are safe to call concurrently
 *     from several threads on different streams;
 *   - completion callbacks go to Charm++ HAPI when built inside ChaNGa
 *     (-DCB200_WITH_CHARM_HAPI) and to a registered C handler otherwise.
 *
 * There is no CPU fallback: without a CUDA device every entry point aborts
 * with the reference's "Fatal CUDA Error" message (HostCUDA.cu:39-47).
 */
#include <atomic>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/changa_b200_api.h"
#include "gravity_kernels.cuh"
#include "pp_stream_kernel.cuh"
#include "moments_build.cuh"
#include "walk_kernels.cuh"
#include "let_kernels.cuh"
#include "tree_kernels.cuh"
#include "ewald_setup.cuh"
#include <nvtx3/nvToolsExt.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_merge.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/reverse_iterator.h>

#ifdef CB200_WITH_CHARM_HAPI
#include "hapi.h"
#endif

using namespace cb200;

/* ------------------------------------------------------------- error policy */
#define cudaChk(code) cb200_cuda_assert((code), __FILE__, __LINE__)
static inline void cb200_cuda_assert(cudaError_t code, const char *file, int line) {
  if (code != cudaSuccess) {
    fprintf(stderr, "Fatal CUDA Error %s at %s:%d\n", cudaGetErrorString(code), file, line);
    abort();
  }
}

/* ------------------------------------------------------------ NVTX ranges */
/* host-side ranges around what each entry point enqueues, named after the reference's HAPI trace
 * ids (cuda_typedef.h:27-47): CUDA_XFER_LOCAL, CUDA_GRAV_LOCAL, ..., CUDA_EWALD, CUDA_SER_TREE /
 * CUDA_SER_LIST for the tree and list generation the device path does itself.  nvtx3 is header-only
 * and costs a null check when no tool is attached. */
static inline void nvtx_push(const char *name) { nvtxRangePushA(name); }
static inline void nvtx_pop() { nvtxRangePop(); }
struct NvtxScope {
  explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
  ~NvtxScope() { nvtxRangePop(); }
};

/* ------------------------------------------------------- completion callbacks */
static std::atomic<cb200_callback_fn> g_handler{nullptr};

#ifndef CB200_WITH_CHARM_HAPI
static void CUDART_CB cb200_trampoline(void *cb) {
  cb200_callback_fn h = g_handler.load(std::memory_order_acquire);
  if (h) h(cb);
}
/* stand-in for Charm++'s hapiAddCallback(stream, cb): run the registered
 * handler on a CUDA callback thread once the stream reaches this point */
static void hapiAddCallback(cudaStream_t stream, void *cb) {
  if (!cb || !g_handler.load(std::memory_order_acquire)) return;
  cudaChk(cudaLaunchHostFunc(stream, cb200_trampoline, cb));
}
#endif

/* --------------------------------------------------------- per-device state */
struct DeviceInfo {
  std::atomic<bool> ready{false};
  int sms = 0;
};
static DeviceInfo g_dev[64];
static std::mutex g_devMutex;

static const DeviceInfo &device_info() {
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  DeviceInfo &d = g_dev[dev];
  if (!d.ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(g_devMutex);
    if (!d.ready.load(std::memory_order_relaxed)) {
      cudaChk(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
      cudaMemPool_t pool;
      cudaChk(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t keep = UINT64_MAX; /* the arena: never hand memory back to the driver */
      cudaChk(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      /* a block freed on one stream is re-used on another only once that free has
       * completed: never by making the allocating (copy) stream wait for a kernel */
      int off = 0;
      cudaChk(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
      d.ready.store(true, std::memory_order_release);
    }
  }
  return d;
}

/* Exact-size block cache in front of the stream-ordered pool.  A force step asks for the same
 * sizes step after step (request blocks, tree arrays, walk pools: gigabytes at 10^7 particles);
 * handing every one of them back to cudaFreeAsync and asking again makes the driver re-stitch
 * virtual ranges, which shows up as random 20-200 ms stalls (measured on the 4 M box: 34 ms
 * steps with occasional 300 ms ones).  A block of >= 1 MiB freed on a stream is parked under
 * (device, stream, size) and handed to the next request of that size ON THAT STREAM -- stream order
 * makes that safe with no event -- so in steady state the large blocks never reach the driver.
 * The cache is capped at half of the device memory; over the cap it is flushed. */
struct BlockKey {
  int dev; cudaStream_t stream; size_t bytes;
  bool operator==(const BlockKey &o) const { return dev == o.dev && stream == o.stream && bytes == o.bytes; }
};
struct BlockKeyHash {
  size_t operator()(const BlockKey &k) const {
    return std::hash<size_t>()(k.bytes) ^ (std::hash<void *>()((void *)k.stream) * 1315423911u) ^ (size_t)k.dev;
  }
};
static std::mutex g_blockMutex;
static std::unordered_map<BlockKey, std::vector<void *>, BlockKeyHash> g_blockFree;
struct BlockInfo { size_t bytes; cudaStream_t stream; };
static std::unordered_map<void *, BlockInfo> g_blockSize; /* large blocks handed out by pool_alloc */
static size_t g_blockCached[64], g_blockCap[64]; /* per device (indices as g_dev) */
constexpr size_t kBlockCacheMin = 1u << 20;

static void block_cache_flush_locked(int dev, cudaStream_t only, bool matchStream) {
  for (auto it = g_blockFree.begin(); it != g_blockFree.end();) {
    if (it->first.dev == dev && (!matchStream || it->first.stream == only)) {
      for (void *p : it->second) {
        /* the remembered stream may be gone (a caller-owned stream destroyed without
         * cb200_stream_destroy) or about to be: cudaFree synchronises instead of naming it */
        (void)matchStream;
        cudaChk(cudaFree(p));
        g_blockCached[it->first.dev & 63] -= it->first.bytes;
      }
      it = g_blockFree.erase(it);
    } else {
      ++it;
    }
  }
}

/* sizes learned from a previous step jitter from step to step; rounded up to three significant bits (steps of at
 * most 12.5 %) they stay the same and the exact-size block cache keeps hitting */
static unsigned long long round_up_coarse(unsigned long long x) {
  if (x < 16) return x;
  int shift = 0;
  while ((x >> shift) >= 16) ++shift;
  const unsigned long long top = ((x + ((1ull << shift) - 1)) >> shift);
  return top << shift;
}
/* everything parked in the block cache of the current device goes back to the driver, and the driver's pool gives
 * unused memory back to the system: called once by a step object after its sizes have settled */
static void pool_trim_device() {
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  cudaChk(cudaDeviceSynchronize());
  {
    std::lock_guard<std::mutex> lock(g_blockMutex);
    block_cache_flush_locked(dev, nullptr, false);
  }
  cudaMemPool_t pool;
  cudaChk(cudaDeviceGetDefaultMemPool(&pool, dev));
  cudaChk(cudaMemPoolTrimTo(pool, 0));
}
static void *pool_alloc(size_t bytes, cudaStream_t stream) {
  void *p = nullptr;
  if (bytes == 0) return nullptr;
  device_info();
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  if (bytes >= kBlockCacheMin) {
    std::lock_guard<std::mutex> lock(g_blockMutex);
    auto it = g_blockFree.find(BlockKey{dev, stream, bytes});
    if (it != g_blockFree.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      g_blockCached[dev & 63] -= bytes;
      g_blockSize[p] = BlockInfo{bytes, stream};
      return p;
    }
  }
  cudaChk(cudaMallocAsync(&p, bytes, stream));
  {
    /* every block is recorded, small ones too: a caller may cudaFree() what we returned
     * (DataManager.cpp:992-996), and the address can come back for a block of another size */
    std::lock_guard<std::mutex> lock(g_blockMutex);
    g_blockSize[p] = BlockInfo{bytes, stream};
  }
  return p;
}
static void pool_free(void *p, cudaStream_t stream) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lock(g_blockMutex);
    auto it = g_blockSize.find(p);
    if (it != g_blockSize.end()) {
      const size_t bytes = it->second.bytes;
      /* a block allocated on one stream and released on another (a staging block: filled by the
       * copy stream, released by the kernel's stream) goes back to the driver pool, which
       * re-uses it only once this release has completed */
      const bool sameStream = it->second.stream == stream;
      g_blockSize.erase(it);
      if (!sameStream || bytes < kBlockCacheMin) goto to_driver;
      int dev = 0;
      cudaChk(cudaGetDevice(&dev));
      size_t &cached = g_blockCached[dev & 63], &cap = g_blockCap[dev & 63];
      if (cap == 0) {
        size_t freeB = 0, totalB = 0;
        cudaChk(cudaMemGetInfo(&freeB, &totalB));
        cap = totalB / 2;
      }
      if (cached + bytes > cap) block_cache_flush_locked(dev, nullptr, false);
      if (cached + bytes <= cap) {
        g_blockFree[BlockKey{dev, stream, bytes}].push_back(p);
        cached += bytes;
        return;
      }
    }
  }
to_driver:
  cudaChk(cudaFreeAsync(p, stream));
}
static void pool_forget_stream(cudaStream_t stream) {
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_blockMutex);
  block_cache_flush_locked(dev, stream, true);
}

/* One bucket counter per (device, stream) for the device-resident list launches: zeroed once when
 * it is made; every list kernel leaves it at zero again (grab_bucket), and launches on one stream
 * are ordered, so no memset node sits between the kernels of a step. */
struct CounterKey {
  int dev; cudaStream_t stream;
  bool operator==(const CounterKey &o) const { return dev == o.dev && stream == o.stream; }
};
struct CounterKeyHash {
  size_t operator()(const CounterKey &k) const { return std::hash<void *>()((void *)k.stream) * 31u + (size_t)k.dev; }
};
static std::mutex g_counterMutex;
static std::unordered_map<CounterKey, unsigned *, CounterKeyHash> g_streamCounter;
static unsigned *stream_counter(cudaStream_t stream) {
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_counterMutex);
  auto it = g_streamCounter.find({dev, stream});
  if (it != g_streamCounter.end()) return it->second;
  unsigned *p = nullptr;
  cudaChk(cudaMalloc((void **)&p, 256));
  cudaChk(cudaMemsetAsync(p, 0, 256, stream));
  g_streamCounter[{dev, stream}] = p;
  return p;
}
static void drop_stream_counter(cudaStream_t stream) {
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_counterMutex);
  auto it = g_streamCounter.find({dev, stream});
  if (it == g_streamCounter.end()) return;
  cudaFree(it->second);
  g_streamCounter.erase(it);
}

/* ------------------------------------------------------- copy/compute overlap */
/* A request arrives on ONE caller stream (TreePiece: streams[thisIndex % numStreams],
 * TreePiece.cpp:5380), and the reference enqueues its host->device list copy and
 * its kernel there back to back, so the copy engine idles while the SMs work and
 * vice versa.  Here every caller stream gets a private companion stream: the
 * request's staging copies go on the companion, an event hands them to the caller
 * stream, the kernel and the completion callback stay on the caller stream.  The
 * order the caller observes is unchanged (kernels and callbacks in submission
 * order; the pinned list buffers are released by the callback, which still fires
 * after the kernel that waited for the copy), but the copy of request i+1 runs
 * under the kernel of request i.  CB200_NO_COPY_STREAM=1 restores the serial order. */
struct Companion {
  cudaStream_t copy = nullptr;
  cudaEvent_t ev[16];
  unsigned next = 0;
};
static std::mutex g_compMutex;
static std::unordered_map<cudaStream_t, Companion *> g_comp;
static bool copy_stream_enabled() {
  static const bool on = getenv("CB200_NO_COPY_STREAM") == nullptr;
  return on;
}
static Companion *companion(cudaStream_t user) {
  std::lock_guard<std::mutex> lock(g_compMutex);
  auto it = g_comp.find(user);
  if (it != g_comp.end()) return it->second;
  Companion *c = new Companion;
  cudaChk(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  for (cudaEvent_t &e : c->ev) cudaChk(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  g_comp[user] = c;
  return c;
}
static void drop_companion(cudaStream_t user) {
  std::lock_guard<std::mutex> lock(g_compMutex);
  auto it = g_comp.find(user);
  if (it == g_comp.end()) return;
  pool_forget_stream(it->second->copy);
  cudaStreamDestroy(it->second->copy);
  for (cudaEvent_t &e : it->second->ev) cudaEventDestroy(e);
  delete it->second;
  g_comp.erase(it);
}
/* stream on which a request's staging copies are issued */
static cudaStream_t staging_stream(cudaStream_t user) { return copy_stream_enabled() ? companion(user)->copy : user; }
/* everything issued on the staging stream so far becomes visible to `user` */
static void staging_handoff(cudaStream_t user) {
  if (!copy_stream_enabled()) return;
  Companion *c = companion(user);
  cudaEvent_t e;
  {
    std::lock_guard<std::mutex> lock(g_compMutex);
    e = c->ev[c->next++ % 16];
  }
  cudaChk(cudaEventRecord(e, c->copy));
  cudaChk(cudaStreamWaitEvent(user, e, 0));
}

/* ------------------------------------------------------------- timing taps */
enum { TAP_CELL = 0, TAP_PART = 1, TAP_EWALD = 2, TAP_N = 3 };
struct Tap { cudaEvent_t a, b; int kind; };
static std::atomic<int> g_timing{0};
static std::mutex g_tapMutex;
static std::vector<Tap> g_taps;
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_kindLaunches[TAP_N];

struct TapScope {
  cudaStream_t stream;
  Tap tap;
  bool on;
  TapScope(int kind, cudaStream_t s) : stream(s), on(g_timing.load() != 0) {
    g_launches.fetch_add(1);
    g_kindLaunches[kind].fetch_add(1);
    if (!on) return;
    tap.kind = kind;
    cudaChk(cudaEventCreate(&tap.a));
    cudaChk(cudaEventCreate(&tap.b));
    cudaChk(cudaEventRecord(tap.a, stream));
  }
  ~TapScope() {
    if (!on) return;
    cudaChk(cudaEventRecord(tap.b, stream));
    std::lock_guard<std::mutex> lock(g_tapMutex);
    g_taps.push_back(tap);
  }
};

/* ---------------------------------------------------------- kernel launchers */
/* resident CTAs per SM of a list kernel, and the opt-in to its dynamic shared memory: both are
 * per-device facts (the attribute is set on the current device's copy of the function), so they are
 * cached per (kernel instantiation, device) -- a process may drive several GPUs (cb200_set_device) */
struct CtaCache { std::atomic<int> n[64]; };
template <typename K>
static int resident_ctas(CtaCache &cache, K kernel, size_t smem) {
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache.n[dev].load(std::memory_order_acquire);
  if (n > 0) return n;
  cudaChk(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaChk(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kListWarps * 32, smem));
  n = n > 0 ? n : 1;
  cache.n[dev].store(n, std::memory_order_release);
  return n;
}

static int list_grid(int nBuckets, int ctasPerSm) {
  int need = (nBuckets + kListWarps - 1) / kListWarps;
  int cap = device_info().sms * ctasPerSm; /* persistent: one wave, sized to the SM count */
  return need < cap ? need : cap;
}

template <int PB, int MINB, bool PAIR = false>
static void launch_cell_list(const PackedPart *parts, VariablePartData *vars, const PackedCell *cells,
                             const ILCell *list, const int *markers, const int *starts,
                             const int *sizes, int nBuckets, real fperiod, unsigned *counter,
                             cudaStream_t stream) {
  static CtaCache cache;
  const int ctas = resident_ctas(cache, cell_list_kernel<PB, MINB, PAIR>, cell_list_smem_bytes<PB>());
  cell_list_kernel<PB, MINB, PAIR><<<list_grid(nBuckets, ctas), kListWarps * 32, cell_list_smem_bytes<PB>(), stream>>>(
      parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter);
  cudaChk(cudaPeekAtLastError());
}

#ifndef CUDA_USE_DOUBLE
template <int PB, int MINB>
static void launch_cell_list_x2(const PackedPart *parts, VariablePartData *vars, const PackedCell *cells,
                                const ILCell *list, const int *markers, const int *starts,
                                const int *sizes, int nBuckets, real fperiod, unsigned *counter,
                                cudaStream_t stream) {
  static CtaCache cache;
  const int ctas = resident_ctas(cache, cell_list_x2_kernel<PB, MINB>, cell_list_x2_smem_bytes<PB>());
  cell_list_x2_kernel<PB, MINB><<<list_grid(nBuckets, ctas), kListWarps * 32, cell_list_x2_smem_bytes<PB>(), stream>>>(
      parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter);
  cudaChk(cudaPeekAtLastError());
}
#endif

template <int PB, int MINB>
static void launch_part_list(const PackedPart *parts, VariablePartData *vars, const PackedPart *sources,
                             const ILCell *list, const int *markers, const int *starts,
                             const int *sizes, int nBuckets, real fperiod, unsigned *counter,
                             cudaStream_t stream) {
  static CtaCache cache;
  const int ctas = resident_ctas(cache, part_list_kernel<PB, MINB>, part_list_smem_bytes<PB>());
  part_list_kernel<PB, MINB><<<list_grid(nBuckets, ctas), kListWarps * 32, part_list_smem_bytes<PB>(), stream>>>(
      parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter);
  cudaChk(cudaPeekAtLastError());
}

#ifndef CUDA_USE_DOUBLE
template <int PB, int MINB>
static void launch_part_list_x2(const PackedPart *parts, VariablePartData *vars, const PackedPart *sources,
                                const ILCell *list, const int *markers, const int *starts,
                                const int *sizes, int nBuckets, real fperiod, unsigned *counter,
                                cudaStream_t stream) {
  static CtaCache cache;
  const int ctas = resident_ctas(cache, part_list_x2_kernel<PB, MINB>, part_list_x2_smem_bytes<PB>());
  part_list_x2_kernel<PB, MINB><<<list_grid(nBuckets, ctas), kListWarps * 32, part_list_x2_smem_bytes<PB>(), stream>>>(
      parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter);
  cudaChk(cudaPeekAtLastError());
}
#endif

#ifndef CUDA_USE_DOUBLE
template <int PB, int MINB>
static void launch_part_list_stream(const PackedPart *parts, VariablePartData *vars, const PackedPart *sources,
                                    const ILCell *list, const int *markers, const int *starts,
                                    const int *sizes, int nBuckets, real fperiod, unsigned *counter,
                                    cudaStream_t stream) {
  static CtaCache cache;
  const int ctas = resident_ctas(cache, part_list_stream_kernel<PB, MINB>, part_list_stream_smem_bytes<PB>());
  part_list_stream_kernel<PB, MINB><<<list_grid(nBuckets, ctas), kListWarps * 32, part_list_stream_smem_bytes<PB>(), stream>>>(
      parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter);
  cudaChk(cudaPeekAtLastError());
}
#endif

/* maxBucket picks the register tile: targets per pass */
static void dispatch_cell_list(int maxBucket, const PackedPart *parts, VariablePartData *vars,
                               const PackedCell *cells, const ILCell *list, const int *markers,
                               const int *starts, const int *sizes, int nBuckets, real fperiod,
                               unsigned *counter, cudaStream_t stream) {
  TapScope tap(TAP_CELL, stream);
#ifdef CUDA_USE_DOUBLE
  /* doubles take two registers each: 8 targets per pass is what fits in 255 */
  static const int variant = getenv("CB200_PC64_VARIANT") ? atoi(getenv("CB200_PC64_VARIANT")) : 0; /* tuning switch */
  (void)maxBucket;
  if (variant == 1)
    launch_cell_list<6, 3>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else if (variant == 2)
    launch_cell_list<12, 2>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else if (variant == 3)
    launch_cell_list<4, 3>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else if (variant == 4) /* 4-6: two targets per basic block */
    launch_cell_list<4, 2, true>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else if (variant == 5)
    launch_cell_list<6, 2, true>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else if (variant == 6)
    launch_cell_list<8, 2, true>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else if (variant == 7) /* the round-1 default */
    launch_cell_list<8, 2>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  else /* measured (profiles/r02f_f64_pc_variants.log): 0.4225 ms on cube300 against 0.4308 for <8,2> */
    launch_cell_list<6, 2, true>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
#else
  static const bool scalar = getenv("CB200_PC_SCALAR") != nullptr; /* A/B switch: the pre-FFMA2 kernel */
  if (scalar) {
    if (maxBucket <= 8)
      launch_cell_list<8, 4>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    else if (maxBucket <= 12)
      launch_cell_list<12, 3>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    else
      launch_cell_list<16, 3>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  } else if (maxBucket <= 8) {
    launch_cell_list_x2<8, 3>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  } else {
    launch_cell_list_x2<12, 3>(parts, vars, cells, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  }
#endif
}

static void dispatch_part_list(int maxBucket, const PackedPart *parts, VariablePartData *vars,
                               const PackedPart *sources, const ILCell *list, const int *markers,
                               const int *starts, const int *sizes, int nBuckets, real fperiod,
                               unsigned *counter, cudaStream_t stream) {
  TapScope tap(TAP_PART, stream);
#ifdef CUDA_USE_DOUBLE
  (void)maxBucket;
  launch_part_list<8, 2>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
#else
  static const bool scalar = getenv("CB200_PP_SCALAR") != nullptr; /* A/B switch: the pre-FFMA2 kernel */
  if (scalar) {
    if (maxBucket <= 8)
      launch_part_list<8, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    else if (maxBucket <= 12)
      launch_part_list<12, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    else
      launch_part_list<16, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
  } else {
    /* CB200_PP_VARIANT: 0 = part_list_stream_kernel at 4 CTAs/SM (default), 3 = the same at 3 CTAs/SM;
     * 1/2 = the round-1 part_list_x2_kernel (A/B) */
    static const int variant = getenv("CB200_PP_VARIANT") ? atoi(getenv("CB200_PP_VARIANT")) : 0;
    if (variant == 1 || variant == 2) {
      if (maxBucket <= 8)
        launch_part_list_x2<8, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
      else if (variant == 1)
        launch_part_list_x2<12, 3>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
      else
        launch_part_list_x2<12, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    } else if (maxBucket <= 8) {
      launch_part_list_stream<8, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    } else if (variant == 3) {
      launch_part_list_stream<12, 3>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    } else {
      launch_part_list_stream<12, 4>(parts, vars, sources, list, markers, starts, sizes, nBuckets, fperiod, counter, stream);
    }
  }
#endif
}

static void repack_cells(const void *d_raw, PackedCell *d_out, int n, cudaStream_t stream) {
  if (n <= 0) return;
  repack_cells_kernel<<<(n + 127) / 128, 128, 0, stream>>>((const real *)d_raw, d_out, n);
  cudaChk(cudaPeekAtLastError());
  g_launches.fetch_add(1);
}
static void repack_parts(const void *d_raw, PackedPart *d_out, int n, cudaStream_t stream) {
  if (n <= 0) return;
  repack_parts_kernel<<<(n + 255) / 256, 256, 0, stream>>>((const real *)d_raw, d_out, n);
  cudaChk(cudaPeekAtLastError());
  g_launches.fetch_add(1);
}

/* upload caller records into pool scratch, rewrite them into `packed` */
static void upload_cells(const void *h_raw, size_t bytes, PackedCell *packed, cudaStream_t stream) {
  int n = (int)(bytes / sizeof(CudaMultipoleMoments));
  if (n == 0) return;
  void *raw = pool_alloc(bytes, stream);
  cudaChk(cudaMemcpyAsync(raw, h_raw, bytes, cudaMemcpyHostToDevice, stream));
  repack_cells(raw, packed, n, stream);
  pool_free(raw, stream);
}
static void upload_parts(const void *h_raw, size_t bytes, PackedPart *packed, cudaStream_t stream) {
  int n = (int)(bytes / sizeof(CompactPartData));
  if (n == 0) return;
  void *raw = pool_alloc(bytes, stream);
  cudaChk(cudaMemcpyAsync(raw, h_raw, bytes, cudaMemcpyHostToDevice, stream));
  repack_parts(raw, packed, n, stream);
  pool_free(raw, stream);
}

/* =================== PART 1: the reference's entry points ================== */

void allocatePinnedHostMemory(void **ptr, size_t size) {
  if (size <= 0) { /* HostCUDA.cu:67-74 */
    *ptr = NULL;
    fprintf(stderr, "allocatePinnedHostMemory: 0 size!\n");
    assert(0);
    return;
  }
#ifdef CB200_WITH_CHARM_HAPI
  hapiMallocHost(ptr, size, true);
#else
  cudaChk(cudaHostAlloc(ptr, size, cudaHostAllocPortable));
#endif
}

void freePinnedHostMemory(void *ptr) {
  if (ptr == NULL) { /* HostCUDA.cu:87-92 */
    fprintf(stderr, "freePinnedHostMemory: NULL ptr!\n");
    assert(0);
    return;
  }
#ifdef CB200_WITH_CHARM_HAPI
  hapiFreeHost(ptr, true);
#else
  cudaChk(cudaFreeHost(ptr));
#endif
}

void DataManagerTransferLocalTree(void *moments, size_t sMoments, void *compactParts,
                                  size_t sCompactParts, void *varParts, size_t sVarParts,
                                  void **d_localMoments, void **d_compactParts, void **d_varParts,
                                  cudaStream_t stream, int numParticles, void *callback) {
  NvtxScope range("CUDA_XFER_LOCAL");
  const size_t nCells = sMoments / sizeof(CudaMultipoleMoments);
  const size_t nParts = sCompactParts / sizeof(CompactPartData);
  *d_localMoments = pool_alloc(nCells * sizeof(PackedCell), stream);
  *d_compactParts = pool_alloc(nParts * sizeof(PackedPart), stream);
  *d_varParts = pool_alloc(sVarParts, stream);
  upload_cells(moments, sMoments, (PackedCell *)*d_localMoments, stream);
  upload_parts(compactParts, sCompactParts, (PackedPart *)*d_compactParts, stream);
  /* the accumulators start at zero whatever the host buffer holds
   * (ZeroVars, HostCUDA.cu:141-143); rows past numParticles keep the
   * caller's values like the reference's upload at :139 */
  size_t zeroed = (size_t)numParticles * sizeof(VariablePartData);
  if (zeroed > sVarParts) zeroed = sVarParts;
  if (zeroed) cudaChk(cudaMemsetAsync(*d_varParts, 0, zeroed, stream));
  if (sVarParts > zeroed)
    cudaChk(cudaMemcpyAsync((char *)*d_varParts + zeroed, (char *)varParts + zeroed, sVarParts - zeroed,
                            cudaMemcpyHostToDevice, stream));
  hapiAddCallback(stream, callback);
}

void DataManagerTransferRemoteChunk(void *moments, size_t sMoments, void *remoteParts,
                                    size_t sRemoteParts, void **d_remoteMoments,
                                    void **d_remoteParts, cudaStream_t stream, void *callback) {
  NvtxScope range("CUDA_XFER_REMOTE");
  const size_t nCells = sMoments / sizeof(CudaMultipoleMoments);
  const size_t nParts = sRemoteParts / sizeof(CompactPartData);
  *d_remoteMoments = pool_alloc(nCells * sizeof(PackedCell), stream);
  *d_remoteParts = pool_alloc(nParts * sizeof(PackedPart), stream);
  upload_cells(moments, sMoments, (PackedCell *)*d_remoteMoments, stream);
  upload_parts(remoteParts, sRemoteParts, (PackedPart *)*d_remoteParts, stream);
  hapiAddCallback(stream, callback);
}

void TransferParticleVarsBack(VariablePartData *hostBuffer, size_t size, void *d_varParts,
                              cudaStream_t stream, void *cb) {
  NvtxScope range("CUDA_XFER_BACK");
  if (size) cudaChk(cudaMemcpyAsync(hostBuffer, d_varParts, size, cudaMemcpyDeviceToHost, stream));
  hapiAddCallback(stream, cb);
}

/* one request = one pool block: [counter | markers | starts | sizes | list | missed] */
struct RequestScratch {
  char *base = nullptr;
  unsigned *counter = nullptr;
  int *markers = nullptr, *starts = nullptr, *sizes = nullptr;
  ILCell *list = nullptr;
  void *missedPacked = nullptr;
  void *missedRaw = nullptr;
};

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static int max_bucket_size(const int *sizes, int n) {
  int m = 0;
  for (int i = 0; i < n; ++i) m = sizes[i] > m ? sizes[i] : m;
  return m;
}

static RequestScratch stage_request(CudaRequest *data, size_t missedPackedBytes, size_t missedRawBytes) {
  cudaStream_t stream = staging_stream(data->stream);
  const int nb = data->numBucketsPlusOne - 1;
  const size_t sList = (size_t)data->numInteractions * sizeof(ILCell);
  const size_t sMark = (size_t)(nb + 1) * sizeof(int), sStart = (size_t)nb * sizeof(int);
  size_t off = 0;
  const size_t oCounter = off; off += 256;
  const size_t oMark = off; off += align_up(sMark);
  const size_t oStart = off; off += align_up(sStart);
  const size_t oSize = off; off += align_up(sStart);
  const size_t oList = off; off += align_up(sList);
  const size_t oMissP = off; off += align_up(missedPackedBytes);
  const size_t oMissR = off; off += align_up(missedRawBytes);
  RequestScratch s;
  s.base = (char *)pool_alloc(off, stream);
  s.counter = (unsigned *)(s.base + oCounter);
  s.markers = (int *)(s.base + oMark);
  s.starts = (int *)(s.base + oStart);
  s.sizes = (int *)(s.base + oSize);
  s.list = (ILCell *)(s.base + oList);
  s.missedPacked = s.base + oMissP;
  s.missedRaw = s.base + oMissR;
  cudaChk(cudaMemsetAsync(s.counter, 0, 256, stream));
  cudaChk(cudaMemcpyAsync(s.markers, data->bucketMarkers, sMark, cudaMemcpyHostToDevice, stream));
  if (sStart) {
    cudaChk(cudaMemcpyAsync(s.starts, data->bucketStarts, sStart, cudaMemcpyHostToDevice, stream));
    cudaChk(cudaMemcpyAsync(s.sizes, data->bucketSizes, sStart, cudaMemcpyHostToDevice, stream));
  }
  if (sList) cudaChk(cudaMemcpyAsync(s.list, data->list, sList, cudaMemcpyHostToDevice, stream));
  return s;
}

enum Gather { G_LOCAL, G_REMOTE, G_MISSED };

static void cell_list_request(CudaRequest *data, Gather g) {
  NvtxScope range(g == G_LOCAL ? "CUDA_GRAV_LOCAL" : (g == G_REMOTE ? "CUDA_GRAV_REMOTE" : "CUDA_REMOTE_RESUME"));
  cudaStream_t stream = data->stream;
  const int nb = data->numBucketsPlusOne - 1;
  if (nb <= 0) { hapiAddCallback(stream, data->cb); return; }
  const size_t nMissed = (g == G_MISSED) ? data->sMissed / sizeof(CudaMultipoleMoments) : 0;
  RequestScratch s = stage_request(data, nMissed * sizeof(PackedCell), (g == G_MISSED) ? data->sMissed : 0);
  const PackedCell *cells = (const PackedCell *)(g == G_LOCAL ? data->d_localMoments : data->d_remoteMoments);
  if (g == G_MISSED) /* moments travel with the request (HostCUDA.cu:296-342) */
    cudaChk(cudaMemcpyAsync(s.missedRaw, data->missedNodes, data->sMissed, cudaMemcpyHostToDevice,
                            staging_stream(stream)));
  staging_handoff(stream);
  if (g == G_MISSED) {
    repack_cells(s.missedRaw, (PackedCell *)s.missedPacked, (int)nMissed, stream);
    cells = (const PackedCell *)s.missedPacked;
  }
  dispatch_cell_list(max_bucket_size(data->bucketSizes, nb), (const PackedPart *)data->d_localParts,
                     data->d_localVars, cells, s.list, s.markers, s.starts, s.sizes, nb, data->fperiod,
                     s.counter, stream);
  pool_free(s.base, stream);
  hapiAddCallback(stream, data->cb);
}

static void part_list_request(CudaRequest *data, Gather g, const CompactPartData *h_small, int nSmall) {
  NvtxScope range(g == G_LOCAL ? "CUDA_PART_GRAV_LOCAL" : (g == G_REMOTE ? "CUDA_PART_GRAV_REMOTE"
                  : (h_small ? "CUDA_PART_GRAV_LOCAL_SMALL" : "CUDA_REMOTE_RESUME")));
  cudaStream_t stream = data->stream;
  const int nb = data->numBucketsPlusOne - 1;
  if (nb <= 0) { hapiAddCallback(stream, data->cb); return; }
  const void *h_extra = nullptr;
  size_t sExtra = 0;
  if (g == G_MISSED) { h_extra = h_small ? (const void *)h_small : data->missedParts;
                       sExtra = h_small ? (size_t)nSmall * sizeof(CompactPartData) : data->sMissed; }
  const size_t nExtra = sExtra / sizeof(CompactPartData);
  RequestScratch s = stage_request(data, nExtra * sizeof(PackedPart), sExtra);
  const PackedPart *src = (const PackedPart *)(g == G_LOCAL ? data->d_localParts : data->d_remoteParts);
  if (g == G_MISSED)
    cudaChk(cudaMemcpyAsync(s.missedRaw, h_extra, sExtra, cudaMemcpyHostToDevice, staging_stream(stream)));
  staging_handoff(stream);
  if (g == G_MISSED) {
    repack_parts(s.missedRaw, (PackedPart *)s.missedPacked, (int)nExtra, stream);
    src = (const PackedPart *)s.missedPacked;
  }
  dispatch_part_list(max_bucket_size(data->bucketSizes, nb), (const PackedPart *)data->d_localParts,
                     data->d_localVars, src, s.list, s.markers, s.starts, s.sizes, nb, data->fperiod,
                     s.counter, stream);
  pool_free(s.base, stream);
  hapiAddCallback(stream, data->cb);
}

void TreePieceCellListDataTransferLocal(CudaRequest *data) { cell_list_request(data, G_LOCAL); }
void TreePieceCellListDataTransferRemote(CudaRequest *data) { cell_list_request(data, G_REMOTE); }
void TreePieceCellListDataTransferRemoteResume(CudaRequest *data) { cell_list_request(data, G_MISSED); }

void TreePiecePartListDataTransferLocal(CudaRequest *data) { part_list_request(data, G_LOCAL, nullptr, 0); }
void TreePiecePartListDataTransferRemote(CudaRequest *data) { part_list_request(data, G_REMOTE, nullptr, 0); }
void TreePiecePartListDataTransferRemoteResume(CudaRequest *data) { part_list_request(data, G_MISSED, nullptr, 0); }
/* sources are an ad-hoc host array shipped with the request (HostCUDA.cu:345-408) */
void TreePiecePartListDataTransferLocalSmallPhase(CudaRequest *data, CompactPartData *parts, int len) {
  part_list_request(data, G_MISSED, parts, len);
}

/* CudaFunctions.h:7-8 -- kept for callers that stage a request themselves */
void TreePieceDataTransferBasic(CudaRequest *data, CudaDevPtr *ptr) {
  cudaStream_t stream = data->stream;
  const int nb = data->numBucketsPlusOne - 1;
  const size_t sList = (size_t)data->numInteractions * sizeof(ILCell);
  const size_t sMark = (size_t)(nb + 1) * sizeof(int), sStart = (size_t)nb * sizeof(int);
  ptr->d_list = pool_alloc(sList, stream);
  ptr->d_bucketMarkers = (int *)pool_alloc(sMark, stream);
  ptr->d_bucketStarts = (int *)pool_alloc(sStart, stream);
  ptr->d_bucketSizes = (int *)pool_alloc(sStart, stream);
  if (sList) cudaChk(cudaMemcpyAsync(ptr->d_list, data->list, sList, cudaMemcpyHostToDevice, stream));
  cudaChk(cudaMemcpyAsync(ptr->d_bucketMarkers, data->bucketMarkers, sMark, cudaMemcpyHostToDevice, stream));
  if (sStart) {
    cudaChk(cudaMemcpyAsync(ptr->d_bucketStarts, data->bucketStarts, sStart, cudaMemcpyHostToDevice, stream));
    cudaChk(cudaMemcpyAsync(ptr->d_bucketSizes, data->bucketSizes, sStart, cudaMemcpyHostToDevice, stream));
  }
}
void TreePieceDataTransferBasicCleanup(CudaDevPtr *ptr) {
  cudaChk(cudaFree(ptr->d_list));
  cudaChk(cudaFree(ptr->d_bucketMarkers));
  cudaChk(cudaFree(ptr->d_bucketStarts));
  cudaChk(cudaFree(ptr->d_bucketSizes));
}

/* ------------------------------------------------------------------ Ewald */
void EwaldHostMemorySetup(EwaldData *h_idata, int nParticles, int nEwhLoop, int largephase) {
  if (largephase)
    allocatePinnedHostMemory((void **)&(h_idata->EwaldMarkers), (size_t)nParticles * sizeof(int));
  else
    h_idata->EwaldMarkers = NULL;
  allocatePinnedHostMemory((void **)&(h_idata->ewt), (size_t)nEwhLoop * sizeof(EwtData));
  allocatePinnedHostMemory((void **)&(h_idata->cachedData), sizeof(EwaldReadOnlyData));
}

void EwaldHostMemoryFree(EwaldData *h_idata, int largephase) {
  if (largephase) freePinnedHostMemory(h_idata->EwaldMarkers);
  freePinnedHostMemory(h_idata->ewt);
  freePinnedHostMemory(h_idata->cachedData);
}

static void launch_ewald(const PackedPart *parts, VariablePartData *vars, const int *d_markers,
                         int first, int last, int n, const EwaldReadOnlyData *ro, const EwtData *ewt,
                         cudaStream_t stream) {
  if (n <= 0) return;
  assert(ro->nEwhLoop <= NEWH); /* HostCUDA.cu:1923 */
  EwaldParams P;
  memset(&P, 0, sizeof P);
  P.ro = *ro;
  memcpy(P.ewt, ewt, (size_t)ro->nEwhLoop * sizeof(EwtData));
  TapScope tap(TAP_EWALD, stream);
  ewald_kernel<<<(n + kEwaldThreads - 1) / kEwaldThreads, kEwaldThreads, 0, stream>>>(parts, vars, d_markers,
                                                                                    first, last, P);
  cudaChk(cudaPeekAtLastError());
}

void EwaldHost(CompactPartData *d_localParts, VariablePartData *d_localVars, EwaldData *h_idata,
               cudaStream_t stream, void *cb, int myIndex, int largephase) {
  NvtxScope range("CUDA_EWALD");
  (void)myIndex;
  const int n = h_idata->cachedData->n;
  int *d_markers = nullptr;
  if (largephase && n > 0) {
    cudaStream_t st = staging_stream(stream);
    d_markers = (int *)pool_alloc((size_t)n * sizeof(int), st);
    cudaChk(cudaMemcpyAsync(d_markers, h_idata->EwaldMarkers, (size_t)n * sizeof(int),
                            cudaMemcpyHostToDevice, st));
    staging_handoff(stream);
  }
  if (!largephase || d_markers)
    launch_ewald((const PackedPart *)d_localParts, d_localVars, d_markers, h_idata->EwaldRange[0],
                 h_idata->EwaldRange[1], n, h_idata->cachedData, h_idata->ewt, stream);
  pool_free(d_markers, stream);
  hapiAddCallback(stream, cb);
}

/* ============================ PART 2: the C ABI ============================ */
extern "C" {

int cb200_abi_version(void) { return 1; }
int cb200_real_bytes(void) { return (int)sizeof(cudatype); }
const char *cb200_build_info(void) {
#ifdef CUDA_USE_DOUBLE
  return "changa_b200 sm_100a real=f64 hexadecapole kernels: cell_list{8} part_list{8} ewald moments";
#else
  return "changa_b200 sm_100a real=f32 hexadecapole kernels: cell_list{8,12,16} part_list{8,12,16} ewald moments";
#endif
}

void cb200_set_callback_handler(cb200_callback_fn handler) { g_handler.store(handler, std::memory_order_release); }

void *cb200_stream_create(void) {
  cudaStream_t s;
  cudaChk(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  stream_counter(s); /* the bucket counter of this stream exists before anything can be captured on it */
  return (void *)s;
}
void cb200_stream_destroy(void *stream) {
  pool_forget_stream((cudaStream_t)stream);
  drop_companion((cudaStream_t)stream);
  drop_stream_counter((cudaStream_t)stream);
  cudaChk(cudaStreamDestroy((cudaStream_t)stream));
}
void cb200_stream_synchronize(void *stream) { cudaChk(cudaStreamSynchronize((cudaStream_t)stream)); }
void cb200_device_synchronize(void) { cudaChk(cudaDeviceSynchronize()); }
void cb200_set_device(int ordinal) { cudaChk(cudaSetDevice(ordinal)); }
void cb200_device_free(void *dptr) { cudaChk(cudaFree(dptr)); }

void cb200_allocatePinnedHostMemory(void **ptr, size_t size) { allocatePinnedHostMemory(ptr, size); }
void cb200_freePinnedHostMemory(void *ptr) { freePinnedHostMemory(ptr); }
void cb200_DataManagerTransferLocalTree(void *moments, size_t sMoments, void *compactParts,
                                        size_t sCompactParts, void *varParts, size_t sVarParts,
                                        void **d_localMoments, void **d_compactParts, void **d_varParts,
                                        void *stream, int numParticles, void *callback) {
  DataManagerTransferLocalTree(moments, sMoments, compactParts, sCompactParts, varParts, sVarParts,
                               d_localMoments, d_compactParts, d_varParts, (cudaStream_t)stream,
                               numParticles, callback);
}
void cb200_DataManagerTransferRemoteChunk(void *moments, size_t sMoments, void *compactParts,
                                          size_t sCompactParts, void **d_remoteMoments,
                                          void **d_remoteParts, void *stream, void *callback) {
  DataManagerTransferRemoteChunk(moments, sMoments, compactParts, sCompactParts, d_remoteMoments,
                                 d_remoteParts, (cudaStream_t)stream, callback);
}
void cb200_TransferParticleVarsBack(void *hostBuffer, size_t size, void *d_varParts, void *stream, void *cb) {
  TransferParticleVarsBack((VariablePartData *)hostBuffer, size, d_varParts, (cudaStream_t)stream, cb);
}
void cb200_TreePieceCellListDataTransferLocal(CudaRequest *d) { TreePieceCellListDataTransferLocal(d); }
void cb200_TreePieceCellListDataTransferRemote(CudaRequest *d) { TreePieceCellListDataTransferRemote(d); }
void cb200_TreePieceCellListDataTransferRemoteResume(CudaRequest *d) { TreePieceCellListDataTransferRemoteResume(d); }
void cb200_TreePiecePartListDataTransferLocal(CudaRequest *d) { TreePiecePartListDataTransferLocal(d); }
void cb200_TreePiecePartListDataTransferLocalSmallPhase(CudaRequest *d, CompactPartData *parts, int len) {
  TreePiecePartListDataTransferLocalSmallPhase(d, parts, len);
}
void cb200_TreePiecePartListDataTransferRemote(CudaRequest *d) { TreePiecePartListDataTransferRemote(d); }
void cb200_TreePiecePartListDataTransferRemoteResume(CudaRequest *d) { TreePiecePartListDataTransferRemoteResume(d); }
void cb200_EwaldHostMemorySetup(EwaldData *h, int size, int nEwhLoop, int largephase) {
  EwaldHostMemorySetup(h, size, nEwhLoop, largephase);
}
void cb200_EwaldHostMemoryFree(EwaldData *h, int largephase) { EwaldHostMemoryFree(h, largephase); }
void cb200_EwaldHost(void *d_localParts, void *d_localVars, EwaldData *h, void *stream, void *cb,
                     int myIndex, int largephase) {
  EwaldHost((CompactPartData *)d_localParts, (VariablePartData *)d_localVars, h, (cudaStream_t)stream, cb,
            myIndex, largephase);
}

/* ---- device-resident variants: lists already in HBM, no copy issued ---- */
static int device_max_bucket(const int *d_sizes, int n, int hint) {
  (void)d_sizes; (void)n;
  return hint > 0 ? hint : 16; /* unknown: the widest register tile handles any size */
}

void cb200_cell_list_device(void *d_parts, void *d_vars, void *d_moments, const ILCell *d_list,
                            const int *d_markers, const int *d_starts, const int *d_sizes,
                            int numBuckets, cudatype fperiod, void *stream) {
  cb200_cell_list_device_ex(d_parts, d_vars, d_moments, d_list, d_markers, d_starts, d_sizes, numBuckets,
                            fperiod, 0, stream);
}
void cb200_cell_list_device_ex(void *d_parts, void *d_vars, void *d_moments, const ILCell *d_list,
                               const int *d_markers, const int *d_starts, const int *d_sizes,
                               int numBuckets, cudatype fperiod, int maxBucketSize, void *stream) {
  if (numBuckets <= 0) return;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned *counter = stream_counter(s);
  dispatch_cell_list(device_max_bucket(d_sizes, numBuckets, maxBucketSize), (const PackedPart *)d_parts,
                     (VariablePartData *)d_vars, (const PackedCell *)d_moments, d_list, d_markers, d_starts,
                     d_sizes, numBuckets, fperiod, counter, s);
}
void cb200_part_list_device(void *d_parts, void *d_vars, void *d_sources, const ILCell *d_list,
                            const int *d_markers, const int *d_starts, const int *d_sizes,
                            int numBuckets, cudatype fperiod, void *stream) {
  cb200_part_list_device_ex(d_parts, d_vars, d_sources, d_list, d_markers, d_starts, d_sizes, numBuckets,
                            fperiod, 0, stream);
}
void cb200_part_list_device_ex(void *d_parts, void *d_vars, void *d_sources, const ILCell *d_list,
                               const int *d_markers, const int *d_starts, const int *d_sizes,
                               int numBuckets, cudatype fperiod, int maxBucketSize, void *stream) {
  if (numBuckets <= 0) return;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned *counter = stream_counter(s);
  dispatch_part_list(device_max_bucket(d_sizes, numBuckets, maxBucketSize), (const PackedPart *)d_parts,
                     (VariablePartData *)d_vars, (const PackedPart *)d_sources, d_list, d_markers, d_starts,
                     d_sizes, numBuckets, fperiod, counter, s);
}
void cb200_ewald_device(void *d_parts, void *d_vars, const int *d_markers, int nActive,
                        const EwaldReadOnlyData *h_ro, const EwtData *h_ewt, void *stream) {
  launch_ewald((const PackedPart *)d_parts, (VariablePartData *)d_vars, d_markers, 0, nActive - 1, nActive,
               h_ro, h_ewt, (cudaStream_t)stream);
}

/* layout conversion for callers that keep their own device arrays (multi-GPU
 * driver: records arrive by all-gather, not from the host) */
void cb200_pack_moments_device(const void *d_raw, void *d_packed, int n, void *stream) {
  repack_cells(d_raw, (PackedCell *)d_packed, n, (cudaStream_t)stream);
}
void cb200_pack_particles_device(const void *d_raw, void *d_packed, int n, void *stream) {
  repack_parts(d_raw, (PackedPart *)d_packed, n, (cudaStream_t)stream);
}
void cb200_zero_vars_device(void *d_vars, int n, void *stream) {
  if (n > 0) cudaChk(cudaMemsetAsync(d_vars, 0, (size_t)n * sizeof(VariablePartData), (cudaStream_t)stream));
}
void cb200_copy_device(void *dst, const void *src, size_t bytes, void *stream) {
  if (bytes) cudaChk(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
}
size_t cb200_packed_moment_bytes(void) { return sizeof(PackedCell); }
size_t cb200_packed_particle_bytes(void) { return sizeof(PackedPart); }

/* ---- timing taps ---- */
void cb200_timing_enable(int on) { g_timing.store(on); }
void cb200_timing_reset(void) {
  std::lock_guard<std::mutex> lock(g_tapMutex);
  for (Tap &t : g_taps) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  g_taps.clear();
  for (int k = 0; k < TAP_N; ++k) g_kindLaunches[k].store(0);
}
void cb200_timing_read(double out[6]) {
  std::lock_guard<std::mutex> lock(g_tapMutex);
  for (int k = 0; k < 6; ++k) out[k] = 0.0;
  for (Tap &t : g_taps) {
    float ms = 0.f;
    cudaChk(cudaEventSynchronize(t.b));
    cudaChk(cudaEventElapsedTime(&ms, t.a, t.b));
    out[t.kind] += ms;
  }
  for (int k = 0; k < TAP_N; ++k) out[3 + k] = (double)g_kindLaunches[k].load();
}
long long cb200_kernel_launches(void) { return g_launches.load(); }

/* ---- device moment build ---- */
/* the locally essential build of a multi-GPU step (let_kernels.cuh); exchange() is called once, between the
 * build and the export of the block level, with the component-major work records */
struct MomentLet {
  int blockLevel;
  const unsigned char *flag; /* per node: 1 = built by this rank (meaningful at and below the block level) */
  std::function<void(double *work, size_t numNodes, int words, int lo, int n)> exchange;
};
static void build_moments_impl(const double *d_pos_xyz, const double *d_mass, const double *d_soft, const int *d_child0,
                               const int *d_child1, const int *d_firstPart, const int *d_lastPart,
                               const double *d_geolo_xyz, const double *d_geohi_xyz, const double *d_boxlo_xyz,
                               const double *d_boxhi_xyz, const int *h_levelStart, int numLevels, int numNodes,
                               real *d_moments_out, double *d_moments_f64_out, PackedCell *d_packed_out, cudaStream_t s,
                               const MomentLet *let = nullptr) {
  if (numNodes <= 0) return;
  MomentNode *work = (MomentNode *)pool_alloc((size_t)numNodes * sizeof(MomentNode), s);
  /* CB200_MOM_VARIANT: resident CTAs (of 64 threads) per SM the kernel is compiled for -- default 8
   * (128 registers, some spills: 0.886 ms at 4 M particles), 4: no bound (192 registers, no spills:
   * 1.043 ms), 3: 5 (168: 0.915), 2: 10 (96: 1.039) -- the FP64 chains want warps more than registers */
  static const int variant = getenv("CB200_MOM_VARIANT") ? atoi(getenv("CB200_MOM_VARIANT")) : 0;
  auto launch = [&](int lo, int n, const unsigned char *flag, int mode) {
#define CB200_MOM_LAUNCH(MINB)                                                                                     \
  build_moments_level_kernel<MINB><<<(n + kMomThreads - 1) / kMomThreads, kMomThreads, 0, s>>>(                     \
      d_pos_xyz, d_mass, d_soft, d_child0, d_child1, d_firstPart, d_lastPart, d_geolo_xyz, d_geohi_xyz, d_boxlo_xyz, \
      d_boxhi_xyz, lo, n, numNodes, work, d_moments_out, d_moments_f64_out, d_packed_out, flag, mode)
    if (variant == 4) CB200_MOM_LAUNCH(1);
    else if (variant == 2) CB200_MOM_LAUNCH(10);
    else if (variant == 3) CB200_MOM_LAUNCH(5);
    else CB200_MOM_LAUNCH(8);
#undef CB200_MOM_LAUNCH
    cudaChk(cudaPeekAtLastError());
    g_launches.fetch_add(1);
  };
  for (int lvl = numLevels - 1; lvl >= 0; --lvl) { /* bottom-up: children are on deeper levels */
    const int lo = h_levelStart[lvl], n = h_levelStart[lvl + 1] - lo;
    if (n <= 0) continue;
    if (let && lvl > let->blockLevel) {
      launch(lo, n, let->flag, 0);
    } else if (let && lvl == let->blockLevel) {
      launch(lo, n, let->flag, 1); /* my blocks (own and halo) */
      let->exchange(reinterpret_cast<double *>(work), (size_t)numNodes, (int)(sizeof(MomentNode) / sizeof(double)), lo, n);
      launch(lo, n, nullptr, 2);   /* every block of the level, from the exchanged records */
    } else {
      launch(lo, n, nullptr, 0);
    }
  }
  pool_free(work, s);
}

void cb200_build_moments(const double *d_pos_xyz, const double *d_mass, const double *d_soft,
                         int numParticles, const int *d_child0, const int *d_child1,
                         const int *d_firstPart, const int *d_lastPart, const double *d_geolo_xyz,
                         const double *d_geohi_xyz, const double *d_boxlo_xyz, const double *d_boxhi_xyz,
                         const int *h_levelStart, int numLevels, int numNodes, void *d_moments_out,
                         double *d_moments_f64_out, void *stream) {
  (void)numParticles;
  build_moments_impl(d_pos_xyz, d_mass, d_soft, d_child0, d_child1, d_firstPart, d_lastPart, d_geolo_xyz, d_geohi_xyz,
                     d_boxlo_xyz, d_boxhi_xyz, h_levelStart, numLevels, numNodes, (real *)d_moments_out, d_moments_f64_out,
                     nullptr, (cudaStream_t)stream);
}

/* ---- tree topology on the device ---- */
static int tree_level_grid() {
  static CtaCache cache;
  int dev = 0;
  cudaChk(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache.n[dev].load(std::memory_order_acquire);
  if (n > 0) return n;
  int per = 0;
  cudaChk(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, tree_levels_kernel, kTreeLevelThreads, 0));
  n = device_info().sms * (per > 4 ? 4 : (per > 0 ? per : 1)); /* co-resident by construction: cooperative launch */
  cache.n[dev].store(n, std::memory_order_release);
  return n;
}

/* Rank boundaries on the sorted particle array: boundary r sits at the first bucket that starts at or
 * after particle target[r] (never splits a bucket; changa_b200.multigpu.bucket_range_by_starts is the
 * host statement).  rank[p] = number of bucket starts in [0, p). */
__global__ void tree_cuts_kernel(const int *__restrict__ flag, const int *__restrict__ rank, int n, const int *__restrict__ targets,
                                 int nCuts, TreeMeta *meta) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nCuts) return;
  int p = targets[r];
  if (p > n) p = n;
  while (p < n && !flag[p]) ++p; /* at most maxBucket steps */
  meta->cuts[2 * r] = rank[p];
  meta->cuts[2 * r + 1] = p;
}
__global__ void tree_set_buckets_kernel(const int *__restrict__ rank, int n, TreeMeta *meta) { meta->numBuckets = rank[n]; }

/* the build; input as three arrays or one record array (TreeInput).  h_targets (optional, nCuts <= 17 particle
 * indices, host): rank boundaries looked up on the device and returned in cuts[2r] (bucket) / cuts[2r+1]
 * (particle).  ONE stream synchronisation (the node / bucket / level counts size everything downstream). */
/* keys of rows [0, n) of `in` and their stable sort (ties keep the row order, as the host comparator): keysOut /
 * orderOut get the sorted keys and the rows' indices idxBase + row */
static void tree_sorted_keys(const TreeInput &in, int n, const double *rootlo, const double *roothi, int idxBase,
                             unsigned long long *keysOut, int *orderOut, cudaStream_t s) {
  if (n <= 0) return;
  TreeBox box;
  for (int d = 0; d < 3; ++d) { box.lo[d] = rootlo[d]; box.inv[d] = 1.0 / (roothi[d] - rootlo[d]); }
  unsigned long long *keysIn = (unsigned long long *)pool_alloc((size_t)n * 8, s);
  int *idxIn = (int *)pool_alloc((size_t)n * 4, s);
  tree_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(in, n, box, keysIn, idxIn, idxBase);
  cudaChk(cudaPeekAtLastError());
  size_t tmpBytes = 0;
  cudaChk(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysIn, keysOut, idxIn, orderOut, n, 0, kTreeKeyBits, s));
  void *tmp = pool_alloc(tmpBytes, s);
  cudaChk(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysIn, keysOut, idxIn, orderOut, n, 0, kTreeKeyBits, s));
  pool_free(tmp, s); pool_free(keysIn, s); pool_free(idxIn, s);
  g_launches.fetch_add(3);
}

/* Stable merge of sorted runs into one: run r = (keys + off[r], idx + off[r], len[r]), runs in rank order, so equal
 * keys end up in the caller's order exactly as one stable sort of the whole box leaves them.  Pairwise rounds of
 * cub::DeviceMerge between two buffer pairs; returns 0 / 1: the pair that holds the n merged entries from offset 0. */
static int merge_sorted_runs(unsigned long long *keys[2], int *idx[2], std::vector<long long> off, std::vector<int> len,
                             cudaStream_t s) {
  int cur = 0;
  size_t maxTmp = 0;
  void *tmp = nullptr;
  bool compact = false; /* the first round also closes the gaps between the padded slices */
  while (len.size() > 1 || !compact) {
    std::vector<long long> noff;
    std::vector<int> nlen;
    long long at = 0;
    for (size_t r = 0; r < len.size(); r += 2) {
      if (r + 1 < len.size()) {
        size_t bytes = 0;
        cudaChk(cub::DeviceMerge::MergePairs(nullptr, bytes, keys[cur] + off[r], idx[cur] + off[r], len[r], keys[cur] + off[r + 1],
                                             idx[cur] + off[r + 1], len[r + 1], keys[cur ^ 1] + at, idx[cur ^ 1] + at,
                                             ::cuda::std::less<>{}, s));
        if (bytes > maxTmp) { pool_free(tmp, s); tmp = pool_alloc(bytes, s); maxTmp = bytes; }
        cudaChk(cub::DeviceMerge::MergePairs(tmp, bytes, keys[cur] + off[r], idx[cur] + off[r], len[r], keys[cur] + off[r + 1],
                                             idx[cur] + off[r + 1], len[r + 1], keys[cur ^ 1] + at, idx[cur ^ 1] + at,
                                             ::cuda::std::less<>{}, s));
        noff.push_back(at); nlen.push_back(len[r] + len[r + 1]);
        at += len[r] + len[r + 1];
      } else {
        if (len[r] > 0) {
          cudaChk(cudaMemcpyAsync(keys[cur ^ 1] + at, keys[cur] + off[r], (size_t)len[r] * 8, cudaMemcpyDeviceToDevice, s));
          cudaChk(cudaMemcpyAsync(idx[cur ^ 1] + at, idx[cur] + off[r], (size_t)len[r] * 4, cudaMemcpyDeviceToDevice, s));
        }
        noff.push_back(at); nlen.push_back(len[r]);
        at += len[r];
      }
      g_launches.fetch_add(1);
    }
    off.swap(noff); len.swap(nlen);
    cur ^= 1;
    compact = true;
  }
  pool_free(tmp, s);
  return cur;
}

/* preKeys / preOrder (or NULL): the sorted keys and the sorted rows' caller indices, made elsewhere (the multi-GPU
 * step sorts every rank's slice on its own GPU and merges the runs); both pool blocks of stream s, owned from here */
static void build_tree_impl(const TreeInput &in, int n, int maxBucket, const double *rootlo, const double *roothi,
                            cb200_tree *out, const int *h_targets, int nCuts, int *h_cuts, cudaStream_t s,
                            unsigned long long *preKeys = nullptr, int *preOrder = nullptr, double capFactor = 1.5) {
  memset(out, 0, sizeof *out);
  out->numParticles = n;
  if (n <= 0) return;
  TreeBox box;
  for (int d = 0; d < 3; ++d) { box.lo[d] = rootlo[d]; box.inv[d] = 1.0 / (roothi[d] - rootlo[d]); }
  const int tb = 256;
  /* keys, stable sort by key (ties keep the caller's order), sorted particle arrays */
  unsigned long long *keys = preKeys;
  out->d_order = preOrder;
  if (!keys) {
    keys = (unsigned long long *)pool_alloc((size_t)n * 8, s);
    out->d_order = (int *)pool_alloc((size_t)n * 4, s);
    tree_sorted_keys(in, n, rootlo, roothi, 0, keys, out->d_order, s);
  }
  out->d_pos = (double *)pool_alloc((size_t)n * 24, s);
  out->d_mass = (double *)pool_alloc((size_t)n * 8, s);
  out->d_soft = (double *)pool_alloc((size_t)n * 8, s);
  out->d_packedParts = pool_alloc((size_t)n * sizeof(PackedPart), s);
  tree_gather_kernel<<<(n + tb - 1) / tb, tb, 0, s>>>(in, out->d_order, n, out->d_pos, out->d_mass, out->d_soft,
                                                      (PackedPart *)out->d_packedParts);
  cudaChk(cudaPeekAtLastError());
  g_launches.fetch_add(2);

  /* nodes, level by level inside one cooperative kernel; capacity: a node holds at least one
   * particle, chains of single children are the only way past ~n/3 nodes */
  /* the kernel reports an overflow (error 1): the caller may retry with more */
  const int cap = (int)round_up_coarse((unsigned long long)((double)n * capFactor) + 4096ull);
  TreeArrays t;
  t.child0 = out->d_child0 = (int *)pool_alloc((size_t)cap * 4, s);
  t.child1 = out->d_child1 = (int *)pool_alloc((size_t)cap * 4, s);
  t.parent = out->d_parent = (int *)pool_alloc((size_t)cap * 4, s);
  t.first = out->d_first = (int *)pool_alloc((size_t)cap * 4, s);
  t.last = out->d_last = (int *)pool_alloc((size_t)cap * 4, s);
  t.geolo = out->d_geolo = (double *)pool_alloc((size_t)cap * 24, s);
  t.geohi = out->d_geohi = (double *)pool_alloc((size_t)cap * 24, s);
  TreeMeta *meta = (TreeMeta *)pool_alloc(sizeof(TreeMeta), s);
  cudaChk(cudaMemsetAsync(meta, 0, sizeof(TreeMeta), s));
  tree_root_kernel<<<1, 1, 0, s>>>(t, n, box, roothi[0], roothi[1], roothi[2]);
  int *split = (int *)pool_alloc((size_t)(n + 1) * 4, s); /* a level has at most n nodes */
  int *slot = (int *)pool_alloc((size_t)(n + 1) * 4, s);
  const int grid = tree_level_grid();
  int *blockSums = (int *)pool_alloc((size_t)grid * 4, s);
  {
    const unsigned long long *ckeys = keys;
    int mb = maxBucket, cp = cap;
    void *args[] = {&t, &ckeys, &mb, &cp, &split, &slot, &blockSums, &meta};
    cudaChk(cudaLaunchCooperativeKernel((void *)tree_levels_kernel, dim3(grid), dim3(kTreeLevelThreads), args, 0, s));
  }
  /* buckets in particle order: flags at bucket starts, their exclusive scan */
  int *flag = (int *)pool_alloc((size_t)(n + 1) * 4, s);
  int *rank = (int *)pool_alloc((size_t)(n + 1) * 4, s);
  int *leafAt = (int *)pool_alloc((size_t)n * 4, s);
  cudaChk(cudaMemsetAsync(flag, 0, (size_t)(n + 1) * 4, s));
  tree_leaf_flags_kernel<<<(cap + tb - 1) / tb, tb, 0, s>>>(t, meta, flag, leafAt); /* node count read on the device */
  size_t scanBytes = 0;
  cudaChk(cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, flag, rank, n + 1, s));
  void *scanTmp = pool_alloc(scanBytes, s);
  cudaChk(cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, flag, rank, n + 1, s));
  tree_set_buckets_kernel<<<1, 1, 0, s>>>(rank, n, meta);
  int *d_targets = nullptr;
  if (nCuts > 0) {
    d_targets = (int *)pool_alloc((size_t)nCuts * 4, s);
    cudaChk(cudaMemcpyAsync(d_targets, h_targets, (size_t)nCuts * 4, cudaMemcpyHostToDevice, s));
    tree_cuts_kernel<<<1, 160, 0, s>>>(flag, rank, n, d_targets, nCuts, meta); /* nCuts <= kTreeMaxCuts */
  }
  g_launches.fetch_add(6);
  TreeMeta hm;
  cudaChk(cudaMemcpyAsync(&hm, meta, sizeof hm, cudaMemcpyDeviceToHost, s));
  cudaChk(cudaStreamSynchronize(s)); /* the only one: node, bucket and level counts size what follows */
  out->numLevels = hm.numLevels;
  out->numNodes = hm.numNodes;
  out->numBuckets = hm.numBuckets;
  out->error = hm.error;
  for (int l = 0; l <= hm.numLevels && l < 66; ++l) out->levelStart[l] = hm.levelStart[l];
  for (int r = 0; r < 2 * nCuts; ++r) h_cuts[r] = hm.cuts[r];
  const int nn = hm.numNodes, nb = hm.numBuckets;
  if (!out->error) {
    out->d_bucketNode = (int *)pool_alloc((size_t)nb * 4, s);
    out->d_bucketStarts = (int *)pool_alloc((size_t)nb * 4, s);
    out->d_bucketSizes = (int *)pool_alloc((size_t)nb * 4, s);
    out->d_bucketFirst = (int *)pool_alloc((size_t)nn * 4, s);
    out->d_bucketCount = (int *)pool_alloc((size_t)nn * 4, s);
    const int m = nn > n ? nn : n;
    tree_buckets_kernel<<<(m + tb - 1) / tb, tb, 0, s>>>(t, nn, n, flag, rank, leafAt, out->d_bucketNode,
                                                       out->d_bucketFirst, out->d_bucketCount, out->d_bucketStarts,
                                                       out->d_bucketSizes);
    cudaChk(cudaPeekAtLastError());
    /* tight bounding boxes, bottom-up */
    out->d_boxlo = (double *)pool_alloc((size_t)nn * 24, s);
    out->d_boxhi = (double *)pool_alloc((size_t)nn * 24, s);
    for (int lvl = out->numLevels - 1; lvl >= 0; --lvl) {
      const int l0 = out->levelStart[lvl], cnt = out->levelStart[lvl + 1] - l0;
      if (cnt <= 0) continue;
      tree_boxes_level_kernel<<<(cnt + 127) / 128, 128, 0, s>>>(t, out->d_pos, l0, cnt, out->d_boxlo, out->d_boxhi);
      cudaChk(cudaPeekAtLastError());
      g_launches.fetch_add(1);
    }
    g_launches.fetch_add(1);
  }
  pool_free(flag, s); pool_free(rank, s); pool_free(leafAt, s); pool_free(scanTmp, s); pool_free(d_targets, s);
  pool_free(split, s); pool_free(slot, s); pool_free(blockSums, s); pool_free(meta, s); pool_free(keys, s);
}

void cb200_build_tree(const double *d_pos_xyz, const double *d_mass, const double *d_soft, int n, int maxBucket,
                      const double *rootlo, const double *roothi, cb200_tree *out, void *stream) {
  TreeInput in;
  in.pos = d_pos_xyz; in.mass = d_mass; in.soft = d_soft;
  in.posStride = 3; in.attrStride = 1;
  build_tree_impl(in, n, maxBucket, rootlo, roothi, out, nullptr, 0, nullptr, (cudaStream_t)stream);
}

void cb200_tree_free(cb200_tree *t, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  pool_free(t->d_pos, s); pool_free(t->d_mass, s); pool_free(t->d_soft, s); pool_free(t->d_packedParts, s);
  pool_free(t->d_order, s); pool_free(t->d_child0, s); pool_free(t->d_child1, s); pool_free(t->d_parent, s);
  pool_free(t->d_first, s); pool_free(t->d_last, s); pool_free(t->d_geolo, s); pool_free(t->d_geohi, s);
  pool_free(t->d_boxlo, s); pool_free(t->d_boxhi, s); pool_free(t->d_bucketNode, s);
  pool_free(t->d_bucketFirst, s); pool_free(t->d_bucketCount, s); pool_free(t->d_bucketStarts, s);
  pool_free(t->d_bucketSizes, s);
  memset(t, 0, sizeof *t);
}

/* ---- interaction lists on the device ---- */
void cb200_walk_device(int numNodes, int numBuckets, int numLevels, const int *h_levelStart,
                       const int *d_child0, const int *d_child1, const int *d_parent,
                       const int *d_firstPart, const int *d_lastPart, const int *d_bucketFirst,
                       const int *d_bucketCount, const int *d_bucketNode, const double *d_boxlo_xyz,
                       const double *d_boxhi_xyz, const double *d_moments_f64, double theta, int nReplicas,
                       double period, int bucketLo, int bucketHi, cb200_lists *out, void *stream) {
  cb200_walk_device_active(numNodes, numBuckets, numLevels, h_levelStart, d_child0, d_child1, d_parent, d_firstPart,
                           d_lastPart, d_bucketFirst, d_bucketCount, d_bucketNode, d_boxlo_xyz, d_boxhi_xyz,
                           d_moments_f64, theta, nReplicas, period, bucketLo, bucketHi, nullptr, out, stream);
}

struct MinOp {
  __host__ __device__ int operator()(int a, int b) const { return a < b ? a : b; }
};

/* what a multi-GPU force step adds to the walk (force_step.cuh sets it around its call, on its own thread): the
 * part of the tree it built and the reduction of the largest node softening over the ranks */
struct WalkExtras {
  const unsigned char *built = nullptr;
  int builtAlways = 0;
  int markEnd = 0; /* end of the level below the block level: unbuilt nodes in [builtAlways, markEnd) carry the mark */
  std::function<void(unsigned long long *d_softMaxBits, cudaStream_t)> reduceSoftMax;
  /* in: capacities (entries) of the three pools (cells, buckets, undecided), 0 = sized from the tree;
   * out: entries the walk reserved.  A step sizes the pools of the next one from these */
  mutable unsigned long long poolHint[3] = {0, 0, 0};
};
static thread_local const WalkExtras *tl_walkExtras = nullptr;
void cb200_walk_device_active(int numNodes, int numBuckets, int numLevels, const int *h_levelStart,
                              const int *d_child0, const int *d_child1, const int *d_parent,
                              const int *d_firstPart, const int *d_lastPart, const int *d_bucketFirst,
                              const int *d_bucketCount, const int *d_bucketNode, const double *d_boxlo_xyz,
                              const double *d_boxhi_xyz, const double *d_moments_f64, double theta, int nReplicas,
                              double period, int bucketLo, int bucketHi, const unsigned char *d_bucketActive,
                              cb200_lists *out, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  memset(out, 0, sizeof *out);
  out->numBuckets = numBuckets;
  if (numNodes <= 0 || numBuckets <= 0) return;
  const int sms = device_info().sms;
  WalkTree t;
  t.numNodes = numNodes; t.numBuckets = numBuckets;
  t.child0 = d_child0; t.child1 = d_child1; t.parent = d_parent; t.first = d_firstPart; t.last = d_lastPart;
  t.bucketFirst = d_bucketFirst; t.bucketCount = d_bucketCount; t.bucketNode = d_bucketNode;
  t.boxlo = d_boxlo_xyz; t.boxhi = d_boxhi_xyz; t.mom = d_moments_f64;
  WalkParams p;
  p.theta = theta; p.thetaMono = theta * theta * theta * theta; /* TreePiece.cpp:5036 */
  p.period = period; p.nReplicas = nReplicas;
  p.bucketLo = bucketLo < 0 ? 0 : bucketLo;
  p.bucketHi = bucketHi > numBuckets ? numBuckets : bucketHi;
  p.nextActive = nullptr;
  int *nextActive = nullptr, *activeIdx = nullptr;
  void *scanTmp = nullptr;
  if (d_bucketActive) { /* multistep: only buckets with rungs >= activeRung are walked */
    const int n1 = numBuckets + 1;
    activeIdx = (int *)pool_alloc((size_t)n1 * sizeof(int), s);
    nextActive = (int *)pool_alloc((size_t)n1 * sizeof(int), s);
    walk_active_index_kernel<<<(n1 + 255) / 256, 256, 0, s>>>(d_bucketActive, numBuckets, activeIdx);
    cudaChk(cudaPeekAtLastError());
    auto in = thrust::make_reverse_iterator(activeIdx + n1);
    auto outIt = thrust::make_reverse_iterator(nextActive + n1);
    size_t bytes = 0;
    cudaChk(cub::DeviceScan::InclusiveScan(nullptr, bytes, in, outIt, MinOp(), n1, s));
    scanTmp = pool_alloc(bytes, s);
    cudaChk(cub::DeviceScan::InclusiveScan(scanTmp, bytes, in, outIt, MinOp(), n1, s));
    g_launches.fetch_add(2);
    p.nextActive = nextActive;
  }

  /* per-node list slices come from three pools sized from the tree (host walk: ~110 cell,
   * ~45 undecided, ~12 bucket entries per node); the error flag reports an overflow */
  /* only nodes above the range's buckets are visited: a rank that walks 1/N of the buckets
   * needs ~1/N of the pools (plus the shared top of the tree) */
  const double frac = (double)(p.bucketHi > p.bucketLo ? p.bucketHi - p.bucketLo : 0) / (double)numBuckets;
  unsigned long long visited = (unsigned long long)((double)numNodes * (frac * 1.25 < 1.0 ? frac * 1.25 : 1.0)) +
                               65536ull + 4096ull * (unsigned long long)numLevels;
  if (visited > (unsigned long long)numNodes) visited = (unsigned long long)numNodes;
  WalkPools pools;
  /* + what the warps' last chunks of every level may leave unused (walk_level_kernel reserves in chunks) */
  const unsigned long long chunkSlack = (unsigned long long)sms * CB200_WALK_MINB * kWalkWarps * 4096ull * 8ull +
                                        (unsigned long long)numLevels * 4096ull * 64ull;
  pools.capC = visited * 256 + (1u << 16) + chunkSlack;
  pools.capU = visited * 128 + (1u << 16) + chunkSlack / 4;
  pools.capL = visited * 64 + (1u << 16) + chunkSlack / 2;
  if (tl_walkExtras && tl_walkExtras->poolHint[0]) { /* capacities the caller learned from earlier steps */
    const unsigned long long *hint = tl_walkExtras->poolHint;
    if (hint[0] < pools.capC) pools.capC = hint[0];
    if (hint[1] < pools.capL) pools.capL = hint[1];
    if (hint[2] < pools.capU) pools.capU = hint[2];
  }
  pools.clist = (WalkEntry *)pool_alloc(pools.capC * sizeof(WalkEntry), s);
  pools.lplist = (WalkEntry *)pool_alloc(pools.capL * sizeof(WalkEntry), s);
  pools.undlist = (WalkEntry *)pool_alloc(pools.capU * sizeof(WalkEntry), s);
  char *ctl = (char *)pool_alloc(256, s);
  cudaChk(cudaMemsetAsync(ctl, 0, 256, s));
  pools.used = (unsigned long long *)ctl;
  pools.error = (int *)(ctl + 64);
  pools.stats = (int *)(ctl + 192);
  NodeLists *lists = (NodeLists *)pool_alloc((size_t)numNodes * sizeof(NodeLists), s);
  WalkNodeRec *rec = (WalkNodeRec *)pool_alloc((size_t)numNodes * sizeof(WalkNodeRec), s);
  t.softMaxBits = (const unsigned long long *)(ctl + 128); /* ctl is zeroed above */
  const WalkExtras *extras = tl_walkExtras;
  WalkNodeRecF *recf = (WalkNodeRecF *)pool_alloc((size_t)numNodes * sizeof(WalkNodeRecF), s);
  walk_pack_nodes_kernel<<<(numNodes + 255) / 256, 256, 0, s>>>(t, p.theta, p.thetaMono, p.period, rec, recf,
                                                                (unsigned long long *)(ctl + 128),
                                                                extras ? extras->built : nullptr,
                                                                extras ? extras->builtAlways : 0,
                                                                extras ? extras->markEnd : 0);
  cudaChk(cudaPeekAtLastError());
  if (extras && extras->reduceSoftMax) extras->reduceSoftMax((unsigned long long *)(ctl + 128), s);
  t.recf = recf;
  t.ftol = (const WalkFloatTol *)(ctl + 224);
  walk_float_tol_kernel<<<1, 1, 0, s>>>(t, p.period, (WalkFloatTol *)(ctl + 224));
  cudaChk(cudaPeekAtLastError());
  g_launches.fetch_add(1);
  g_launches.fetch_add(1);
  t.rec = rec;
  /* CB200_WALK_MINB (5) CTAs of 4 warps per SM: what the kernel's 96 registers and its 44 KB of shared memory
   * (the heads of the per-node lists) both allow; the walk is latency-bound and wants the warps */
  const int walkCtas = sms * CB200_WALK_MINB; /* resident CTAs per SM the level kernel is compiled for */
  {
    static CtaCache attrSet;
    int dev = 0;
    cudaChk(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (attrSet.n[dev].load(std::memory_order_acquire) == 0) {
      cudaChk(cudaFuncSetAttribute(walk_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWalkSmemBytes));
      attrSet.n[dev].store(1, std::memory_order_release);
    }
  }
  static const int generalOnly = getenv("CB200_WALK_GENERAL") != nullptr; /* A/B switch: every node through walk_node_general */
  WalkEntry *scratch = (WalkEntry *)pool_alloc((size_t)walkCtas * kWalkWarps * 4 * kWalkCap * sizeof(WalkEntry), s);
  /* per level, the nodes above the range's buckets (all of them on one GPU) */
  int2 *levelRange = (int2 *)pool_alloc(66 * sizeof(int2), s);
  {
    WalkLevels lv;
    lv.n = numLevels < 65 ? numLevels : 65;
    for (int l = 0; l <= lv.n; ++l) lv.start[l] = h_levelStart[l];
    walk_level_ranges_kernel<<<1, 96, 0, s>>>(t, lv, p.bucketLo, p.bucketHi, levelRange);
    cudaChk(cudaPeekAtLastError());
    g_launches.fetch_add(1);
  }
  const double rangeFrac = frac * 1.25 + 1e-3 < 1.0 ? frac * 1.25 + 1e-3 : 1.0; /* grid bound only: the kernel loops */
  for (int lvl = 0; lvl < numLevels; ++lvl) {
    const int nAll = h_levelStart[lvl + 1] - h_levelStart[lvl];
    if (nAll <= 0) continue;
    const int n = (int)((double)nAll * rangeFrac) + 64 < nAll ? (int)((double)nAll * rangeFrac) + 64 : nAll;
    const int need = (n + kWalkWarps - 1) / kWalkWarps;
    const int grid = need < walkCtas ? need : walkCtas;
    /* pool chunk of a warp: ~128 cell entries per node it will see, between 256 and 4096 */
    const long long perWarp = ((long long)n + (long long)grid * kWalkWarps - 1) / ((long long)grid * kWalkWarps);
    const int chunk = (int)(perWarp * 128 < 256 ? 256 : (perWarp * 128 > 4096 ? 4096 : perWarp * 128));
    walk_level_kernel<<<grid, kWalkWarps * 32, kWalkSmemBytes, s>>>(t, p, levelRange + lvl, lists, pools, scratch, generalOnly,
                                                                  chunk);
    cudaChk(cudaPeekAtLastError());
    g_launches.fetch_add(1);
  }

  /* per-bucket sizes -> markers */
  const size_t nb1 = (size_t)numBuckets + 1;
  int *counts = (int *)pool_alloc(3 * nb1 * sizeof(int), s);
  cudaChk(cudaMemsetAsync(counts, 0, 3 * nb1 * sizeof(int), s));
  out->d_cellMarkers = (int *)pool_alloc(nb1 * sizeof(int), s);
  out->d_softMarkers = (int *)pool_alloc(nb1 * sizeof(int), s);
  out->d_partMarkers = (int *)pool_alloc(nb1 * sizeof(int), s);
  out->d_starts = (int *)pool_alloc(nb1 * sizeof(int), s);
  out->d_sizes = (int *)pool_alloc(nb1 * sizeof(int), s);
  int *flaggedBuckets = (int *)pool_alloc(nb1 * sizeof(int), s);
  emit_count_kernel<<<(numBuckets + 255) / 256, 256, 0, s>>>(t, p, lists, counts, counts + nb1, counts + 2 * nb1, out->d_starts,
                                                            out->d_sizes, flaggedBuckets, (int *)(ctl + 208));
  cudaChk(cudaPeekAtLastError());
  emit_count_flagged_kernel<<<sms * 8, kWalkWarps * 32, 0, s>>>(t, p, lists, pools, counts, counts + nb1, flaggedBuckets,
                                                               (const int *)(ctl + 208));
  cudaChk(cudaPeekAtLastError());
  g_launches.fetch_add(1);
  walk_totals_kernel<<<((int)nb1 + 255) / 256, 256, 0, s>>>(counts, (int)nb1, (unsigned long long *)(ctl + 160));
  size_t tmpBytes = 0;
  cudaChk(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, counts, out->d_cellMarkers, (int)nb1, s));
  void *tmp = pool_alloc(tmpBytes, s);
  cudaChk(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts, out->d_cellMarkers, (int)nb1, s));
  cudaChk(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts + nb1, out->d_softMarkers, (int)nb1, s));
  cudaChk(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, counts + 2 * nb1, out->d_partMarkers, (int)nb1, s));
  g_launches.fetch_add(4);
  int totals[4] = {0, 0, 0, 0};
  cudaChk(cudaMemcpyAsync(&totals[0], out->d_cellMarkers + numBuckets, sizeof(int), cudaMemcpyDeviceToHost, s));
  cudaChk(cudaMemcpyAsync(&totals[1], out->d_softMarkers + numBuckets, sizeof(int), cudaMemcpyDeviceToHost, s));
  cudaChk(cudaMemcpyAsync(&totals[2], out->d_partMarkers + numBuckets, sizeof(int), cudaMemcpyDeviceToHost, s));
  cudaChk(cudaMemcpyAsync(&totals[3], pools.error, sizeof(int), cudaMemcpyDeviceToHost, s));
  unsigned long long wide[3] = {0, 0, 0}, usedPool[3] = {0, 0, 0};
  cudaChk(cudaMemcpyAsync(wide, ctl + 160, sizeof wide, cudaMemcpyDeviceToHost, s));
  cudaChk(cudaMemcpyAsync(usedPool, ctl, sizeof usedPool, cudaMemcpyDeviceToHost, s));
  cudaChk(cudaStreamSynchronize(s)); /* the list sizes decide the allocations below */
  if (tl_walkExtras) for (int k = 0; k < 3; ++k) tl_walkExtras->poolHint[k] = usedPool[k];
#ifdef CB200_WALK_STATS
  { int st[4]; cudaChk(cudaMemcpy(st, ctl + 192, sizeof st, cudaMemcpyDeviceToHost));
    fprintf(stderr, "walk stats: fast routine gave up on clist %d, buckets %d, undecided %d, ring %d of %d nodes\n", st[0], st[1], st[2], st[3], numNodes); }
#endif
  out->nCell = (long long)wide[0]; out->nSoft = (long long)wide[1]; out->nPart = (long long)wide[2];
  out->error = totals[3];
  for (int k = 0; k < 3; ++k)
    if (wide[k] > 0x7fffffffull) out->error = 3; /* 32-bit markers: walk a narrower bucket range */
  if (!out->error) {
    out->d_cell = (ILCell *)pool_alloc((size_t)(totals[0] > 0 ? totals[0] : 1) * sizeof(ILCell), s);
    out->d_soft = (ILCell *)pool_alloc((size_t)(totals[1] > 0 ? totals[1] : 1) * sizeof(ILCell), s);
    out->d_part = (ILCell *)pool_alloc((size_t)(totals[2] > 0 ? totals[2] : 1) * sizeof(ILCell), s);
    /* the path table (walk_paths_kernel): one row of numLevels node indices per bucket of the range */
    static const bool oldEmit = getenv("CB200_EMIT_CLIMB") != nullptr; /* A/B switch: the parent-link climb */
    const int nRange = p.bucketHi > p.bucketLo ? p.bucketHi - p.bucketLo : 0;
    int *path = nullptr;
    if (!oldEmit && nRange > 0) {
      path = (int *)pool_alloc((size_t)nRange * numLevels * sizeof(int), s);
      walk_paths_kernel<<<(nRange + 127) / 128, 128, 0, s>>>(t, p.bucketLo, p.bucketHi, numLevels, path);
      cudaChk(cudaPeekAtLastError());
      g_launches.fetch_add(1);
    }
    if (nRange > 0) {
      emit_fill_kernel<<<(nRange + kWalkWarps - 1) / kWalkWarps, kWalkWarps * 32, 0, s>>>(
          t, p, lists, pools, out->d_cellMarkers, out->d_softMarkers, out->d_partMarkers, out->d_cell, out->d_soft, out->d_part,
          path, numLevels);
      cudaChk(cudaPeekAtLastError());
    }
    pool_free(path, s);
    if (out->nSoft > 0) { /* the sources of the softened-cell list: made only when some cell is softened */
      out->d_nodeParticles = pool_alloc((size_t)numNodes * sizeof(PackedPart), s);
      nodes_as_particles_kernel<<<(numNodes + 255) / 256, 256, 0, s>>>(d_moments_f64, (PackedPart *)out->d_nodeParticles,
                                                                      numNodes);
      cudaChk(cudaPeekAtLastError());
      g_launches.fetch_add(1);
    }
    g_launches.fetch_add(2);
  }
  pool_free(flaggedBuckets, s); pool_free(levelRange, s);
  pool_free(tmp, s); pool_free(counts, s); pool_free(scratch, s); pool_free(lists, s); pool_free(ctl, s);
  pool_free(rec, s); pool_free(recf, s); pool_free(scanTmp, s); pool_free(activeIdx, s); pool_free(nextActive, s);
  pool_free(pools.clist, s); pool_free(pools.lplist, s); pool_free(pools.undlist, s);
}

void cb200_active_sets_device(const unsigned char *d_rung, const int *d_order, int numParticles,
                              const int *d_bucketStarts, const int *d_bucketSizes, int numBuckets, int activeRung,
                              unsigned char *d_bucketActive, int *d_ewaldMarkers, int *h_counts, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  h_counts[0] = h_counts[1] = 0;
  if (numParticles <= 0 || numBuckets <= 0) return;
  unsigned char *flags = (unsigned char *)pool_alloc((size_t)numParticles, s);
  int *ctl = (int *)pool_alloc(256, s);
  cudaChk(cudaMemsetAsync(ctl, 0, 256, s));
  active_particle_flags_kernel<<<(numParticles + 255) / 256, 256, 0, s>>>(d_rung, d_order, numParticles, activeRung, flags);
  cudaChk(cudaPeekAtLastError());
  active_bucket_flags_kernel<<<(numBuckets + 255) / 256, 256, 0, s>>>(flags, d_bucketStarts, d_bucketSizes, numBuckets,
                                                                     d_bucketActive, ctl);
  cudaChk(cudaPeekAtLastError());
  thrust::counting_iterator<int> idx(0);
  size_t bytes = 0;
  cudaChk(cub::DeviceSelect::Flagged(nullptr, bytes, idx, flags, d_ewaldMarkers, ctl + 1, numParticles, s));
  void *tmp = pool_alloc(bytes, s);
  cudaChk(cub::DeviceSelect::Flagged(tmp, bytes, idx, flags, d_ewaldMarkers, ctl + 1, numParticles, s));
  g_launches.fetch_add(3);
  cudaChk(cudaMemcpyAsync(h_counts, ctl, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  cudaChk(cudaStreamSynchronize(s)); /* the caller sizes the Ewald launch by the count */
  pool_free(tmp, s); pool_free(flags, s); pool_free(ctl, s);
}

void cb200_lists_free(cb200_lists *l, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  pool_free(l->d_cell, s); pool_free(l->d_soft, s); pool_free(l->d_part, s);
  pool_free(l->d_cellMarkers, s); pool_free(l->d_softMarkers, s); pool_free(l->d_partMarkers, s);
  pool_free(l->d_starts, s); pool_free(l->d_sizes, s); pool_free(l->d_nodeParticles, s);
  memset(l, 0, sizeof *l);
}

/* ---- bucket partitioner: contiguous SFC ranges of equal cost ---- */
void cb200_partition_buckets(const double *cost, int numBuckets, int nRanks, int *cuts) {
  double total = 0.0;
  for (int i = 0; i < numBuckets; ++i) total += cost[i];
  cuts[0] = 0;
  double run = 0.0;
  int b = 0;
  for (int r = 1; r < nRanks; ++r) {
    const double target = total * r / nRanks;
    /* advance while adding the next bucket keeps us at or below the target,
     * or lands closer to it than stopping short would */
    while (b < numBuckets && (run + cost[b] <= target || (target - run) > (run + cost[b] - target))) {
      run += cost[b];
      ++b;
    }
    cuts[r] = b;
  }
  cuts[nRanks] = numBuckets;
}

} /* extern "C" */

#include "force_step.cuh"
