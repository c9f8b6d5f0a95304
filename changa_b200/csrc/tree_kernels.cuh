/* tree_kernels.cuh -- the tree TOPOLOGY on the device.
 *
 * Same tree as csrc/treewalk.cpp builds on the host (which is its bit-exact
 * specification and the statement of GenericTreeNode.h:473-598, Compute.cpp:2476-2477,
 * DataManager.cpp:797-828): 63-bit Morton keys of the particles in the root box,
 * particles sorted by (key, original index), a binary tree that splits on key bit
 * 62 - level (dimension level % 3, geometric box halved at its midpoint), a node is a
 * bucket when lastParticle - firstParticle < maxBucket, nodes numbered breadth first
 * (children 0 then 1, empty children skipped), buckets numbered in particle order.
 *
 *   tree_keys_kernel        key + original index per particle
 *   cub::DeviceRadixSort    stable: ties keep index order, as the host comparator
 *   tree_gather_kernel      sorted positions / masses / softenings (+ PackedPart)
 *   per level:  tree_split_kernel (binary search of the split, child count)
 *               cub::DeviceScan   (child slots in parent order)
 *               tree_emit_kernel  (children: particle range, geometric box, links)
 *   tree_leaf_flags_kernel + scan + tree_buckets_kernel   buckets in particle order,
 *                                   first bucket / bucket count of every node
 *   tree_boxes_level_kernel   tight bounding boxes, bottom-up
 *
 * All geometry in double with explicitly rounded operations (no FMA contraction),
 * so every array equals the host's bit for bit (tests/test_gpu_parity.py).
 */
#ifndef CB200_TREE_KERNELS_CUH
#define CB200_TREE_KERNELS_CUH

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdint>
#include "device_layout.cuh"

namespace cb200 {

constexpr int kTreeKeyBitsPerDim = 21;
constexpr int kTreeKeyBits = 63;
constexpr int kTreeMaxLevels = 64;

struct TreeBox { double lo[3], inv[3]; };

/* The unsorted input: positions pos[i * posStride + d], masses mass[i * attrStride], softenings
 * soft[i * attrStride] -- three separate arrays (strides 3, 1) or one array of {x, y, z, m, soft}
 * records as it arrives from the host or from the all-gather (strides 5, 5). */
struct TreeInput {
  const double *pos, *mass, *soft;
  int posStride, attrStride;
};

/* idxBase: index of row 0 in the caller's order (a rank keys its own slice of the box, force_step.cuh) */
__global__ void tree_keys_kernel(TreeInput in, int n, TreeBox box,
                                 unsigned long long *__restrict__ keys, int *__restrict__ idx, int idxBase = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned q[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double f = __dmul_rn(__dsub_rn(in.pos[(size_t)in.posStride * i + d], box.lo[d]), box.inv[d]);
    if (f < 0.0) f = 0.0;
    const double s = __dmul_rn(f, (double)(1u << kTreeKeyBitsPerDim));
    q[d] = s >= (double)(1u << kTreeKeyBitsPerDim) ? (1u << kTreeKeyBitsPerDim) - 1 : __double2uint_rz(s);
  }
  unsigned long long k = 0;
#pragma unroll
  for (int b = kTreeKeyBitsPerDim - 1; b >= 0; --b)
    k = (k << 3) | (unsigned long long)((((q[0] >> b) & 1) << 2) | (((q[1] >> b) & 1) << 1) | ((q[2] >> b) & 1));
  keys[i] = k;
  idx[i] = idxBase + i;
}

__global__ void tree_gather_kernel(TreeInput in, const int *__restrict__ order, int n,
                                   double *__restrict__ spos, double *__restrict__ smass,
                                   double *__restrict__ ssoft, PackedPart *__restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = order[i];
  const double *pp = in.pos + (size_t)in.posStride * o;
  const double x = pp[0], y = pp[1], z = pp[2];
  const double m = in.mass[(size_t)in.attrStride * o], h = in.soft[(size_t)in.attrStride * o];
  spos[3 * (size_t)i] = x; spos[3 * (size_t)i + 1] = y; spos[3 * (size_t)i + 2] = z;
  smass[i] = m; ssoft[i] = h;
  if (packed) {
    PackedPart q;
    q.x = (real)x; q.y = (real)y; q.z = (real)z; q.mass = (real)m;
    q.soft = (real)h; q.pad0 = q.pad1 = q.pad2 = 0;
    packed[i] = q;
  }
}

struct TreeArrays {
  int *child0, *child1, *parent, *first, *last;
  double *geolo, *geohi;
};

__global__ void tree_root_kernel(TreeArrays t, int n, TreeBox box, double hx, double hy, double hz) {
  t.child0[0] = t.child1[0] = -1; t.parent[0] = -1; t.first[0] = 0; t.last[0] = n - 1;
  t.geolo[0] = box.lo[0]; t.geolo[1] = box.lo[1]; t.geolo[2] = box.lo[2];
  t.geohi[0] = hx; t.geohi[1] = hy; t.geohi[2] = hz;
}

/* where node i (particles [f, l], depth `level`) splits and how many children it gets; -1: a bucket */
__device__ __forceinline__ int tree_split_of(const unsigned long long *__restrict__ keys, int f, int l, int level,
                                             int maxBucket, int &nkids) {
  nkids = 0;
  if (l - f < maxBucket || level >= kTreeKeyBits - 3) return -1;
  const int bit = kTreeKeyBits - 1 - level;
  const unsigned long long mask = 1ull << bit;
  const unsigned long long kf = keys[f], kl = keys[l];
  int s;
  if ((kf & mask) == (kl & mask)) {
    s = (kf & mask) ? f : l + 1;
  } else { /* first key of [f, l] with the split bit set (GenericTreeNode.h:573-578) */
    const unsigned long long probe = (kl & (~0ull << bit)) | mask;
    int a = f, b = l + 1;
    while (a < b) {
      const int mid = a + ((b - a) >> 1);
      if (keys[mid] < probe) a = mid + 1; else b = mid;
    }
    s = a;
  }
  nkids = (s > f) + (s <= l);
  return s;
}

/* children of node i (split at s) at slots c, c+1: particle range, geometric box halved along
 * level % 3, links */
__device__ __forceinline__ void tree_emit_children(TreeArrays t, int i, int s, int level, int c) {
  const int f = t.first[i], l = t.last[i];
  const int dim = level % 3;
  double glo[3], ghi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { glo[d] = t.geolo[3 * (size_t)i + d]; ghi[d] = t.geohi[3 * (size_t)i + d]; }
  const double mid = __dmul_rn(0.5, __dadd_rn(ghi[dim], glo[dim]));
  if (s > f) {
    t.child0[i] = c;
    t.child0[c] = t.child1[c] = -1; t.parent[c] = i; t.first[c] = f; t.last[c] = s - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) { t.geolo[3 * (size_t)c + d] = glo[d]; t.geohi[3 * (size_t)c + d] = d == dim ? mid : ghi[d]; }
    ++c;
  }
  if (s <= l) {
    t.child1[i] = c;
    t.child0[c] = t.child1[c] = -1; t.parent[c] = i; t.first[c] = s; t.last[c] = l;
#pragma unroll
    for (int d = 0; d < 3; ++d) { t.geolo[3 * (size_t)c + d] = d == dim ? mid : glo[d]; t.geohi[3 * (size_t)c + d] = ghi[d]; }
  }
}

/* what the host needs to know about the finished tree: one 4-byte-aligned record, copied back once */
constexpr int kTreeMaxCuts = 144; /* 17 ranks x 8 output slabs + 1, rounded up */
struct TreeMeta {
  int numLevels, numNodes, numBuckets, error;
  int levelStart[kTreeMaxLevels + 2];
  int cuts[2 * kTreeMaxCuts]; /* bucket and particle index of the rank (and output-slab) boundaries (tree_cuts_kernel) */
};

/* The whole level loop in ONE cooperative launch (round 1 launched split / scan / emit per level and
 * read the size of the next level back over PCIe before launching it: one stream synchronisation per
 * level).  Per level every CTA takes a contiguous slice of the level's nodes: phase 1 finds the splits
 * and scans the child counts inside the CTA; after a grid-wide barrier every CTA sums the CTA totals
 * in front of it (a few hundred values) and places its children -- parents in order, child 0 before
 * child 1, so the numbering is the breadth-first nodeArrayIndex (DataManager.cpp:797-828) -- and a
 * second barrier publishes the new level.  `split` / `slot`: scratch of n + 1 ints. */
constexpr int kTreeLevelThreads = 256;
__global__ void __launch_bounds__(kTreeLevelThreads)
tree_levels_kernel(TreeArrays t, const unsigned long long *__restrict__ keys, int maxBucket, int cap,
                   int *__restrict__ split, int *__restrict__ slot, int *__restrict__ blockSums, TreeMeta *meta) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_warp[kTreeLevelThreads / 32];
  __shared__ int s_pair[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int lo = 0, hi = 1, level = 0;
  bool overflow = false;
  while (lo < hi && level < kTreeMaxLevels) {
    const int cnt = hi - lo;
    const int per = (cnt + (int)gridDim.x - 1) / (int)gridDim.x;
    const int b0 = min(cnt, (int)blockIdx.x * per), b1 = min(cnt, b0 + per);
    int carry = 0;
    for (int base = b0; base < b1; base += kTreeLevelThreads) {
      const int w = base + threadIdx.x;
      int kids = 0;
      if (w < b1) split[w] = tree_split_of(keys, t.first[lo + w], t.last[lo + w], level, maxBucket, kids);
      int incl = kids;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int k = 0; k < kTreeLevelThreads / 32; ++k) {
        const int v = s_warp[k];
        if (k < warp) before += v;
        total += v;
      }
      if (w < b1) slot[w] = carry + before + incl - kids;
      carry += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) blockSums[blockIdx.x] = carry;
    grid.sync();
    int off = 0, tot = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kTreeLevelThreads) {
      const int v = blockSums[i];
      tot += v;
      if (i < (int)blockIdx.x) off += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if (threadIdx.x < 2) s_pair[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) { atomicAdd(&s_pair[0], off); atomicAdd(&s_pair[1], tot); }
    __syncthreads();
    const int offset = s_pair[0], total = s_pair[1];
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) meta->levelStart[level + 1] = hi;
    if (hi + total > cap) { overflow = true; break; } /* the same decision in every CTA */
    for (int w = b0 + threadIdx.x; w < b1; w += kTreeLevelThreads) {
      const int s = split[w];
      if (s >= 0) tree_emit_children(t, lo + w, s, level, hi + offset + slot[w]);
    }
    grid.sync();
    lo = hi;
    hi += total;
    ++level;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    meta->levelStart[0] = 0;
    meta->numLevels = level;
    meta->numNodes = hi;
    meta->error = overflow ? 1 : 0;
  }
}

/* flag[p] = 1 where a bucket starts (p = its first particle); flag has n+1 entries */
__global__ void tree_leaf_flags_kernel(TreeArrays t, const TreeMeta *__restrict__ meta, int *__restrict__ flag,
                                       int *__restrict__ leafAt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= meta->numNodes) return;
  if (t.child0[i] < 0 && t.child1[i] < 0) { flag[t.first[i]] = 1; leafAt[t.first[i]] = i; }
}

/* rank = exclusive scan of flag: bucket index of the bucket starting at p; rank[n] = number of buckets */
__global__ void tree_buckets_kernel(TreeArrays t, int numNodes, int n, const int *__restrict__ flag,
                                    const int *__restrict__ rank, const int *__restrict__ leafAt,
                                    int *__restrict__ bucketNode, int *__restrict__ bucketFirst,
                                    int *__restrict__ bucketCount, int *__restrict__ bucketStarts,
                                    int *__restrict__ bucketSizes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) {
    const int node = leafAt[i], b = rank[i];
    bucketNode[b] = node;
    bucketStarts[b] = t.first[node];
    bucketSizes[b] = t.last[node] - t.first[node] + 1;
  }
  if (i < numNodes) {
    bucketFirst[i] = rank[t.first[i]];
    bucketCount[i] = rank[t.last[i] + 1] - rank[t.first[i]];
  }
}

__global__ void tree_boxes_level_kernel(TreeArrays t, const double *__restrict__ spos, int lo, int cnt,
                                        double *__restrict__ boxlo, double *__restrict__ boxhi) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= cnt) return;
  const int i = lo + w;
  double bl[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, bh[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
  const int c0 = t.child0[i], c1 = t.child1[i];
  if (c0 < 0 && c1 < 0) {
    for (int p = t.first[i]; p <= t.last[i]; ++p)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double v = spos[3 * (size_t)p + d];
        bl[d] = fmin(bl[d], v); bh[d] = fmax(bh[d], v);
      }
  } else {
    const int ch[2] = {c0, c1};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (ch[k] < 0) continue;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        bl[d] = fmin(bl[d], boxlo[3 * (size_t)ch[k] + d]);
        bh[d] = fmax(bh[d], boxhi[3 * (size_t)ch[k] + d]);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) { boxlo[3 * (size_t)i + d] = bl[d]; boxhi[3 * (size_t)i + d] = bh[d]; }
}

}  // namespace cb200
#endif
