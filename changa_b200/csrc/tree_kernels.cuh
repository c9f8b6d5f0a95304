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
#include <cstdint>
#include "device_layout.cuh"

namespace cb200 {

constexpr int kTreeKeyBitsPerDim = 21;
constexpr int kTreeKeyBits = 63;
constexpr int kTreeMaxLevels = 64;

struct TreeBox { double lo[3], inv[3]; };

__global__ void tree_keys_kernel(const double *__restrict__ pos, int n, TreeBox box,
                                 unsigned long long *__restrict__ keys, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned q[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double f = __dmul_rn(__dsub_rn(pos[3 * (size_t)i + d], box.lo[d]), box.inv[d]);
    if (f < 0.0) f = 0.0;
    const double s = __dmul_rn(f, (double)(1u << kTreeKeyBitsPerDim));
    q[d] = s >= (double)(1u << kTreeKeyBitsPerDim) ? (1u << kTreeKeyBitsPerDim) - 1 : __double2uint_rz(s);
  }
  unsigned long long k = 0;
#pragma unroll
  for (int b = kTreeKeyBitsPerDim - 1; b >= 0; --b)
    k = (k << 3) | (unsigned long long)((((q[0] >> b) & 1) << 2) | (((q[1] >> b) & 1) << 1) | ((q[2] >> b) & 1));
  keys[i] = k;
  idx[i] = i;
}

__global__ void tree_gather_kernel(const double *__restrict__ pos, const double *__restrict__ mass,
                                   const double *__restrict__ soft, const int *__restrict__ order, int n,
                                   double *__restrict__ spos, double *__restrict__ smass,
                                   double *__restrict__ ssoft, PackedPart *__restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = order[i];
  const double x = pos[3 * (size_t)o], y = pos[3 * (size_t)o + 1], z = pos[3 * (size_t)o + 2];
  const double m = mass[o], h = soft[o];
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

/* nodes [lo, lo+cnt) of one level: where each splits and how many children it gets */
__global__ void tree_split_kernel(TreeArrays t, const unsigned long long *__restrict__ keys, int lo, int cnt,
                                  int level, int maxBucket, int *__restrict__ split, int *__restrict__ nkids) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > cnt) return;
  if (w == cnt) { nkids[w] = 0; return; } /* the scan's total lands here */
  const int i = lo + w;
  const int f = t.first[i], l = t.last[i];
  if (l - f < maxBucket || level >= kTreeKeyBits - 3) { split[w] = -1; nkids[w] = 0; return; }
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
  split[w] = s;
  nkids[w] = (s > f) + (s <= l);
}

__global__ void tree_emit_kernel(TreeArrays t, int lo, int cnt, int level, const int *__restrict__ split,
                                 const int *__restrict__ slot, int nextLo, int cap, int *__restrict__ error) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= cnt) return;
  const int s = split[w];
  if (s < 0) return;
  const int i = lo + w;
  const int f = t.first[i], l = t.last[i];
  const int dim = level % 3;
  double glo[3], ghi[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { glo[d] = t.geolo[3 * (size_t)i + d]; ghi[d] = t.geohi[3 * (size_t)i + d]; }
  const double mid = __dmul_rn(0.5, __dadd_rn(ghi[dim], glo[dim]));
  int c = nextLo + slot[w];
  if (s > f) {
    if (c >= cap) { *error = 1; return; }
    t.child0[i] = c;
    t.child0[c] = t.child1[c] = -1; t.parent[c] = i; t.first[c] = f; t.last[c] = s - 1;
#pragma unroll
    for (int d = 0; d < 3; ++d) { t.geolo[3 * (size_t)c + d] = glo[d]; t.geohi[3 * (size_t)c + d] = d == dim ? mid : ghi[d]; }
    ++c;
  }
  if (s <= l) {
    if (c >= cap) { *error = 1; return; }
    t.child1[i] = c;
    t.child0[c] = t.child1[c] = -1; t.parent[c] = i; t.first[c] = s; t.last[c] = l;
#pragma unroll
    for (int d = 0; d < 3; ++d) { t.geolo[3 * (size_t)c + d] = d == dim ? mid : glo[d]; t.geohi[3 * (size_t)c + d] = ghi[d]; }
  }
}

/* flag[p] = 1 where a bucket starts (p = its first particle); flag has n+1 entries */
__global__ void tree_leaf_flags_kernel(TreeArrays t, int numNodes, int *__restrict__ flag, int *__restrict__ leafAt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numNodes) return;
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
