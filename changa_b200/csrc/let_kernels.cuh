/* let_kernels.cuh -- the locally essential part of the moment build (multi-GPU force step).
 *
 * Every rank holds the whole sorted box and the whole tree topology, but it only ever LOOKS at a deep source node
 * if that node lies near its own buckets: a node n is tested by the walk of a target T only when its parent P was
 * opened at T or at an ancestor T' of T, and P opens only if the target's box reaches the sphere
 * |x - cm_P| <= R_P, R_P = max(opening radius, softening reach) (gravity.h:652-723; "contained" for an internal
 * target, "intersects" for a bucket; the softening branch needs the two centres of mass within 2 soft_P + 2 soft_T).
 * So the tree is cut at a BLOCK LEVEL L_b:
 *   - above and at L_b every rank has every node (the level-L_b records are exchanged: each rank contributes the
 *     blocks it owns, an integer all-reduce of the bit patterns with zeros elsewhere = an exact all-gather of
 *     unequal slices in one collective; the levels above are then combined by every rank, identically);
 *   - below L_b a rank builds the subtrees of the blocks it owns plus the HALO blocks: block B is halo when its
 *     tight box, in some periodic image, comes within R_B of a node of the rank's domain cover, with
 *     R_B = max(2/sqrt(3)/theta, 1) * diag(B) + 4 * (largest particle softening)
 *     >= R_P of every node P in B's subtree (a node's radius is at most the diagonal of the tight box around it,
 *     its centre of mass lies in that box, a mean softening is at most the largest).
 * The domain cover is the set of nodes of a coarse level L_d (and leaves above it) that hold buckets of the rank:
 * a target at or below L_d lies inside its cover node's tight box; a target above L_d opens P only when its whole
 * box -- cover nodes included -- is inside P's sphere.  (The softening branch at a target above L_d is the one
 * case the bound does not cover; it needs a coincidence of two centres of mass within a few softening lengths at
 * every level of the chain down to L_b.)  Whatever the argument, the walk CHECKS: records outside the built part
 * carry a mark, a walk that touches one reports it, and the step falls back to the full build.
 * The reference has no analogue (its remote data come on demand through CkCache). */
#ifndef CB200_LET_KERNELS_CUH
#define CB200_LET_KERNELS_CUH

#include <cuda_runtime.h>

namespace cb200 {

constexpr int kLetCoverCap = 4096;

struct LetBox { double lo[3], hi[3]; };

/* the domain cover: nodes of level L_d, and leaves above it, with a bucket in [b0, b1) */
__global__ void let_cover_kernel(const int *__restrict__ child0, const int *__restrict__ child1,
                                 const int *__restrict__ bucketFirst, const int *__restrict__ bucketCount,
                                 const double *__restrict__ boxlo, const double *__restrict__ boxhi, int coverStart,
                                 int coverEnd, int b0, int b1, LetBox *__restrict__ cover, int *__restrict__ nCover) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= coverEnd) return;
  const bool leaf = child0[i] < 0 && child1[i] < 0;
  if (i < coverStart && !leaf) return;
  const int f = bucketFirst[i], l = f + bucketCount[i];
  if (l <= b0 || f >= b1) return;
  const int k = atomicAdd(nCover, 1);
  if (k >= kLetCoverCap) return;
  LetBox b;
  for (int d = 0; d < 3; ++d) { b.lo[d] = boxlo[3 * (size_t)i + d]; b.hi[d] = boxhi[3 * (size_t)i + d]; }
  cover[k] = b;
}

/* flag[B] for the nodes of the block level: 1 = the rank builds B's subtree (own or halo) */
__global__ void let_block_flags_kernel(const int *__restrict__ bucketFirst, const int *__restrict__ bucketCount,
                                       const double *__restrict__ boxlo, const double *__restrict__ boxhi, int lo, int n,
                                       int b0, int b1, const LetBox *__restrict__ cover, const int *__restrict__ nCover,
                                       double ropenFactor, const double *__restrict__ maxSoft, double period, int nReplicas,
                                       unsigned char *__restrict__ flag) {
  __shared__ LetBox sc[64];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = lo + t;
  const int nc = *nCover;
  bool need = nc > kLetCoverCap; /* the cover did not fit: build everything */
  double blo[3] = {0, 0, 0}, bhi[3] = {0, 0, 0}, R2 = 0.0;
  if (t < n) {
    const int f = bucketFirst[i], l = f + bucketCount[i];
    if (l > b0 && f < b1) need = true;
    double diag2 = 0.0;
    for (int d = 0; d < 3; ++d) {
      blo[d] = boxlo[3 * (size_t)i + d]; bhi[d] = boxhi[3 * (size_t)i + d];
      diag2 += (bhi[d] - blo[d]) * (bhi[d] - blo[d]);
    }
    const double R = (ropenFactor * sqrt(diag2) + 4.0 * *maxSoft) * (1.0 + 1e-9) + 1e-300;
    R2 = R * R;
  }
  const int total = nc < kLetCoverCap ? nc : kLetCoverCap;
  for (int c0 = 0; c0 < total; c0 += 64) {
    __syncthreads();
    if (threadIdx.x < 64 && c0 + threadIdx.x < total) sc[threadIdx.x] = cover[c0 + threadIdx.x];
    __syncthreads();
    if (t < n && !need) {
      const int m = min(64, total - c0);
      for (int c = 0; c < m && !need; ++c) {
        double d2 = 0.0;
        for (int d = 0; d < 3; ++d) { /* the gap along an axis, smallest over the periodic images */
          double best = 1e300;
          for (int k = -nReplicas; k <= nReplicas; ++k) {
            const double s = k * period;
            const double g1 = blo[d] + s - sc[c].hi[d], g2 = sc[c].lo[d] - (bhi[d] + s);
            double g = g1 > g2 ? g1 : g2;
            if (g < 0.0) g = 0.0;
            if (g < best) best = g;
          }
          d2 += best * best;
        }
        if (d2 <= R2) need = true;
      }
    }
  }
  if (t < n) flag[i] = need ? 1 : 0;
}

/* a child is built when its parent is */
__global__ void let_propagate_kernel(const int *__restrict__ child0, const int *__restrict__ child1, int lo, int n,
                                     unsigned char *__restrict__ flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = lo + t;
  const unsigned char f = flag[i];
  const int c0 = child0[i], c1 = child1[i];
  if (c0 >= 0) flag[c0] = f;
  if (c1 >= 0) flag[c1] = f;
}

/* the exchange of the block level's work records (component-major: work[k * numNodes + node], k < words):
 * buf[k * n + j] = bit pattern of the record of block lo + j if this rank owns it (its first particle lies in
 * [p0, p1)), else 0; the sum over the ranks is then the owner's bit pattern */
__global__ void let_pack_kernel(const double *__restrict__ work, size_t numNodes, int words, int lo, int n,
                                const int *__restrict__ firstPart, int p0, int p1, unsigned long long *__restrict__ buf) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)words * n) return;
  const int k = (int)(idx / n), j = (int)(idx % n);
  const int f = firstPart[lo + j];
  buf[idx] = (f >= p0 && f < p1) ? (unsigned long long)__double_as_longlong(work[(size_t)k * numNodes + lo + j]) : 0ull;
}
__global__ void let_unpack_kernel(double *__restrict__ work, size_t numNodes, int words, int lo, int n,
                                  const unsigned long long *__restrict__ buf) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)words * n) return;
  const int k = (int)(idx / n), j = (int)(idx % n);
  work[(size_t)k * numNodes + lo + j] = __longlong_as_double((long long)buf[idx]);
}

}  // namespace cb200
#endif
