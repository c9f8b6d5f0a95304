/* walk_kernels.cuh -- interaction-list generation on the device.
 *
 * Emits, per bucket, exactly the ILCell / ILPart / softened-cell lists that
 * ChaNGa's host walk produces (LocalTargetWalk::dft, TreeWalk.cpp:308-397;
 * ListCompute::doWork, Compute.cpp:690-884; stateReady, Compute.cpp:1608-1863)
 * -- same entries, same order, same offsetID bits -- so the 8-byte-per-entry
 * lists never cross PCIe.  The specification is csrc/treewalk.cpp (host, bit-
 * exact against the sequential oracle); the device version is the same walk
 * organised for a GPU:
 *
 *   walk_level_kernel   level-synchronous over the LOCAL tree: one warp owns
 *                       one local node, starts from its parent's undecided
 *                       list (the 27 root replicas for the root), and drains a
 *                       FIFO checklist 32 entries at a time.  Processing an
 *                       entry depends only on (source node, local node), so a
 *                       batch evaluated in parallel and appended in lane order
 *                       (ballot + popc compaction) reproduces the sequential
 *                       order exactly.  Outputs: clist, lplist, undlist of the
 *                       node, copied into exactly-sized slices of a pool.
 *   emit_count/fill     one warp per bucket: concatenates the per-node lists
 *                       along the path root -> bucket, level by level, cells
 *                       (minus the ones openSoftening sends to the softened
 *                       list) and then particle buckets expanded per particle
 *                       (GenericList<ILPart>::serialize, Compute.cpp:1174-1187).
 *
 * Every opening decision is the double-precision one of the host walk (gravity.h:251-260,
 * 652-723): walk_node_fast evaluates a test in single precision first and accepts the result
 * only outside a rigorous error band (WalkNodeRecF); inside it the test is repeated in double.
 */
#ifndef CB200_WALK_KERNELS_CUH
#define CB200_WALK_KERNELS_CUH

#include <cuda_runtime.h>
#include "device_layout.cuh"

namespace cb200 {

constexpr int kWalkCap = 2048;          /* entries per node list / checklist (host walk: max ~640) */
constexpr int kWalkWarps = 4;           /* warps per CTA */
/* The head of every per-node list lives in shared memory; what does not fit spills to the warp's
 * slice of the global scratch at the same index.  Sizes cover the typical node (host walk: ~110
 * accepted cells, ~45 undecided, ~12 buckets, ~300 appended checklist entries at the leaves). */
constexpr int kWalkRingS = 512, kWalkClS = 256, kWalkUnS = 128, kWalkLpS = 64;
/* staged source records of a batch: the fast routine keeps two buffers of 32 float records at a 48-byte pitch, the
 * general routine one buffer of 32 double records at an 80-byte pitch */
constexpr int kWalkRowBytes = 2 * 32 * 48;
static_assert(kWalkRowBytes >= 32 * 80, "the general routine's single buffer fits");
constexpr int kWalkLpHead = 96;                                                  /* bucket-list head of the fast routine */
constexpr int kWalkWarpSmem = kWalkRowBytes + 8 * (kWalkRingS + kWalkClS + kWalkUnS + kWalkLpHead);
constexpr int kWalkSmemBytes = kWalkWarps * kWalkWarpSmem;
constexpr int kWalkOffsetMask = 0x1ff << 22;
constexpr int kWalkBucketMask = (1 << 22) - 1;

struct WalkEntry { int node; int offsetID; };

/* what one opening test reads of a SOURCE node, gathered into one 64-byte record (two
 * sectors) instead of five scattered arrays: the walk is a stream of dependent gathers */
struct __align__(16) WalkNodeRec {
  double cx, cy, cz;
  double ropen;     /* max(2/sqrt(3) * radius / theta, radius): the opening radius of gravity.h:665-668 */
  double soft;
  int child0, child1, first, last;
  double ropenMono; /* 2/sqrt(3) * radius / theta^4 (gravity.h:712) */
};
static_assert(sizeof(WalkNodeRec) == 64, "WalkNodeRec");

/* The same node for the FLOAT FILTER of walk_node_fast: an opening test is first evaluated in single precision with
 * a rigorous error band; only a test that lands inside the band (or near the softening reach) is repeated in double
 * from the 64-byte record, so every decision is the double one.  With u = 2^-24 and Mc = (largest coordinate of the
 * root's tight box) + 3 * period:
 *   - the shifted centre c, the box edges and their differences carry at most E2 = 6.2 u Mc each (conversion of cm,
 *     shift, edge; two rounded float operations), so |delta_f - delta| <= E2 per axis;
 *   - q_f = fl(sum delta_f^2) differs from the double dsq by at most 3.5 E2 sqrt(dsq) + 3 E2^2 + 3.2 u dsq, and
 *     sqrt(dsq) <= (dsq / R + R) / 2 for the node's opening radius R, so
 *       |q_f - R2_f| > tau (q_f + R2_f) + absTol,   tau = 1.8 E2 / R + 3.3 u (+ the cast of R^2),  absTol = 3 E2^2
 *     decides dsq <= R^2 either way (tau and absTol below carry a further factor 4 on E2).
 * tau travels with the node (it depends on R), absTol and the softening gate with the walk. */
struct __align__(16) WalkNodeRecF {
  float cx, cy, cz, ropen2;
  float tau;
  int child0, child1;
  int npart; /* last - first + 1; -1: not built by this rank (let_kernels.cuh) */
};
static_assert(sizeof(WalkNodeRecF) == 32, "WalkNodeRecF");
struct WalkFloatTol {
  float absTol;   /* 3 E2^2, rounded up */
  float softGate; /* q_f above it: the box is out of reach of every softening test and of the maybe-softened flag */
  float e2;       /* E2 (with the safety factor), for the gate */
};

/* moments arrive as 27 doubles per node in CudaMultipoleMoments order */
struct WalkTree {
  int numNodes, numBuckets;
  const int *child0, *child1, *first, *last, *bucketFirst, *bucketCount, *parent;
  const double *boxlo, *boxhi, *mom;
  const int *bucketNode;
  const WalkNodeRec *rec;
  const WalkNodeRecF *recf;
  const WalkFloatTol *ftol; /* filled by walk_float_tol_kernel once the largest softening is known */
  const unsigned long long *softMaxBits; /* max over nodes of `soft`, as the bits of a non-negative double */
};

constexpr int kWalkMaybeSoft = 1 << 31; /* clist entries only: some bucket below the owner MAY see this cell softened */

/* built (or NULL = every node): nodes from index builtAlways on carry a record only where built[] is set (the
 * locally essential build of a multi-GPU step, let_kernels.cuh); the others get a harmless record with the mark
 * first = -1, which the walk reports if it ever meets one */
constexpr int kWalkNotBuilt = 4; /* cb200_lists.error: the walk touched a node outside the built part of the tree */
/* scale of the float filter's error band: (largest |coordinate| of the root's tight box) + 3 |period| */
__device__ __forceinline__ double walk_coord_scale(const WalkTree &t, double period) {
  double m = 0.0;
  for (int d = 0; d < 3; ++d) { m = fmax(m, fabs(t.boxlo[d])); m = fmax(m, fabs(t.boxhi[d])); }
  return m + 3.0 * fabs(period);
}
constexpr double kWalkU = 5.9604644775390625e-08;           /* 2^-24 */
constexpr double kWalkE2Factor = 4.0 * 6.2 * kWalkU;         /* E2 = this x Mc (safety factor 4) */
__global__ void walk_float_tol_kernel(WalkTree t, double period, WalkFloatTol *__restrict__ out) {
  const double e2 = kWalkE2Factor * walk_coord_scale(t, period);
  const double B = 2.0 * (2.0 * __longlong_as_double((long long)*t.softMaxBits)) * 1.0001; /* (2 soft + rmMax) <= 2 rmMax */
  WalkFloatTol w;
  w.absTol = (float)(3.0 * e2 * e2 * 1.001) + 1e-37f;
  w.softGate = (float)((B * B * (1.0 + 4.0 * kWalkU) + 3.5 * e2 * B + 3.0 * e2 * e2) * 1.001) + 1e-37f;
  w.e2 = (float)(e2 * 1.001);
  *out = w;
}
__global__ void walk_pack_nodes_kernel(WalkTree t, double theta, double thetaMono, double period, WalkNodeRec *__restrict__ out,
                                       WalkNodeRecF *__restrict__ outf, unsigned long long *softMaxBits,
                                       const unsigned char *__restrict__ built, int builtAlways, int markEnd) {
  /* markEnd: nodes in [builtAlways, markEnd) -- the level right below the block level -- get the "not built"
   * mark when they are not built; deeper unbuilt nodes get nothing: a walk can only reach them through one of
   * those marked records, whose child links are empty */
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool have = i < t.numNodes && (!built || i < builtAlways || built[i]);
  double soft = 0.0;
  if (have) soft = fmax(t.mom[(size_t)i * 27 + 1], 0.0);
  /* non-negative doubles order like their bit patterns */
  unsigned long long b = (unsigned long long)__double_as_longlong(soft);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long v = __shfl_xor_sync(0xffffffffu, b, o);
    b = v > b ? v : b;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(softMaxBits, b);
  if (i >= t.numNodes) return;
  if (!have) {
    if (i >= markEnd) return;
    WalkNodeRec r;
    r.cx = r.cy = r.cz = r.ropen = r.soft = r.ropenMono = 0.0;
    r.child0 = r.child1 = -1; r.first = -1; r.last = -2;
    out[i] = r;
    WalkNodeRecF f;
    f.cx = f.cy = f.cz = f.ropen2 = f.tau = 0.0f;
    f.child0 = f.child1 = -1; f.npart = -1;
    outf[i] = f;
    return;
  }
  const double *m = t.mom + (size_t)i * 27;
  WalkNodeRec r;
  /* the two radii of openCriterionNode depend on the source node only: computed once here with the
   * host walk's operations (one product, one quotient, each rounded) instead of per opening test */
  const double geom = 2.0 / sqrt(3.0);
  const double gr = __dmul_rn(geom, m[0]);
  double ropen = __ddiv_rn(gr, theta);
  if (ropen < m[0]) ropen = m[0];
  r.ropen = ropen; r.ropenMono = __ddiv_rn(gr, thetaMono);
  r.soft = m[1]; r.cx = m[3]; r.cy = m[4]; r.cz = m[5];
  r.child0 = t.child0[i]; r.child1 = t.child1[i]; r.first = t.first[i]; r.last = t.last[i];
  out[i] = r;
  WalkNodeRecF f;
  f.cx = (float)r.cx; f.cy = (float)r.cy; f.cz = (float)r.cz;
  f.ropen2 = (float)__dmul_rn(ropen, ropen);
  const double e2 = kWalkE2Factor * walk_coord_scale(t, period);
  /* R == 0 (a point): the band is everything, the test goes to double */
  const double tau = ropen > 0.0 ? (1.8 * e2 / ropen + 4.5 * kWalkU) * 1.001 : 3.0e38;
  f.tau = tau < 3.0e38 ? (float)tau : 3.0e38f;
  f.child0 = r.child0; f.child1 = r.child1; f.npart = r.last - r.first + 1;
  outf[i] = f;
}

struct WalkParams {
  double theta, thetaMono, period;
  int nReplicas, bucketLo, bucketHi;
  /* multistep (bucketList[b]->rungs >= activeRung, Compute.cpp:1278,1574): nextActive[b] = smallest
   * active bucket >= b (numBuckets when there is none), numBuckets + 1 entries; NULL = every bucket */
  const int *nextActive;
};

/* per-node results: slices of the three pools */
struct NodeLists {
  int cOff, cLen, lOff, lLen, uOff, uLen;
  int visited;
  /* totals along the visited part of the path root -> this node: accepted cells, expanded
   * particle entries, cells flagged maybe-softened.  A bucket's list sizes are its node's
   * totals (minus the softened cells, which only flagged entries can be) */
  int pathCells, pathParts, pathFlagged;
  int ownParts; /* expanded particle entries of this node's own lplist */
  int parent;   /* copy of the tree's parent link: emit climbs bucket -> root with ONE dependent load per level */
};
static_assert(sizeof(NodeLists) == 48, "NodeLists");

struct WalkPools {
  WalkEntry *clist, *lplist, *undlist;
  unsigned long long capC, capL, capU;  /* entries per pool */
  unsigned long long *used;             /* [3] bump counters */
  int *error;                           /* != 0: a capacity was exceeded */
  int *stats;                           /* CB200_WALK_STATS builds: nodes the fast routine gave up on, by list */
};

/* Space::intersect(box, sphere) as the host walk states it (treewalk.cpp box_sphere; restated
 * in-tree at CUDAMoments.cu:137-159), without its early exits: the distance only grows, so
 * "rsq < dsq at some axis" and "dsq > rsq at the end" are the same answer; a zero term adds
 * exactly 0.  Products and sums are rounded separately, like the host build (-ffp-contract=off). */
/* max without the NaN bookkeeping of fmax (a dozen SASS instructions per call in double): the
 * operands are differences of finite coordinates */
__device__ __forceinline__ double walk_max(double a, double b) { return a > b ? a : b; }
/* squared distance from the point c to the box: what Space::intersect compares with r^2 */
__device__ __forceinline__ double walk_box_dist2(const double *lo, const double *hi, const double *c) {
  double dsq = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double delta = walk_max(walk_max(__dsub_rn(lo[d], c[d]), __dsub_rn(c[d], hi[d])), 0.0);
    dsq = __dadd_rn(dsq, __dmul_rn(delta, delta));
  }
  return dsq;
}
__device__ __forceinline__ bool walk_box_sphere(const double *lo, const double *hi, const double *c, double r) {
  return walk_box_dist2(lo, hi, c) <= __dmul_rn(r, r);
}
__device__ __forceinline__ bool walk_box_inside_sphere(const double *lo, const double *hi, const double *c, double r) {
  double s = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double w = walk_max(fabs(__dsub_rn(lo[d], c[d])), fabs(__dsub_rn(hi[d], c[d])));
    s = __dadd_rn(s, __dmul_rn(w, w));
  }
  return s <= __dmul_rn(r, r);
}
__device__ __forceinline__ void walk_shifted_cm(const WalkNodeRec &m, int offsetID, double period, double *c) {
  c[0] = __dadd_rn(m.cx, __dmul_rn((double)(((offsetID >> 22) & 7) - 3), period));
  c[1] = __dadd_rn(m.cy, __dmul_rn((double)(((offsetID >> 25) & 7) - 3), period));
  c[2] = __dadd_rn(m.cz, __dmul_rn((double)(((offsetID >> 28) & 7) - 3), period));
}
/* gravity.h:251-260; mm: the local node */
__device__ __forceinline__ bool walk_open_softening(const WalkNodeRec &m, const double *c, const WalkNodeRec &mm,
                                                    const double *mylo, const double *myhi) {
  const double rs = 2.0 * m.soft, rm = 2.0 * mm.soft;
  const double dx = __dsub_rn(mm.cx, c[0]), dy = __dsub_rn(mm.cy, c[1]), dz = __dsub_rn(mm.cz, c[2]);
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  const double rr = __dadd_rn(rs, rm);
  if (d2 <= __dmul_rn(rr, rr)) return true;
  return walk_box_sphere(mylo, myhi, c, rs);
}
/* gravity.h:652-723: 1 open, -1 undecided, 0 accept.  Every sphere of the test is centred on the
 * same shifted centre of mass c, so the box distance dsq (walk_box_dist2 of the local node's box)
 * is computed once by the caller and compared with each radius */
__device__ __forceinline__ int walk_open_criterion(const WalkNodeRec &m, const double *c, double dsq,
                                                   const WalkNodeRec &mine, const double *lo, const double *hi,
                                                   bool myIsBucket) {
  if (m.last - m.first + 1 <= 6) return 1;
  if (dsq <= __dmul_rn(m.ropen, m.ropen)) {
    if (myIsBucket) return 1;
    return walk_box_inside_sphere(lo, hi, c, m.ropen) ? 1 : -1;
  }
  /* openSoftening (gravity.h:251-260) */
  const double rs = 2.0 * m.soft, rm = 2.0 * mine.soft;
  const double dx = __dsub_rn(mine.cx, c[0]), dy = __dsub_rn(mine.cy, c[1]), dz = __dsub_rn(mine.cz, c[2]);
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  const double rr = __dadd_rn(rs, rm);
  if (!(d2 <= __dmul_rn(rr, rr)) && !(dsq <= __dmul_rn(rs, rs))) return 0;
  return dsq <= __dmul_rn(m.ropenMono, m.ropenMono) ? 1 : 0;
}

/* active buckets of a node under the [bucketLo, bucketHi) restriction and the active mask;
 * firstActive is the walk's target bucket (treewalk.cpp firstActive[]) */
__device__ __forceinline__ bool walk_node_active(const WalkTree &t, const WalkParams &p, int node, int &firstActive) {
  const int b0 = t.bucketFirst[node], b1 = b0 + t.bucketCount[node];
  int lo = b0 > p.bucketLo ? b0 : p.bucketLo;
  const int hi = b1 < p.bucketHi ? b1 : p.bucketHi;
  if (p.nextActive && lo < hi) lo = p.nextActive[lo];
  firstActive = lo;
  return lo < hi;
}
__device__ __forceinline__ bool walk_bucket_active(const WalkParams &p, int b) {
  return b >= p.bucketLo && b < p.bucketHi && (!p.nextActive || p.nextActive[b] == b);
}

/* nextActive from the per-bucket flags: own index where active, numBuckets elsewhere (and at the
 * end); a suffix minimum (cub, reversed) finishes it */
__global__ void walk_active_index_kernel(const unsigned char *__restrict__ active, int nb, int *__restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= nb) out[b] = (b < nb && active[b]) ? b : nb;
}

/* ordered append of the lanes whose flag is set */
__device__ __forceinline__ void walk_append(WalkEntry *dst, int &count, bool flag, WalkEntry e, int lane, int cap,
                                            int *error) {
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  const int pos = count + __popc(ballot & ((1u << lane) - 1));
  if (flag) {
    if (pos < cap) dst[pos] = e;
    else *error = 1;
  }
  count += __popc(ballot);
}

/* ---- the per-node walk ---------------------------------------------------------------------------
 * walk_level_kernel hands one local node to one warp.  The node's checklist is drained by one of two
 * routines that leave the node's three lists in the same place (heads in shared memory, what does not
 * fit in the warp's slice of the global scratch):
 *   walk_node_fast<LEAF>   the common case: everything the node touches fits the shared-memory heads
 *                          (checklist ring 512, clist 256, undecided 128, buckets 128; a head that fills up
 *                          is moved to the scratch as a whole), so no append has to choose between two
 *                          homes; specialised for bucket nodes (no undecided list, no contained-in-sphere
 *                          test).  Returns false when the live checklist would outgrow the ring;
 *   walk_node_general      the first version of this kernel's loop, kept out of line for those nodes. */

/* squared distance from the point c to the box [lo, hi] -- the value walk_box_dist2 computes, bit for
 * bit: at most one of lo - c and c - hi is positive, so max(lo - c, c - hi, 0) is the larger of the two
 * with negative values replaced by +0, which is done on the sign bit instead of a second
 * double-precision max (DSETP + selects + NaN bookkeeping) */
__device__ __forceinline__ double walk_pos_part(double d) {
  const int hi = __double2hiint(d), lo = __double2loint(d);
  const int keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, lo & keep);
}
__device__ __forceinline__ double walk_box_dist2_lean(const double (&lo)[3], const double (&hi)[3], double cx, double cy, double cz) {
  const double c[3] = {cx, cy, cz};
  double dsq = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double a = __dsub_rn(lo[d], c[d]), b = __dsub_rn(c[d], hi[d]);
    const double delta = walk_pos_part(a > b ? a : b);
    dsq = __dadd_rn(dsq, __dmul_rn(delta, delta));
  }
  return dsq;
}

constexpr int kWalkLpFast = kWalkLpHead; /* the general routine uses kWalkLpS of it */
static_assert(kWalkLpFast >= kWalkLpS, "one shared-memory area serves both routines");
constexpr int kWalkRowF = 3; /* uint4 per staged float record: 32 bytes + 16 of padding (48-byte pitch: LDS.128 conflict-free) */

/* the local node of a warp, in double, for the (rare) tests that are repeated in double: kept in shared memory so
 * that the hot loop holds only the float copies.  wd[0..2] box lo, [3..5] box hi, [6..8] centre of mass, [9] soft */
constexpr int kWalkMineDoubles = 10;

/* one opening test in double, exactly as walk_node_general evaluates it (openCriterionNode, gravity.h:652-723;
 * openSoftening :251-260; the maybe-softened flag of the emit step): 1 open, -1 undecided, 0 accept */
template <bool LEAF>
__device__ __noinline__ int walk_test_double(const WalkNodeRec *__restrict__ rec, int node, int offsetID, const double *wd,
                                             const double *shiftTab, double rmMax, int *flaggedOut) {
  const WalkNodeRec src = rec[node];
  const double lo[3] = {wd[0], wd[1], wd[2]}, hi[3] = {wd[3], wd[4], wd[5]};
  const double cx = __dadd_rn(src.cx, shiftTab[(offsetID >> 22) & 7]);
  const double cy = __dadd_rn(src.cy, shiftTab[(offsetID >> 25) & 7]);
  const double cz = __dadd_rn(src.cz, shiftTab[(offsetID >> 28) & 7]);
  const double dsq = walk_box_dist2_lean(lo, hi, cx, cy, cz);
  int open = 0;
  *flaggedOut = 0;
  if (src.last - src.first + 1 <= 6) return 1;
  if (dsq <= __dmul_rn(src.ropen, src.ropen)) {
    if (LEAF) return 1;
    const double c[3] = {cx, cy, cz};
    return walk_box_inside_sphere(lo, hi, c, src.ropen) ? 1 : -1;
  }
  const double rs = 2.0 * src.soft, rm2 = 2.0 * wd[9];
  const double rflag = __dadd_rn(rs, rmMax);
  const double rflag2 = __dmul_rn(rflag, rflag);
  const double dx = __dsub_rn(wd[6], cx), dy = __dsub_rn(wd[7], cy), dz = __dsub_rn(wd[8], cz);
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  const double rr = __dadd_rn(rs, rm2);
  if ((d2 <= __dmul_rn(rr, rr)) || (dsq <= __dmul_rn(rs, rs)))
    open = dsq <= __dmul_rn(src.ropenMono, src.ropenMono) ? 1 : 0;
  *flaggedOut = (open == 0 && dsq <= rflag2) ? 1 : 0;
  return open;
}

template <bool LEAF>
__device__ __forceinline__ bool walk_node_fast(const WalkTree &t, const WalkParams &p, const WalkPools &pools,
                                               const double *wd, int target, int par, const NodeLists *__restrict__ lists,
                                               double rmMax, const double *shiftTab, const float *shiftTabF, uint4 *rows,
                                               WalkEntry *sring, WalkEntry *scl, WalkEntry *sund, WalkEntry *slp, WalkEntry *cl,
                                               WalkEntry *lp, WalkEntry *und, int lane, int &nc, int &nl, int &nu, int &flC,
                                               int &flL, int &flU, int &myParts, int &myFlagged) {
  constexpr int kMask = kWalkRingS - 1;
  static_assert((kWalkRingS & kMask) == 0, "ring size is a power of two");
  int head = 0, tail = 0;
  nc = nl = nu = 0; myParts = 0; myFlagged = 0;
  flC = flL = flU = 0; /* entries already moved from a full head to the warp's global scratch */
  /* initial checklist into the ring: the parent's undecided nodes, or the root replicas (TreePiece.cpp:3748-3757) */
  if (par < 0) {
    const int side = 2 * p.nReplicas + 1, total = side * side * side;
    if (total > kWalkRingS) return false;
    for (int i = lane; i < total; i += 32) {
      const int x = i / (side * side) - p.nReplicas, y = (i / side) % side - p.nReplicas, z = i % side - p.nReplicas;
      sring[i] = {0, (((x + 3) | ((y + 3) << 3) | ((z + 3) << 6)) << 22)};
    }
    tail = total;
  } else {
    const NodeLists pl = lists[par];
    if (pl.uLen > kWalkRingS) return false;
    const long long *src = reinterpret_cast<const long long *>(pools.undlist + pl.uOff);
    long long *dst = reinterpret_cast<long long *>(sring);
    for (int i = lane; i < pl.uLen; i += 32) dst[i] = __ldg(src + i);
    tail = pl.uLen;
  }
  __syncwarp();
  /* the local box in float: its rounding is part of the error band (WalkNodeRecF) */
  const float lox = (float)wd[0], loy = (float)wd[1], loz = (float)wd[2];
  const float hix = (float)wd[3], hiy = (float)wd[4], hiz = (float)wd[5];
  const WalkFloatTol ftol = *t.ftol;
  const unsigned below = (1u << lane) - 1;
  /* records of a batch: two lanes fetch the two 16-byte halves of one 32-byte float record, 16 records per
   * cp.async instruction, into rows of 48-byte pitch (every lane then reads its own row with two LDS.128) */
  auto stage = [&](uint4 *dstRows, int node) {
    const int sub = lane >> 1, piece = lane & 1;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int src = __shfl_sync(0xffffffffu, node, 16 * j + sub);
      if (src >= 0) cp_async16_ca(&dstRows[(16 * j + sub) * kWalkRowF + piece], reinterpret_cast<const uint4 *>(t.recf + src) + piece);
    }
    cp_async_commit();
  };
  int buf = 0;
  WalkEntry eN = {-1, 0};
  if (lane < tail) eN = sring[lane];
  stage(rows, eN.node);
  while (head < tail) {
    const int batch = min(32, tail - head);
    const bool have = lane < batch;
    WalkEntry e = eN;
    cp_async_wait<0>();
    __syncwarp();
    const uint4 ra = rows[buf * (32 * kWalkRowF) + lane * kWalkRowF], rb = rows[buf * (32 * kWalkRowF) + lane * kWalkRowF + 1];
    /* the part of the next batch that is queued already: its records are requested before this batch is tested */
    const int nextHead = head + batch;
    const int avail = min(32, tail - nextHead);
    eN.node = -1;
    if (avail > 0) { /* warp-uniform */
      if (lane < avail) eN = sring[(nextHead + lane) & kMask];
      stage(rows + (buf ^ 1) * (32 * kWalkRowF), eN.node);
    }
    int open = 0;
    bool flagged = false;
    const int c0 = (int)rb.y, c1 = (int)rb.z, npart = (int)rb.w;
    const bool srcBucket = c0 < 0 && c1 < 0;
    if (have) {
      if (npart < 0) *pools.error = kWalkNotBuilt;
      e.offsetID = target | (kWalkOffsetMask & e.offsetID); /* reEncodeOffset, TreePiece.cpp:3647-3652 */
      bool exact = false;
      if (npart <= 6) {
        open = 1; /* openCriterionNode opens a node of at most six particles whatever the distance */
      } else {
        const float cx = __uint_as_float(ra.x) + shiftTabF[(e.offsetID >> 22) & 7];
        const float cy = __uint_as_float(ra.y) + shiftTabF[(e.offsetID >> 25) & 7];
        const float cz = __uint_as_float(ra.z) + shiftTabF[(e.offsetID >> 28) & 7];
        const float R2 = __uint_as_float(ra.w), tau = __uint_as_float(rb.x);
        const float ax = lox - cx, bx = cx - hix, ay = loy - cy, by = cy - hiy, az = loz - cz, bz = cz - hiz;
        const float dx = fmaxf(fmaxf(ax, bx), 0.0f), dy = fmaxf(fmaxf(ay, by), 0.0f), dz = fmaxf(fmaxf(az, bz), 0.0f);
        const float q = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (fabsf(q - R2) <= fmaf(tau, q + R2, ftol.absTol)) {
          exact = true; /* inside the error band of dsq <= ropen^2 */
        } else if (q < R2) {
          if (LEAF) {
            open = 1;
          } else { /* contained in the sphere?  (Space::contained: the farthest corner) */
            const float wx = fmaxf(fabsf(ax), fabsf(bx)), wy = fmaxf(fabsf(ay), fabsf(by)), wz = fmaxf(fabsf(az), fabsf(bz));
            const float sq = fmaf(wz, wz, fmaf(wy, wy, wx * wx));
            if (fabsf(sq - R2) <= fmaf(tau, sq + R2, ftol.absTol)) exact = true;
            else open = sq < R2 ? 1 : -1;
          }
        } else if (!(q > ftol.softGate)) {
          exact = true; /* within reach of a softening test or of the maybe-softened flag */
        } /* else: accepted, and no softening test can say otherwise */
      }
      if (exact) {
        int fl = 0;
        open = walk_test_double<LEAF>(t.rec, e.node, e.offsetID, wd, shiftTab, rmMax, &fl);
        flagged = fl != 0;
      }
    }
    /* ListCompute::doWork with the LocalOpt table (Opt.h:86-128) */
    const bool toC = have && open == 0;
    const bool toL = have && open != 0 && srcBucket;
    const bool expand = have && open != 0 && !srcBucket && (LEAF || open == 1);
    const bool toU = !LEAF && have && open == -1 && !srcBucket;
    if (toC && flagged) { e.offsetID |= kWalkMaybeSoft; ++myFlagged; }
    if (toL) myParts += npart;
    const unsigned bC = __ballot_sync(0xffffffffu, toC), bL = __ballot_sync(0xffffffffu, toL);
    const unsigned bU = LEAF ? 0u : __ballot_sync(0xffffffffu, toU);
    const unsigned k0 = __ballot_sync(0xffffffffu, expand && c0 >= 0), k1 = __ballot_sync(0xffffffffu, expand && c1 >= 0);
    const int totalKids = __popc(k0) + __popc(k1);
    /* the checklist has no second home here: a node whose live entries outgrow the ring goes to walk_node_general
     * (warp-uniform; nothing has been written yet) */
    if (tail + totalKids - nextHead > kWalkRingS) {
#ifdef CB200_WALK_STATS
      if (lane == 0) atomicAdd(pools.stats + 3, 1);
#endif
      cp_async_wait<0>();
      __syncwarp();
      return false;
    }
    /* a full head is moved to the warp's scratch as a whole (rare: 2 % of the nodes of a uniform box, a quarter
     * of a clustered one), so an append never has to choose between two homes */
    auto flush = [&](WalkEntry *head_, WalkEntry *home, int count, int &flushed) {
      for (int i = lane; i < count - flushed; i += 32) home[flushed + i] = head_[i];
      __syncwarp();
      flushed = count;
    };
    if (nc + __popc(bC) > flC + kWalkClS) flush(scl, cl, nc, flC);
    if (nl + __popc(bL) > flL + kWalkLpFast) flush(slp, lp, nl, flL);
    if (!LEAF && nu + __popc(bU) > flU + kWalkUnS) flush(sund, und, nu, flU);
    /* ordered appends (lane order = checklist order): one ballot per list, one store per lane */
    if (toC) scl[nc - flC + __popc(bC & below)] = e;
    if (toL) slp[nl - flL + __popc(bL & below)] = e;
    if (!LEAF && toU) sund[nu - flU + __popc(bU & below)] = e;
    nc += __popc(bC); nl += __popc(bL); nu += __popc(bU);
    /* children in order 0, 1 behind everything already queued */
    if (expand) {
      int pos = tail + __popc(k0 & below) + __popc(k1 & below);
      if (c0 >= 0) sring[pos++ & kMask] = {c0, e.offsetID};
      if (c1 >= 0) sring[pos & kMask] = {c1, e.offsetID};
    }
    const int oldTail = tail;
    tail += totalKids;
    head = nextHead;
    __syncwarp();
    if (avail < 32 && tail > oldTail) { /* warp-uniform: this batch appended entries of the next one */
      int late = -1;
      const int i = head + lane;
      if (lane >= max(avail, 0) && i < tail) { eN = sring[i & kMask]; late = eN.node; }
      stage(rows + (buf ^ 1) * (32 * kWalkRowF), late);
    }
    buf ^= 1;
  }
  return true;
}

/* the general routine: any list may spill from its shared-memory head into the warp's global scratch
 * (chk, cl, lp, und: kWalkCap entries each); the checklist starts in the parent's slice of the pool */
struct WalkNodeCounts { int nc, nl, nu, parts, flagged; };
__device__ __noinline__ WalkNodeCounts walk_node_general(const WalkTree &t, const WalkParams &p, const WalkPools &pools,
                                                        const WalkNodeRec mine, double lo0, double lo1, double lo2, double hi0,
                                                        double hi1, double hi2, int target, int par,
                                                        const NodeLists *__restrict__ lists, double rmMax, uint4 *rows,
                                                        WalkEntry *sring, WalkEntry *scl, WalkEntry *sund, WalkEntry *slp,
                                                        WalkEntry *chk, int lane) {
  /* by value: a reference would pin the caller's copies (which the fast routine uses) in local memory */
  const double mylo[3] = {lo0, lo1, lo2}, myhi[3] = {hi0, hi1, hi2};
  WalkEntry *cl = chk + kWalkCap, *lp = cl + kWalkCap, *und = lp + kWalkCap;
  const bool myIsBucket = mine.child0 < 0 && mine.child1 < 0;
  int head = 0, tail = 0, nc = 0, nl = 0, nu = 0;
  int myParts = 0, myFlagged = 0; /* per lane; summed over the warp at the end */
  /* initial checklist: the parent's undecided nodes, or the root replicas (TreePiece.cpp:3748-3757).
   * The parent's list is read where it lies (its slice of the undecided pool): FIFO entries
   * [0, nInit) come from there, everything this node appends lives in the warp's ring `chk` --
   * copying the slice into the ring first cost a global read-write pass per node (12% of the
   * level kernel's stall samples sat on those stores, profiles/r02c_ncu_walk_level.json) */
  const WalkEntry *init = chk;
  int nInit = 0;
  if (par < 0) {
    const int side = 2 * p.nReplicas + 1, total = side * side * side;
    for (int i = lane; i < total; i += 32) {
      const int x = i / (side * side) - p.nReplicas, y = (i / side) % side - p.nReplicas, z = i % side - p.nReplicas;
      const WalkEntry r0 = {0, (((x + 3) | ((y + 3) << 3) | ((z + 3) << 6)) << 22)};
      if (i < kWalkRingS) sring[i] = r0;
      else if (i < kWalkCap) chk[i] = r0;
    }
    tail = total;
  } else {
    const NodeLists pl = lists[par];
    init = pools.undlist + pl.uOff;
    nInit = pl.uLen;
    tail = pl.uLen;
  }
  if (tail > kWalkCap) { if (lane == 0) *pools.error = 1; tail = kWalkCap; }
  __syncwarp();
  /* FIFO entry i: the parent's slice, then what this node appended (shared ring, global beyond it) */
  auto fifo = [&](int i) -> WalkEntry {
    if (i < nInit) return init[i];
    const int r = i - nInit;
    return r < kWalkRingS ? sring[r] : chk[r & (kWalkCap - 1)];
  };
  auto fifo_put = [&](int i, WalkEntry v) {
    const int r = i - nInit;
    if (r < kWalkRingS) sring[r] = v;
    else chk[r & (kWalkCap - 1)] = v;
  };

  /* Source records are gathered cooperatively: four lanes fetch the four 16-byte pieces of one
   * 64-byte record with ONE cp.async instruction per 8 records, into an 80-byte-pitch row of
   * shared memory (conflict-free 128-bit reads), and every lane then reads its own row.  A
   * per-lane gather costs one L1 tag lookup per lane and load instruction (five instructions x 32
   * lines per batch: the L1 was the busiest unit, 74%); this way it is 32 lookups per batch. */
  auto stage = [&](uint4 *dstRows, int node) {
    const int sub = lane >> 2, piece = lane & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int src = __shfl_sync(0xffffffffu, node, 8 * j + sub);
      if (src >= 0) cp_async16_ca(&dstRows[(8 * j + sub) * 5 + piece], reinterpret_cast<const uint4 *>(t.rec + src) + piece);
    }
    cp_async_commit();
  };
  /* two row buffers: the records of the next batch are requested before the current batch is
   * tested (entries already in the checklist), the ones this batch appends right after it */
  int buf = 0;
  WalkEntry eN = {-1, 0};
  if (lane < tail) eN = fifo(lane);
  stage(rows, eN.node);
  while (head < tail) {
    const int i = head + lane;
    const bool have = i < tail;
    WalkEntry e = eN;
    const int batch = min(32, tail - head);
    cp_async_wait<0>();
    __syncwarp();
    WalkNodeRec src;
    {
      uint4 *d = reinterpret_cast<uint4 *>(&src);
#pragma unroll
      for (int k = 0; k < 4; ++k) d[k] = rows[lane * 5 + k];
    }
    __syncwarp(); /* one buffer: every lane holds its record before the next batch's records are requested */
    const int iN = i + batch, oldTail = tail;
    const bool early = iN < oldTail;
    eN.node = -1;
    if (head + batch < oldTail) { /* warp-uniform */
      if (early) eN = fifo(iN);
      stage(rows, eN.node);
    }
    int open = 0;
    bool srcBucket = false;
    int c0 = -1, c1 = -1;
    if (have) {
      if (src.first < 0) *pools.error = kWalkNotBuilt;
      e.offsetID = target | (kWalkOffsetMask & e.offsetID); /* reEncodeOffset, TreePiece.cpp:3647-3652 */
      double c[3];
      walk_shifted_cm(src, e.offsetID, p.period, c);
      const double dsq = walk_box_dist2(mylo, myhi, c);
      open = walk_open_criterion(src, c, dsq, mine, mylo, myhi, myIsBucket);
      c0 = src.child0; c1 = src.child1;
      srcBucket = c0 < 0 && c1 < 0;
      if (open == 0) {
        /* openSoftening (gravity.h:251-260) is decided per BUCKET at emit time.  Every bucket
         * below this node has its box and centre of mass inside this node's box, so when the
         * cell's softening sphere, grown by the largest bucket softening, misses the box no
         * bucket can see the cell softened: emit then skips the test and the 64-byte gather */
        const double rflag = __dadd_rn(2.0 * src.soft, rmMax);
        if (dsq <= __dmul_rn(rflag, rflag)) {
          e.offsetID |= kWalkMaybeSoft;
          ++myFlagged;
        }
      } else if (srcBucket) {
        myParts += src.last - src.first + 1;
      }
    }
    /* ListCompute::doWork with the LocalOpt table (Opt.h:86-128) */
    const bool toC = have && open == 0;
    const bool toL = have && open != 0 && srcBucket;
    const bool expand = have && open != 0 && !srcBucket && (open == 1 || myIsBucket);
    const bool toU = have && open != 0 && !srcBucket && !expand;
    /* ordered appends (lane order = checklist order): one ballot per list, one store per lane --
     * the three lists are consecutive kWalkCap-slices of the warp's scratch */
    const unsigned below = (1u << lane) - 1;
    const unsigned bC = __ballot_sync(0xffffffffu, toC), bL = __ballot_sync(0xffffffffu, toL),
                   bU = __ballot_sync(0xffffffffu, toU);
    if (toC | toL | toU) {
      const int pos = toC ? nc + __popc(bC & below) : (toL ? nl + __popc(bL & below) : nu + __popc(bU & below));
      const int capS = toC ? kWalkClS : (toL ? kWalkLpS : kWalkUnS);
      WalkEntry *dst = pos < capS ? (toC ? scl : (toL ? slp : sund)) : (toC ? cl : (toL ? lp : und));
      if (pos < kWalkCap) dst[pos] = e;
      else *pools.error = 1;
    }
    nc += __popc(bC); nl += __popc(bL); nu += __popc(bU);
    /* children in order 0, 1 behind everything already queued */
    const unsigned k0 = __ballot_sync(0xffffffffu, expand && c0 >= 0), k1 = __ballot_sync(0xffffffffu, expand && c1 >= 0);
    const int totalKids = __popc(k0) + __popc(k1);
    if (tail - head - batch + totalKids > kWalkCap) { if (lane == 0) *pools.error = 1; break; }
    int pos = tail + __popc(k0 & below) + __popc(k1 & below);
    if (expand) {
      if (c0 >= 0) fifo_put(pos++, {c0, e.offsetID});
      if (c1 >= 0) fifo_put(pos, {c1, e.offsetID});
    }
    head += batch;
    tail += totalKids;
    __syncwarp();
    if (tail > oldTail && oldTail < head + 32) { /* warp-uniform: this batch appended entries of the next one */
      int late = -1;
      if (!early && iN < tail) { eN = fifo(iN); late = eN.node; }
      stage(rows, late);
    }
    buf ^= 1;
  }

  return {nc, nl, nu, myParts, myFlagged};
}

/* The nodes of a level that have a bucket in [bucketLo, bucketHi): nodes of a level are in SFC order and their
 * bucket ranges ascend, so it is one index range per level (two binary searches).  A rank that owns 1/8 of the
 * buckets visits 1/8 of every deep level; walking the whole level to find them (and writing a "not visited"
 * record for the other 7/8) was 40 % of a rank's level kernels at eight ranks (profiles/r02s). */
struct WalkLevels { int start[66]; int n; };
__global__ void walk_level_ranges_kernel(WalkTree t, WalkLevels lv, int bucketLo, int bucketHi, int2 *__restrict__ range) {
  const int l = threadIdx.x;
  if (l >= lv.n) return;
  const int a0 = lv.start[l], b0 = lv.start[l + 1];
  int a = a0, b = b0;
  while (a < b) { /* first node whose buckets end behind bucketLo */
    const int mid = a + ((b - a) >> 1);
    if (t.bucketFirst[mid] + t.bucketCount[mid] <= bucketLo) a = mid + 1; else b = mid;
  }
  const int lo = a;
  b = b0;
  while (a < b) { /* first node whose buckets start at or behind bucketHi */
    const int mid = a + ((b - a) >> 1);
    if (t.bucketFirst[mid] < bucketHi) a = mid + 1; else b = mid;
  }
  range[l] = make_int2(lo, a - lo);
}

/* One level of the local tree: nodes [range->x, range->x + range->y) (walk_level_ranges_kernel).  scratch: per
 * warp 4 x kWalkCap entries (checklist, clist, lplist, undlist) for walk_node_general. */
#ifndef CB200_WALK_MINB
#define CB200_WALK_MINB 5
#endif
__global__ void __launch_bounds__(kWalkWarps * 32, CB200_WALK_MINB)
walk_level_kernel(WalkTree t, WalkParams p, const int2 *__restrict__ range, NodeLists *__restrict__ lists, WalkPools pools,
                  WalkEntry *__restrict__ scratch, int generalOnly, int chunk) {
  /* chunk: entries a warp reserves from the cell pool at a time (half of it from the bucket pool, a quarter from
   * the undecided pool): the host scales it with the nodes a warp will see on this level */
  const int lo = range->x, n = range->y;
  const int lane = threadIdx.x & 31;
  const int warpGlobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int totalWarps = (gridDim.x * blockDim.x) >> 5;
  extern __shared__ __align__(16) unsigned char walkSmem[];
  __shared__ double shiftTab[8];
  __shared__ float shiftTabF[8];
  __shared__ double mineD[kWalkWarps][kWalkMineDoubles];
  __shared__ unsigned long long poolCursor[kWalkWarps][6]; /* {cursor, end} of the warp's chunk of the three pools */
  if (threadIdx.x < kWalkWarps * 6) poolCursor[threadIdx.x / 6][threadIdx.x % 6] = 0ull;
  unsigned long long *wpool = poolCursor[threadIdx.x >> 5];
  if (threadIdx.x < 8) {
    shiftTab[threadIdx.x] = __dmul_rn((double)((int)threadIdx.x - 3), p.period); /* walk_shifted_cm */
    shiftTabF[threadIdx.x] = (float)shiftTab[threadIdx.x];
  }
  __syncthreads();
  double *wd = mineD[threadIdx.x >> 5];
  unsigned char *wsm = walkSmem + (size_t)(threadIdx.x >> 5) * kWalkWarpSmem;
  uint4 *rows = reinterpret_cast<uint4 *>(wsm);
  WalkEntry *sring = reinterpret_cast<WalkEntry *>(wsm + kWalkRowBytes);
  WalkEntry *scl = sring + kWalkRingS, *sund = scl + kWalkClS, *slp = sund + kWalkUnS;
  WalkEntry *chk = scratch + (size_t)warpGlobal * 4 * kWalkCap;
  WalkEntry *cl = chk + kWalkCap, *lp = cl + kWalkCap, *und = lp + kWalkCap;

  const double rmMax = 2.0 * __longlong_as_double((long long)*t.softMaxBits);
  for (int w = warpGlobal; w < n; w += totalWarps) {
    const int my = lo + w;
    NodeLists out = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int target = 0;
    bool go = walk_node_active(t, p, my, target);
    const int par = t.parent[my];
    out.parent = par;
    if (par >= 0) {
      const NodeLists up = lists[par];
      if (go) go = up.visited && up.uLen > 0; /* descend only under a non-empty undecided list */
      out.pathCells = up.pathCells; out.pathParts = up.pathParts; out.pathFlagged = up.pathFlagged;
    }
    if (!go) {
      if (lane == 0) lists[my] = out; /* not visited: the path totals of the last visited ancestor */
      continue;
    }
    target &= kWalkBucketMask;
    /* the local node in double goes to shared memory (walk_test_double, walk_node_general read it from there) */
    __syncwarp();
    if (lane < 3) { wd[lane] = t.boxlo[3 * (size_t)my + lane]; wd[3 + lane] = t.boxhi[3 * (size_t)my + lane]; }
    const WalkNodeRecF mineF = t.recf[my];
    if (lane == 0) {
      const WalkNodeRec mine = t.rec[my];
      wd[6] = mine.cx; wd[7] = mine.cy; wd[8] = mine.cz; wd[9] = mine.soft;
    }
    __syncwarp();
    const bool myIsBucket = mineF.child0 < 0 && mineF.child1 < 0;
    int nc = 0, nl = 0, nu = 0, myParts = 0, myFlagged = 0;
    int flC = 0, flL = 0, flU = 0;
    bool done = false;
    if (!generalOnly) {
      if (myIsBucket)
        done = walk_node_fast<true>(t, p, pools, wd, target, par, lists, rmMax, shiftTab, shiftTabF, rows, sring, scl, sund, slp, cl,
                                    lp, und, lane, nc, nl, nu, flC, flL, flU, myParts, myFlagged);
      else
        done = walk_node_fast<false>(t, p, pools, wd, target, par, lists, rmMax, shiftTab, shiftTabF, rows, sring, scl, sund, slp, cl,
                                     lp, und, lane, nc, nl, nu, flC, flL, flU, myParts, myFlagged);
    }
    if (!done) {
      const WalkNodeRec mine = t.rec[my];
      const WalkNodeCounts r = walk_node_general(t, p, pools, mine, wd[0], wd[1], wd[2], wd[3], wd[4], wd[5], target,
                                                 par, lists, rmMax, rows, sring, scl, sund, slp, chk, lane);
      nc = r.nc; nl = r.nl; nu = r.nu; myParts = r.parts; myFlagged = r.flagged;
    }

    /* Exact-size slices of the pools, cut from CHUNKS the warp reserves: three atomics per node were a full
     * round trip at the end of every node (5 % of the leaf level's stall samples); a chunk lasts for dozens of nodes
     * and its cursor lives in shared memory.  What is left of a chunk when the launch ends is lost (the host sizes
     * the pools for it). */
    unsigned long long oc = 0, ol = 0, ou = 0;
    if (lane == 0) {
      const int need[3] = {nc, nl, nu};
      const unsigned long long cap[3] = {pools.capC, pools.capL, pools.capU};
      unsigned long long off[3];
      bool ok = true;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        unsigned long long cur = wpool[2 * k], end = wpool[2 * k + 1];
        if (cur + (unsigned long long)need[k] > end) {
          const unsigned long long want = (unsigned long long)(need[k] > (chunk >> k) ? need[k] : (chunk >> k));
          cur = atomicAdd(pools.used + k, want);
          end = cur + want;
          if (end > cap[k]) ok = false;
        }
        off[k] = cur;
        wpool[2 * k] = cur + (unsigned long long)need[k];
        wpool[2 * k + 1] = end;
      }
      oc = off[0]; ol = off[1]; ou = off[2];
      if (!ok) { *pools.error = 2; nc = nl = nu = 0; }
    }
    oc = __shfl_sync(0xffffffffu, oc, 0); ol = __shfl_sync(0xffffffffu, ol, 0); ou = __shfl_sync(0xffffffffu, ou, 0);
    nc = __shfl_sync(0xffffffffu, nc, 0); nl = __shfl_sync(0xffffffffu, nl, 0); nu = __shfl_sync(0xffffffffu, nu, 0);
    if (done) { /* fast routine: the first fl* entries were moved to the scratch, the rest is the head */
      for (int i = lane; i < nc; i += 32) pools.clist[oc + i] = i < flC ? cl[i] : scl[i - flC];
      for (int i = lane; i < nl; i += 32) pools.lplist[ol + i] = i < flL ? lp[i] : slp[i - flL];
      for (int i = lane; i < nu; i += 32) pools.undlist[ou + i] = i < flU ? und[i] : sund[i - flU];
    } else { /* general routine: the head is the first part, the scratch holds what did not fit */
      for (int i = lane; i < nc; i += 32) pools.clist[oc + i] = i < kWalkClS ? scl[i] : cl[i];
      for (int i = lane; i < nl; i += 32) pools.lplist[ol + i] = i < kWalkLpS ? slp[i] : lp[i];
      for (int i = lane; i < nu; i += 32) pools.undlist[ou + i] = i < kWalkUnS ? sund[i] : und[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      myParts += __shfl_xor_sync(0xffffffffu, myParts, o);
      myFlagged += __shfl_xor_sync(0xffffffffu, myFlagged, o);
    }
    out.cOff = (int)oc; out.cLen = nc; out.lOff = (int)ol; out.lLen = nl; out.uOff = (int)ou; out.uLen = nu;
    out.visited = 1;
    out.pathCells += nc; out.pathParts += myParts; out.pathFlagged += myFlagged;
    out.ownParts = myParts;
    if (lane == 0) lists[my] = out;
    __syncwarp();
  }
}

/* path root -> deepest visited ancestor-or-self of the bucket's node; returns its length */
__device__ __forceinline__ int walk_path(const WalkTree &t, const NodeLists *lists, int bucketNode, int *path) {
  int chain[64], n = 0;
  for (int v = bucketNode; v >= 0 && n < 64; v = t.parent[v]) chain[n++] = v;
  int len = 0;
  for (int k = n - 1; k >= 0; --k) { /* root first; stop below the lowest node */
    if (!lists[chain[k]].visited) break;
    path[len++] = chain[k];
  }
  return len;
}

/* path[b * stride + l] = the ancestor-or-self of bucket b's node at tree level l (-1 below the bucket).
 * One THREAD per bucket descends from the root by bucket ranges; neighbouring threads read the same
 * nodes, so the descent runs out of L1.  emit_fill then has every node of a bucket's path at once
 * (lane = level) instead of climbing parent links one dependent load at a time: the climb was what
 * bounded it (22 levels x one L2 round trip per bucket-warp, 7.2 ms at 256^3). */
__global__ void walk_paths_kernel(WalkTree t, int bucketLo, int bucketHi, int stride, int *__restrict__ path) {
  const int b = bucketLo + blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bucketHi) return;
  int *row = path + (size_t)(b - bucketLo) * stride;
  int v = 0, l = 0;
  while (v >= 0 && l < stride) {
    row[l++] = v;
    const int c0 = t.child0[v], c1 = t.child1[v];
    if (c0 < 0 && c1 < 0) break;
    v = (c1 >= 0 && (c0 < 0 || b >= t.bucketFirst[c1])) ? c1 : c0;
  }
  for (; l < stride; ++l) row[l] = -1;
}

/* counts[b] = {cells, softened cells, expanded particle entries} of bucket b (0 outside the range).
 * Two kernels: a bucket whose path holds no flagged cell has its sizes in its node's path totals -- one THREAD
 * per bucket (a warp per bucket spent 0.67 ms at 256^3 with one lane working); the others (a cell may be
 * softened for them) are collected in `flagged` and counted by one warp each, which tests the flagged cells. */
__global__ void emit_count_kernel(WalkTree t, WalkParams p, const NodeLists *__restrict__ lists,
                                  int *__restrict__ nCell, int *__restrict__ nSoft, int *__restrict__ nPart,
                                  int *__restrict__ starts, int *__restrict__ sizes, int *__restrict__ flagged,
                                  int *__restrict__ nFlagged) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= t.numBuckets) return;
  const int bn = t.bucketNode[b];
  starts[b] = t.first[bn]; sizes[b] = t.last[bn] - t.first[bn] + 1;
  int cells = 0, part = 0;
  if (walk_bucket_active(p, b)) {
    const NodeLists tot = lists[bn];
    cells = tot.pathCells;
    part = tot.pathParts;
    if (tot.pathFlagged > 0) flagged[atomicAdd(nFlagged, 1)] = b;
  }
  nCell[b] = cells; nSoft[b] = 0; nPart[b] = part;
}
__global__ void __launch_bounds__(kWalkWarps * 32)
emit_count_flagged_kernel(WalkTree t, WalkParams p, const NodeLists *__restrict__ lists, WalkPools pools,
                          int *__restrict__ nCell, int *__restrict__ nSoft, const int *__restrict__ flagged,
                          const int *__restrict__ nFlagged) {
  const int lane = threadIdx.x & 31;
  const int n = *nFlagged;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += (gridDim.x * blockDim.x) >> 5) {
    const int b = flagged[w];
    const int bn = t.bucketNode[b];
    /* only flagged cells can be softened for this bucket: test them (gravity.h:251-260) */
    int path[64], soft = 0;
    const int plen = walk_path(t, lists, bn, path);
    const WalkNodeRec mm = t.rec[bn];
    const double *lo = t.boxlo + 3 * (size_t)bn, *hi = t.boxhi + 3 * (size_t)bn;
    for (int k = 0; k < plen; ++k) {
      const NodeLists nl = lists[path[k]];
      for (int i = lane; i < nl.cLen; i += 32) {
        const WalkEntry e = pools.clist[nl.cOff + i];
        if (e.offsetID & kWalkMaybeSoft) {
          const WalkNodeRec m = t.rec[e.node];
          double c[3];
          walk_shifted_cm(m, e.offsetID, p.period, c);
          if (walk_open_softening(m, c, mm, lo, hi)) ++soft;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) soft += __shfl_xor_sync(0xffffffffu, soft, o);
    if (lane == 0) { nCell[b] -= soft; nSoft[b] = soft; }
  }
}

/* 64-bit totals of the three count arrays: the markers are 32-bit (the ABI's ILCell lists are
 * addressed by int markers), so a range whose lists exceed 2^31-1 entries must be refused */
__global__ void walk_totals_kernel(const int *__restrict__ counts, int nb1, unsigned long long *__restrict__ totals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    unsigned long long v = i < nb1 ? (unsigned long long)counts[(size_t)k * nb1 + i] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(totals + k, v);
  }
}

/* 32 source buckets -> their particles, one ILCell per particle (GenericList<ILPart>::serialize,
 * Compute.cpp:1174-1187).  Every lane holds one source bucket (first particle f, count cnt, replica
 * code); `incl` is the inclusive scan of cnt.  The runs are laid out in shared memory and copied out
 * with full-warp, consecutive 8-byte stores: writing each lane's run straight to global memory is 32
 * strided stores per instruction (a quarter of every 32-byte sector), and the particle lists are 40%
 * of all list bytes. */
constexpr int kEmitStage = 512;
constexpr int kEmitPieces = 64; /* <= 32-entry pieces of the cell slices of one bucket's path (emit_fill) */
__device__ __forceinline__ void emit_expand(ILCell *__restrict__ dst, ILCell *stage, int f, int cnt, int code, int incl,
                                            int lane) {
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total <= kEmitStage) {
    ILCell *mine = stage + (incl - cnt);
    for (int j = 0; j < cnt; ++j) {
      ILCell o;
      o.index = f + j; o.offsetID = code;
      mine[j] = o;
    }
    __syncwarp();
    long long *d64 = reinterpret_cast<long long *>(dst);
    const long long *s64 = reinterpret_cast<const long long *>(stage);
    for (int o = lane; o < total; o += 32) d64[o] = s64[o];
    __syncwarp();
  } else {
    ILCell *mine = dst + (incl - cnt);
    for (int j = 0; j < cnt; ++j) {
      ILCell o;
      o.index = f + j; o.offsetID = code;
      mine[j] = o;
    }
  }
}

/* markers are exclusive prefix sums of the counts (numBuckets + 1 entries each) */
__global__ void __launch_bounds__(kWalkWarps * 32)
emit_fill_kernel(WalkTree t, WalkParams p, const NodeLists *__restrict__ lists, WalkPools pools,
                 const int *__restrict__ cellMark, const int *__restrict__ softMark, const int *__restrict__ partMark,
                 ILCell *__restrict__ cellOut, ILCell *__restrict__ softOut, ILCell *__restrict__ partOut,
                 const int *__restrict__ pathTable, int pathStride) {
  const int lane = threadIdx.x & 31;
  const int b = p.bucketLo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  __shared__ __align__(16) ILCell stageAll[kWalkWarps * kEmitStage];
  __shared__ long long pentAll[kWalkWarps * 64];
  __shared__ int4 pieceAll[kWalkWarps * kEmitPieces];
  ILCell *stage = stageAll + (threadIdx.x >> 5) * kEmitStage;
  long long *pent = pentAll + (threadIdx.x >> 5) * 64;
  int4 *piece = pieceAll + (threadIdx.x >> 5) * kEmitPieces;
  if (b >= p.bucketHi || !walk_bucket_active(p, b)) return;
  const int bn = t.bucketNode[b];
  if (pathTable && lists[bn].pathFlagged == 0) {
    /* The whole path at once: lane l holds the list record of the bucket's ancestor at level l (two
     * dependent loads for the bucket instead of one per level).  Cells: every level's slice goes to the
     * place its path total names.  Particle buckets: the levels' entries are first concatenated in a
     * small staging array, so that the expansion per particle (Compute.cpp:1174-1187) runs on full
     * batches of 32 source buckets instead of one ragged batch per level. */
    const long long *__restrict__ clist64 = reinterpret_cast<const long long *>(pools.clist);
    const long long *__restrict__ lplist64 = reinterpret_cast<const long long *>(pools.lplist);
    long long *__restrict__ cellOut64 = reinterpret_cast<long long *>(cellOut + cellMark[b]);
    int wp = partMark[b], held = 0;
    auto expand_batch = [&](int n) { /* the first n (<= 32) staged source buckets -> their particles */
      int f = 0, cnt = 0, code = 0;
      if (lane < n) {
        const long long e64 = pent[lane];
        const int2 fl = __ldg(reinterpret_cast<const int2 *>(&t.rec[(int)e64].first));
        f = fl.x;
        cnt = fl.y - f + 1;
        code = (int)(e64 >> 32) & kWalkOffsetMask; /* encodeOffset(0, x, y, z) */
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      emit_expand(partOut + wp, stage, f, cnt, code, incl, lane);
      wp += __shfl_sync(0xffffffffu, incl, 31);
    };
    const int *row = pathTable + (size_t)(b - p.bucketLo) * pathStride;
    for (int l0 = 0; l0 < pathStride; l0 += 32) {
      const int v = l0 + lane < pathStride ? __ldg(row + l0 + lane) : -1;
      int nC = 0, srcC = 0, dstC = 0, nL = 0, srcL = 0;
      if (v >= 0) {
        const uint4 *q = reinterpret_cast<const uint4 *>(lists + v);
        const uint4 a = __ldg(q), c = __ldg(q + 1), d = __ldg(q + 2);
        /* NodeLists: {cOff, cLen, lOff, lLen | uOff, uLen, visited, pathCells | pathParts, pathFlagged, ownParts, parent} */
        if (c.z) { nC = (int)a.y; srcC = (int)a.x; dstC = (int)c.w - nC; nL = (int)a.w; srcL = (int)a.z; }
        (void)d;
      }
      /* Cells: the levels' slices are cut into pieces of <= 32 entries, described in a small table, and copied
       * four pieces at a time -- four independent loads in flight per lane.  One slice after the other (a load,
       * then the store that waits for it, 17 times per bucket) left the kernel waiting on its own loads:
       * long-scoreboard 61 % of the stall samples, 1050 of 1580 instructions per bucket in that loop
       * (profiles/r02l_ncu_emit_fill_4M.json). */
      const int myPieces = (nC + 31) >> 5;
      int inclP = myPieces;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inclP, o);
        if (lane >= o) inclP += u;
      }
      const int nPieces = __shfl_sync(0xffffffffu, inclP, 31);
      if (nPieces <= kEmitPieces) {
        for (int k = 0; k < myPieces; ++k)
          piece[inclP - myPieces + k] = make_int4(srcC + 32 * k, dstC + 32 * k, min(32, nC - 32 * k), 0);
        __syncwarp();
        for (int j0 = 0; j0 < nPieces; j0 += 4) {
          int4 d[4];
          long long v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) d[q] = j0 + q < nPieces ? piece[j0 + q] : make_int4(0, 0, 0, 0);
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = lane < d[q].z ? __ldg(clist64 + d[q].x + lane) : 0ll;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (lane < d[q].z) cellOut64[d[q].y + lane] = v[q]; /* {node, offsetID} = {index, offsetID} */
        }
        __syncwarp();
      } else { /* a path with more pieces than the table holds: slice by slice */
        unsigned mc = __ballot_sync(0xffffffffu, nC > 0);
        while (mc) {
          const int sl = __ffs(mc) - 1;
          mc &= mc - 1;
          const int n = __shfl_sync(0xffffffffu, nC, sl), src = __shfl_sync(0xffffffffu, srcC, sl),
                    dst = __shfl_sync(0xffffffffu, dstC, sl);
          for (int i = lane; i < n; i += 32) cellOut64[dst + i] = __ldg(clist64 + src + i);
        }
      }
      unsigned ml = __ballot_sync(0xffffffffu, nL > 0);
      while (ml) {
        const int sl = __ffs(ml) - 1;
        ml &= ml - 1;
        const int n = __shfl_sync(0xffffffffu, nL, sl), src = __shfl_sync(0xffffffffu, srcL, sl);
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int take = min(32, n - i0);
          if (lane < take) pent[held + lane] = __ldg(lplist64 + src + i0 + lane);
          held += take;
          __syncwarp();
          if (held >= 32) {
            expand_batch(32);
            const long long rest = lane < held - 32 ? pent[32 + lane] : 0;
            __syncwarp();
            if (lane < held - 32) pent[lane] = rest;
            held -= 32;
            __syncwarp();
          }
        }
      }
    }
    if (held > 0) expand_batch(held);
    return;
  }
  if (lists[bn].pathFlagged == 0) {
    /* no cell on the path can be softened for any bucket: every level's entries go to a place the
     * path totals name, so one climb bucket -> root (a single chain of parent links) does it all,
     * cells (Compute.cpp:1653-1743) and particle buckets expanded per particle
     * (Compute.cpp:1823-1863, 1174-1187) level by level */
    const int cbase = cellMark[b], pbase = partMark[b];
    const long long *__restrict__ clist64 = reinterpret_cast<const long long *>(pools.clist);
    const long long *__restrict__ lplist64 = reinterpret_cast<const long long *>(pools.lplist);
    NodeLists nl = lists[bn];
    for (;;) {
      /* the parent's record is requested before this level's entries are copied: one dependent
       * load per level instead of three (parent link -> record -> entries) */
      const int par = nl.parent;
      NodeLists up = nl;
      if (par >= 0) up = lists[par];
      if (nl.visited) {
      long long *co = reinterpret_cast<long long *>(cellOut + cbase + (nl.pathCells - nl.cLen));
      for (int i = lane; i < nl.cLen; i += 32) co[i] = __ldg(clist64 + nl.cOff + i); /* {node, offsetID} = {index, offsetID} */
      int wp = pbase + (nl.pathParts - nl.ownParts);
      for (int i0 = 0; i0 < nl.lLen; i0 += 32) {
        const int i = i0 + lane;
        int f = 0, cnt = 0, code = 0;
        if (i < nl.lLen) {
          const long long e64 = __ldg(lplist64 + nl.lOff + i);
          const int2 fl = __ldg(reinterpret_cast<const int2 *>(&t.rec[(int)e64].first));
          f = fl.x;
          cnt = fl.y - f + 1;
          code = (int)(e64 >> 32) & kWalkOffsetMask; /* encodeOffset(0, x, y, z) */
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += u;
        }
        emit_expand(partOut + wp, stage, f, cnt, code, incl, lane);
        wp += __shfl_sync(0xffffffffu, incl, 31);
      }
      }
      if (par < 0) break;
      nl = up;
    }
    return;
  }
  int path[64];
  const int plen = walk_path(t, lists, bn, path);
  const WalkNodeRec mm = t.rec[bn];
  const double *lo = t.boxlo + 3 * (size_t)bn, *hi = t.boxhi + 3 * (size_t)bn;
  int wc = cellMark[b], ws = softMark[b], wp = partMark[b];
  for (int k = 0; k < plen; ++k) { /* cells, level 0 .. maxlevel (Compute.cpp:1653-1743) */
    const NodeLists nl = lists[path[k]];
    for (int i0 = 0; i0 < nl.cLen; i0 += 32) {
      const int i = i0 + lane;
      const bool have = i < nl.cLen;
      WalkEntry e = {0, 0};
      bool isSoft = false;
      if (have) {
        e = pools.clist[nl.cOff + i];
        if (e.offsetID & kWalkMaybeSoft) {
          const WalkNodeRec m = t.rec[e.node];
          double c[3];
          walk_shifted_cm(m, e.offsetID, p.period, c);
          isSoft = walk_open_softening(m, c, mm, lo, hi);
          e.offsetID &= ~kWalkMaybeSoft;
        }
      }
      const unsigned bs = __ballot_sync(0xffffffffu, have && isSoft), bc = __ballot_sync(0xffffffffu, have && !isSoft);
      const unsigned below = (1u << lane) - 1;
      if (have) {
        ILCell o;
        o.index = e.node; o.offsetID = e.offsetID;
        if (isSoft) softOut[ws + __popc(bs & below)] = o;
        else cellOut[wc + __popc(bc & below)] = o;
      }
      ws += __popc(bs);
      wc += __popc(bc);
    }
  }
  for (int k = 0; k < plen; ++k) { /* particle buckets, expanded per particle (Compute.cpp:1823-1863, 1174-1187) */
    const NodeLists nl = lists[path[k]];
    /* 32 source buckets at a time: every lane fetches one entry and its particle range, a warp
     * scan places the ranges, then each lane writes its own <= bucket-size entries */
    for (int i0 = 0; i0 < nl.lLen; i0 += 32) {
      const int i = i0 + lane;
      int f = 0, cnt = 0, code = 0;
      if (i < nl.lLen) {
        const WalkEntry e = pools.lplist[nl.lOff + i];
        const WalkNodeRec &src = t.rec[e.node];
        f = src.first;
        cnt = src.last - f + 1;
        code = e.offsetID & kWalkOffsetMask; /* encodeOffset(0, x, y, z) */
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      emit_expand(partOut + wp, stage, f, cnt, code, incl, lane);
      wp += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

/* multistep sets.  flags[i] = particle i (tree order) has rung >= activeRung (the Ewald markers
 * are the indices of the set flags, Ewald.cpp:416-437); a bucket is active when any of its
 * particles is (GenericTreeNode::rungs is the maximum below the node; Compute.cpp:1278) */
__global__ void active_particle_flags_kernel(const unsigned char *__restrict__ rung, const int *__restrict__ order,
                                             int n, int activeRung, unsigned char *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = rung[order ? order[i] : i] >= activeRung;
}
__global__ void active_bucket_flags_kernel(const unsigned char *__restrict__ flags, const int *__restrict__ starts,
                                           const int *__restrict__ sizes, int nb, unsigned char *__restrict__ active,
                                           int *__restrict__ nActive) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  bool a = false;
  if (b < nb) {
    const int f = starts[b], c = sizes[b];
    for (int j = 0; j < c; ++j) a = a || flags[f + j];
    active[b] = a;
  }
  const unsigned m = __ballot_sync(0xffffffffu, a);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nActive, __popc(m));
}

/* softened cells as ad-hoc source particles: every node as {cm, M | soft} */
__global__ void nodes_as_particles_kernel(const double *__restrict__ mom, PackedPart *__restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *m = mom + (size_t)i * 27;
  PackedPart q;
  q.x = (real)m[3]; q.y = (real)m[4]; q.z = (real)m[5]; q.mass = (real)m[2];
  q.soft = (real)m[1]; q.pad0 = q.pad1 = q.pad2 = 0;
  out[i] = q;
}

}  // namespace cb200
#endif
