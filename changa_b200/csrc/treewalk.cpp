/* treewalk.cpp -- host side of the hot path's INPUT: the tree, its moments and
 * the per-bucket interaction lists, in the exact shape ChaNGa's TreePiece /
 * DataManager hand to the GPU entry points.  Built as libcb200_host.so (no
 * CUDA needed to load it).  It exists so that the force kernels can be fed and
 * measured with the lists a real ChaNGa step produces (bench.py, tests) and is
 * the single-TreePiece statement of:
 *
 *   keys + tree    binary-oct tree split at box midpoints by key bit
 *                  (GenericTreeNode.h:473-598), bucket when
 *                  lastParticle-firstParticle < maxBucketSize (Compute.cpp:2476-2477)
 *   moments        makeBucket / operator+= / calculateRadius* via moments_build.cuh
 *                  (GenericTreeNode.h:219-256, MultipoleMoments.h:207-352,460-526)
 *   node order     breadth first, children 0 then 1, empty children skipped
 *                  = nodeArrayIndex (DataManager.cpp:797-828)
 *   walk           Stadel double walk: LocalTargetWalk::dft (TreeWalk.cpp:308-397),
 *                  ListCompute::doWork (Compute.cpp:690-884), LocalOpt (Opt.h:86-128),
 *                  openCriterionNode / openSoftening (gravity.h:652-723, 251-260),
 *                  27-replica root enqueue (TreePiece.cpp:3741-3766)
 *   emission       ListCompute::stateReady, GPU branch (Compute.cpp:1608-1863):
 *                  per active bucket, level by level, cells then particle buckets;
 *                  softened cells split off for a softened-monopole evaluation
 *                  (Compute.cpp:1683-1699)
 *   serialize      GenericList<T>::serialize (Compute.cpp:1034-1218)
 *
 * The reference walks buckets one after another, restarting at the child of
 * the least common ancestor with the previous bucket and re-using the lists of
 * the levels above (TreePiece.cpp:4757-4775).  Those per-level lists are a pure
 * function of the path root -> node, so here the same walk is a recursion over
 * the local tree in which both children of a node inherit the node's undecided
 * list; subtrees are independent OpenMP tasks.  List contents and order per
 * bucket are identical to the sequential walk, including the low 22 bits of
 * offsetID (the bucket that was the walk target when the entry was decided =
 * the first active bucket under the deciding local node).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

#include "moments_build.cuh"
#include "ewald_setup.cuh"

using cb200::MomentNode;

namespace {

constexpr int kKeyBitsPerDim = 21;
constexpr int kKeyBits = 63;
constexpr int kOffsetMask = 0x1ff << 22; /* TreePiece.cpp:3649 */
constexpr int kBucketMask = (1 << 22) - 1;

inline int encode_offset(int reqID, int x, int y, int z) { /* TreePiece.cpp:3631-3644 */
  return reqID | (((x + 3) | ((y + 3) << 3) | ((z + 3) << 6)) << 22);
}
inline int reencode_offset(int reqID, int offsetID) { return reqID | (kOffsetMask & offsetID); }

struct Tree {
  int n = 0, maxBucket = 12;
  double rootlo[3], roothi[3];
  std::vector<int> order;      /* sorted position -> caller's particle index */
  std::vector<uint64_t> keys;  /* sorted */
  std::vector<double> pos, mass, soft;
  /* nodes, breadth first */
  std::vector<int> child0, child1, parent, first, last, level, bucketFirst, bucketCount;
  std::vector<int> levelStart;
  std::vector<double> geolo, geohi, boxlo, boxhi;
  std::vector<MomentNode> mom;
  std::vector<int> bucketNode; /* bucket index (particle order) -> node */
  int numNodes() const { return (int)child0.size(); }
  int numBuckets() const { return (int)bucketNode.size(); }
  bool isBucket(int i) const { return child0[i] < 0 && child1[i] < 0; }
};

uint64_t morton_key(const double *p, const double *lo, const double *inv) {
  uint32_t q[3];
  for (int d = 0; d < 3; ++d) {
    double f = (p[d] - lo[d]) * inv[d];
    if (f < 0.0) f = 0.0;
    double s = f * (double)(1u << kKeyBitsPerDim);
    uint32_t v = s >= (double)(1u << kKeyBitsPerDim) ? (1u << kKeyBitsPerDim) - 1 : (uint32_t)s;
    q[d] = v;
  }
  uint64_t k = 0;
  for (int b = kKeyBitsPerDim - 1; b >= 0; --b)
    k = (k << 3) | (uint64_t)(((q[0] >> b) & 1) << 2 | ((q[1] >> b) & 1) << 1 | ((q[2] >> b) & 1));
  return k;
}

void build_topology(Tree &t) {
  const int n = t.n;
  auto push = [&](int par, int lvl, int f, int l, const double *glo, const double *ghi) {
    t.child0.push_back(-1); t.child1.push_back(-1); t.parent.push_back(par);
    t.first.push_back(f); t.last.push_back(l); t.level.push_back(lvl);
    for (int d = 0; d < 3; ++d) { t.geolo.push_back(glo[d]); t.geohi.push_back(ghi[d]); }
    return (int)t.child0.size() - 1;
  };
  push(-1, 0, 0, n - 1, t.rootlo, t.roothi);
  t.levelStart.push_back(0);
  int lo = 0;
  for (int lvl = 0;; ++lvl) {
    const int hi = t.numNodes();
    if (lo == hi) break;
    t.levelStart.push_back(hi);
    for (int i = lo; i < hi; ++i) {
      const int f = t.first[i], l = t.last[i];
      /* Compute.cpp:2476-2477; the key runs out of bits at NodeKeyBits-3 */
      if (l - f < t.maxBucket || lvl >= kKeyBits - 3) continue;
      const int bit = kKeyBits - 1 - lvl;
      const uint64_t mask = (uint64_t)1 << bit;
      /* first particle of [f,l] with the split bit set (GenericTreeNode.h:573-578) */
      const uint64_t *b = t.keys.data() + f, *e = t.keys.data() + l + 1;
      const uint64_t probe = (t.keys[l] & (~(uint64_t)0 << bit)) | mask;
      int split;
      if ((t.keys[f] & mask) == (t.keys[l] & mask))
        split = (t.keys[f] & mask) ? f : l + 1;
      else
        split = (int)(std::lower_bound(b, e, probe) - t.keys.data());
      double glo[3], ghi[3];
      for (int d = 0; d < 3; ++d) { glo[d] = t.geolo[3 * i + d]; ghi[d] = t.geohi[3 * i + d]; }
      const int dim = lvl % 3;
      const double mid = 0.5 * (ghi[dim] + glo[dim]);
      if (split > f) {
        double h2[3] = {ghi[0], ghi[1], ghi[2]};
        h2[dim] = mid;
        t.child0[i] = push(i, lvl + 1, f, split - 1, glo, h2);
      }
      if (split <= l) {
        double l2[3] = {glo[0], glo[1], glo[2]};
        l2[dim] = mid;
        t.child1[i] = push(i, lvl + 1, split, l, l2, ghi);
      }
    }
    lo = hi;
  }
  /* buckets in particle order = ChaNGa's bucketList (depth-first build order) */
  const int nn = t.numNodes();
  for (int i = 0; i < nn; ++i)
    if (t.isBucket(i)) t.bucketNode.push_back(i);
  std::sort(t.bucketNode.begin(), t.bucketNode.end(), [&](int a, int b) { return t.first[a] < t.first[b]; });
  t.bucketFirst.assign(nn, 0);
  t.bucketCount.assign(nn, 0);
  for (int b = 0; b < t.numBuckets(); ++b) { t.bucketFirst[t.bucketNode[b]] = b; t.bucketCount[t.bucketNode[b]] = 1; }
  for (int i = nn - 1; i >= 0; --i) {
    if (t.isBucket(i)) continue;
    const int c0 = t.child0[i], c1 = t.child1[i];
    t.bucketFirst[i] = c0 >= 0 ? t.bucketFirst[c0] : t.bucketFirst[c1];
    t.bucketCount[i] = (c0 >= 0 ? t.bucketCount[c0] : 0) + (c1 >= 0 ? t.bucketCount[c1] : 0);
  }
}

void build_boxes_and_moments(Tree &t) {
  const int nn = t.numNodes();
  t.boxlo.assign(3 * (size_t)nn, 0.0);
  t.boxhi.assign(3 * (size_t)nn, 0.0);
  t.mom.resize(nn);
  const int nl = (int)t.levelStart.size() - 1;
  for (int lvl = nl - 1; lvl >= 0; --lvl) {
    const int lo = t.levelStart[lvl], hi = t.levelStart[lvl + 1];
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = lo; i < hi; ++i) {
      double *bl = &t.boxlo[3 * (size_t)i], *bh = &t.boxhi[3 * (size_t)i];
      MomentNode m;
      if (t.isBucket(i)) {
        for (int d = 0; d < 3; ++d) { bl[d] = HUGE_VAL; bh[d] = -HUGE_VAL; }
        for (int p = t.first[i]; p <= t.last[i]; ++p)
          for (int d = 0; d < 3; ++d) {
            bl[d] = std::min(bl[d], t.pos[3 * (size_t)p + d]);
            bh[d] = std::max(bh[d], t.pos[3 * (size_t)p + d]);
          }
        cb200::node_make_bucket(m, t.pos.data(), t.mass.data(), t.soft.data(), t.first[i], t.last[i],
                                &t.geolo[3 * (size_t)i], &t.geohi[3 * (size_t)i]);
      } else {
        cb200::node_clear(m);
        for (int d = 0; d < 3; ++d) { bl[d] = HUGE_VAL; bh[d] = -HUGE_VAL; }
        const int ch[2] = {t.child0[i], t.child1[i]};
        for (int c : ch) {
          if (c < 0) continue;
          cb200::node_add_node(m, t.mom[c]);
          for (int d = 0; d < 3; ++d) {
            bl[d] = std::min(bl[d], t.boxlo[3 * (size_t)c + d]);
            bh[d] = std::max(bh[d], t.boxhi[3 * (size_t)c + d]);
          }
        }
        cb200::node_radius_from_box(m, bl, bh);
      }
      t.mom[i] = m;
    }
  }
}

/* ------------------------------------------------------------------ the walk */
struct OffsetNode { int node; int offsetID; };

struct LevelLists {
  std::vector<OffsetNode> clist, lplist, undlist;
  const LevelLists *up = nullptr;
};

struct BucketSpan { int thread; size_t cell, part, soft; int nCell, nPart, nSoft; };

struct Arena {
  std::vector<int> cell; /* ILCell pairs */
  std::vector<int> part; /* ILPart triples */
  std::vector<int> soft; /* ILCell pairs: node index, offsetID */
  long long opened = 0, tested = 0;
  long long visited = 0, maxChk = 0, maxC = 0, maxL = 0, maxU = 0, sumC = 0, sumL = 0, sumU = 0;
};

struct Walk {
  const Tree &t;
  double theta, thetaMono, period;
  int nReplicas;
  std::vector<int> activeCount; /* active buckets under each node */
  std::vector<int> firstActive; /* first active bucket under each node */
  std::vector<unsigned char> bucketActive;
  std::vector<BucketSpan> spans;
  std::vector<Arena> arenas;
  int taskCutoff = 64; /* buckets under a node below which no task is spawned */

  Walk(const Tree &tree) : t(tree) {}

  static bool box_sphere(const double *lo, const double *hi, const double *c, double r) {
    /* Space::intersect(box, sphere), restated in-tree at CUDAMoments.cu:137-159 */
    double dsq = 0.0, delta;
    const double rsq = r * r;
    for (int d = 0; d < 3; ++d) {
      if ((delta = lo[d] - c[d]) > 0) dsq += delta * delta;
      else if ((delta = c[d] - hi[d]) > 0) dsq += delta * delta;
      if (rsq < dsq) return false;
    }
    return dsq <= rsq;
  }
  static bool box_inside_sphere(const double *lo, const double *hi, const double *c, double r) {
    /* Space::contained(box, sphere): the farthest corner is inside */
    double s = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double a = std::fabs(lo[d] - c[d]), b = std::fabs(hi[d] - c[d]);
      const double w = a > b ? a : b;
      s += w * w;
    }
    return s <= r * r;
  }
  void shifted_cm(int node, int offsetID, double *c) const {
    const MomentNode &m = t.mom[node];
    c[0] = m.cm[0] + (((offsetID >> 22) & 7) - 3) * period;
    c[1] = m.cm[1] + (((offsetID >> 25) & 7) - 3) * period;
    c[2] = m.cm[2] + (((offsetID >> 28) & 7) - 3) * period;
  }
  /* gravity.h:251-260 */
  bool open_softening(int node, const double *c, int my) const {
    const MomentNode &m = t.mom[node], &mm = t.mom[my];
    const double rs = 2.0 * m.soft, rm = 2.0 * mm.soft;
    const double dx = mm.cm[0] - c[0], dy = mm.cm[1] - c[1], dz = mm.cm[2] - c[2];
    if (dx * dx + dy * dy + dz * dz <= (rs + rm) * (rs + rm)) return true;
    return box_sphere(&t.boxlo[3 * (size_t)my], &t.boxhi[3 * (size_t)my], c, rs);
  }
  /* gravity.h:652-723: 1 = open for everything below `my`, -1 = undecided, 0 = accept */
  int open_criterion(int node, int offsetID, int my) const {
    if (t.last[node] - t.first[node] + 1 <= 6) return 1; /* nMinParticleNode */
    const MomentNode &m = t.mom[node];
    const double geom = 2.0 / std::sqrt(3.0); /* TreeNode.h:35 */
    double radius = geom * m.radius / theta;
    if (radius < m.radius) radius = m.radius;
    double c[3];
    shifted_cm(node, offsetID, c);
    const double *lo = &t.boxlo[3 * (size_t)my], *hi = &t.boxhi[3 * (size_t)my];
    if (box_sphere(lo, hi, c, radius)) {
      if (t.isBucket(my)) return 1;
      return box_inside_sphere(lo, hi, c, radius) ? 1 : -1;
    }
    if (!open_softening(node, c, my)) return 0; /* a hexadecapole interaction */
    radius = geom * m.radius / thetaMono;        /* softened: accept only as a monopole */
    return box_sphere(lo, hi, c, radius) ? 1 : 0;
  }

  /* ListCompute::doWork for one source node against the local node `my` */
  void process(int node, int reqID, int my, LevelLists &L, std::vector<OffsetNode> &chk, Arena &A) const {
    const int open = open_criterion(node, reqID, my);
    A.tested++;
    if (open == 0) { /* LocalOpt[0][Internal|Bucket] = COMPUTE */
      L.clist.push_back({node, reqID});
      return;
    }
    A.opened++;
    if (t.isBucket(node)) { /* LocalOpt[1][Bucket] = KEEP_LOCAL_BUCKET */
      L.lplist.push_back({node, reqID});
      return;
    }
    if (open == 1 || t.isBucket(my)) { /* KEEP + CONTAIN: children go on this level's checklist */
      if (t.child0[node] >= 0) chk.push_back({t.child0[node], reqID});
      if (t.child1[node] >= 0) chk.push_back({t.child1[node], reqID});
    } else {
      L.undlist.push_back({node, reqID}); /* INTERSECT: the children of `my` decide */
    }
  }

  void emit(int my, const LevelLists &L) {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    Arena &A = arenas[tid];
    const LevelLists *chain[80];
    int nl = 0;
    for (const LevelLists *p = &L; p; p = p->up) chain[nl++] = p;
    for (int b = t.bucketFirst[my]; b < t.bucketFirst[my] + t.bucketCount[my]; ++b) {
      if (!bucketActive[b]) continue;
      const int bn = t.bucketNode[b];
      BucketSpan s;
      s.thread = tid; s.cell = A.cell.size(); s.part = A.part.size(); s.soft = A.soft.size();
      for (int k = nl - 1; k >= 0; --k) /* level 0 .. maxlevel (Compute.cpp:1653-1743) */
        for (const OffsetNode &e : chain[k]->clist) {
          double c[3];
          shifted_cm(e.node, e.offsetID, c);
          std::vector<int> &dst = open_softening(e.node, c, bn) ? A.soft : A.cell; /* Compute.cpp:1683-1699 */
          dst.push_back(e.node);
          dst.push_back(e.offsetID);
        }
      for (int k = nl - 1; k >= 0; --k) /* Compute.cpp:1823-1863 */
        for (const OffsetNode &e : chain[k]->lplist) {
          A.part.push_back(t.first[e.node]);
          A.part.push_back(e.offsetID & kOffsetMask); /* encodeOffset(0, x, y, z) */
          A.part.push_back(t.last[e.node] - t.first[e.node] + 1);
        }
      s.nCell = (int)((A.cell.size() - s.cell) / 2);
      s.nPart = (int)((A.part.size() - s.part) / 3);
      s.nSoft = (int)((A.soft.size() - s.soft) / 2);
      spans[b] = s;
    }
  }

  void dft(int my, const LevelLists *up) {
    LevelLists L;
    L.up = up;
    std::vector<OffsetNode> chk;
#ifdef _OPENMP
    Arena &A = arenas[omp_get_thread_num()];
#else
    Arena &A = arenas[0];
#endif
    const int target = firstActive[my] & kBucketMask;
    if (!up) { /* TreePiece.cpp:3748-3757 */
      for (int x = -nReplicas; x <= nReplicas; ++x)
        for (int y = -nReplicas; y <= nReplicas; ++y)
          for (int z = -nReplicas; z <= nReplicas; ++z) chk.push_back({0, encode_offset(0, x, y, z)});
    } else { /* the parent's undecided nodes first (TreeWalk.cpp:331-355) */
      for (const OffsetNode &e : up->undlist) process(e.node, reencode_offset(target, e.offsetID), my, L, chk, A);
    }
    for (size_t head = 0; head < chk.size(); ++head) {
      const OffsetNode e = chk[head];
      process(e.node, reencode_offset(target, e.offsetID), my, L, chk, A);
    }
    A.visited++;
    A.maxChk = std::max<long long>(A.maxChk, (long long)chk.size());
    A.maxC = std::max<long long>(A.maxC, (long long)L.clist.size()); A.sumC += L.clist.size();
    A.maxL = std::max<long long>(A.maxL, (long long)L.lplist.size()); A.sumL += L.lplist.size();
    A.maxU = std::max<long long>(A.maxU, (long long)L.undlist.size()); A.sumU += L.undlist.size();
    std::vector<OffsetNode>().swap(chk);
    if (!L.undlist.empty()) {
      const int ch[2] = {t.child0[my], t.child1[my]};
      for (int c : ch) {
        if (c < 0 || activeCount[c] == 0) continue;
        if (t.bucketCount[c] >= taskCutoff) {
#pragma omp task default(shared) firstprivate(c)
          dft(c, &L);
        } else {
          dft(c, &L);
        }
      }
#pragma omp taskwait
    } else {
      emit(my, L); /* lowestNode: every bucket below shares these lists */
    }
  }
};

struct Lists {
  std::vector<int> cell, part, soft;
  std::vector<long long> cellMark, partMark, softMark;
  long long opened = 0, tested = 0;
  long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

}  // namespace

extern "C" {

void *cb200h_tree_build(const double *pos, const double *mass, const double *soft, int n, int maxBucket,
                        const double *rootlo, const double *roothi) {
  Tree *t = new Tree;
  t->n = n;
  t->maxBucket = maxBucket;
  double inv[3];
  for (int d = 0; d < 3; ++d) { t->rootlo[d] = rootlo[d]; t->roothi[d] = roothi[d]; inv[d] = 1.0 / (roothi[d] - rootlo[d]); }
  std::vector<uint64_t> k(n);
#pragma omp parallel for
  for (int i = 0; i < n; ++i) k[i] = morton_key(pos + 3 * (size_t)i, rootlo, inv);
  t->order.resize(n);
  std::iota(t->order.begin(), t->order.end(), 0);
  auto cmp = [&](int a, int b) { return k[a] != k[b] ? k[a] < k[b] : a < b; };
#ifdef _OPENMP
  __gnu_parallel::sort(t->order.begin(), t->order.end(), cmp);
#else
  std::sort(t->order.begin(), t->order.end(), cmp);
#endif
  t->keys.resize(n); t->pos.resize(3 * (size_t)n); t->mass.resize(n); t->soft.resize(n);
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    const int o = t->order[i];
    t->keys[i] = k[o];
    for (int d = 0; d < 3; ++d) t->pos[3 * (size_t)i + d] = pos[3 * (size_t)o + d];
    t->mass[i] = mass[o];
    t->soft[i] = soft[o];
  }
  build_topology(*t);
  build_boxes_and_moments(*t);
  return t;
}

void cb200h_tree_free(void *h) { delete (Tree *)h; }

/* out[0]=nodes out[1]=buckets out[2]=levels out[3]=particles */
void cb200h_tree_sizes(void *h, int *out) {
  Tree *t = (Tree *)h;
  out[0] = t->numNodes(); out[1] = t->numBuckets(); out[2] = (int)t->levelStart.size() - 1; out[3] = t->n;
}

/* any pointer may be NULL.  parts: n x 5 {mass, soft, x, y, z} in sorted order
 * (CompactPartData); moments: nodes x 27 (CudaMultipoleMoments order), double */
void cb200h_tree_export(void *h, int *order, double *parts, double *moments, int *child0, int *child1,
                        int *first, int *last, int *levelStart, double *geolo, double *geohi, double *boxlo,
                        double *boxhi, int *bucketNode, int *bucketStarts, int *bucketSizes) {
  Tree *t = (Tree *)h;
  const int n = t->n, nn = t->numNodes(), nb = t->numBuckets();
  if (order) memcpy(order, t->order.data(), sizeof(int) * n);
  if (parts)
    for (int i = 0; i < n; ++i) {
      double *p = parts + 5 * (size_t)i;
      p[0] = t->mass[i]; p[1] = t->soft[i];
      p[2] = t->pos[3 * (size_t)i]; p[3] = t->pos[3 * (size_t)i + 1]; p[4] = t->pos[3 * (size_t)i + 2];
    }
  if (moments)
    for (int i = 0; i < nn; ++i) cb200::node_export(t->mom[i], moments + 27 * (size_t)i);
  if (child0) memcpy(child0, t->child0.data(), sizeof(int) * nn);
  if (child1) memcpy(child1, t->child1.data(), sizeof(int) * nn);
  if (first) memcpy(first, t->first.data(), sizeof(int) * nn);
  if (last) memcpy(last, t->last.data(), sizeof(int) * nn);
  if (levelStart) memcpy(levelStart, t->levelStart.data(), sizeof(int) * t->levelStart.size());
  if (geolo) memcpy(geolo, t->geolo.data(), sizeof(double) * 3 * nn);
  if (geohi) memcpy(geohi, t->geohi.data(), sizeof(double) * 3 * nn);
  if (boxlo) memcpy(boxlo, t->boxlo.data(), sizeof(double) * 3 * nn);
  if (boxhi) memcpy(boxhi, t->boxhi.data(), sizeof(double) * 3 * nn);
  if (bucketNode) memcpy(bucketNode, t->bucketNode.data(), sizeof(int) * nb);
  for (int b = 0; b < nb; ++b) {
    const int i = t->bucketNode[b];
    if (bucketStarts) bucketStarts[b] = t->first[i];
    if (bucketSizes) bucketSizes[b] = t->last[i] - t->first[i] + 1;
  }
}

/* per node: parent, first bucket beneath, buckets beneath (inputs of the device walk) */
void cb200h_tree_export_links(void *h, int *parent, int *bucketFirst, int *bucketCount) {
  Tree *t = (Tree *)h;
  const size_t nn = t->numNodes();
  if (parent) memcpy(parent, t->parent.data(), sizeof(int) * nn);
  if (bucketFirst) memcpy(bucketFirst, t->bucketFirst.data(), sizeof(int) * nn);
  if (bucketCount) memcpy(bucketCount, t->bucketCount.data(), sizeof(int) * nn);
}

/* Walk buckets [bucketLo, bucketHi) (a rank's SFC range; the whole tree is the
 * source).  bucketActive: one byte per bucket of the tree, or NULL = all active. */
void *cb200h_walk(void *h, double theta, int nReplicas, double period, const unsigned char *bucketActive,
                  int bucketLo, int bucketHi) {
  Tree *t = (Tree *)h;
  Walk w(*t);
  w.theta = theta;
  w.thetaMono = theta * theta * theta * theta; /* TreePiece.cpp:5036 */
  w.period = period;
  w.nReplicas = nReplicas;
  const int nb = t->numBuckets(), nn = t->numNodes();
  w.bucketActive.assign(nb, 0);
  for (int b = std::max(bucketLo, 0); b < std::min(bucketHi, nb); ++b) w.bucketActive[b] = bucketActive ? bucketActive[b] : 1;
  w.activeCount.assign(nn, 0);
  w.firstActive.assign(nn, 0);
  for (int i = nn - 1; i >= 0; --i) {
    if (t->isBucket(i)) {
      w.activeCount[i] = w.bucketActive[t->bucketFirst[i]];
      w.firstActive[i] = t->bucketFirst[i];
    } else {
      const int c0 = t->child0[i], c1 = t->child1[i];
      const int a0 = c0 >= 0 ? w.activeCount[c0] : 0, a1 = c1 >= 0 ? w.activeCount[c1] : 0;
      w.activeCount[i] = a0 + a1;
      w.firstActive[i] = a0 ? w.firstActive[c0] : (a1 ? w.firstActive[c1] : 0);
    }
  }
  BucketSpan none;
  memset(&none, 0, sizeof none);
  w.spans.assign(nb, none);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  w.arenas.resize(nthreads);
  if (nn > 0 && w.activeCount[0] > 0) {
#pragma omp parallel
#pragma omp single
    w.dft(0, nullptr);
  }
  Lists *L = new Lists;
  L->cellMark.assign(nb + 1, 0); L->partMark.assign(nb + 1, 0); L->softMark.assign(nb + 1, 0);
  for (int b = 0; b < nb; ++b) {
    L->cellMark[b + 1] = L->cellMark[b] + w.spans[b].nCell;
    L->partMark[b + 1] = L->partMark[b] + w.spans[b].nPart;
    L->softMark[b + 1] = L->softMark[b] + w.spans[b].nSoft;
  }
  L->cell.resize(2 * (size_t)L->cellMark[nb]);
  L->part.resize(3 * (size_t)L->partMark[nb]);
  L->soft.resize(2 * (size_t)L->softMark[nb]);
#pragma omp parallel for schedule(dynamic, 256)
  for (int b = 0; b < nb; ++b) {
    const BucketSpan &s = w.spans[b];
    const Arena &A = w.arenas[s.thread];
    if (s.nCell) memcpy(&L->cell[2 * (size_t)L->cellMark[b]], &A.cell[s.cell], sizeof(int) * 2 * s.nCell);
    if (s.nPart) memcpy(&L->part[3 * (size_t)L->partMark[b]], &A.part[s.part], sizeof(int) * 3 * s.nPart);
    if (s.nSoft) memcpy(&L->soft[2 * (size_t)L->softMark[b]], &A.soft[s.soft], sizeof(int) * 2 * s.nSoft);
  }
  for (const Arena &A : w.arenas) {
    L->opened += A.opened; L->tested += A.tested;
    L->stats[0] += A.visited; L->stats[1] = std::max(L->stats[1], A.maxChk); L->stats[2] = std::max(L->stats[2], A.maxC);
    L->stats[3] = std::max(L->stats[3], A.maxL); L->stats[4] = std::max(L->stats[4], A.maxU);
    L->stats[5] += A.sumC; L->stats[6] += A.sumL; L->stats[7] += A.sumU;
  }
  return L;
}

void cb200h_lists_free(void *h) { delete (Lists *)h; }

/* visited local nodes, max checklist / clist / lplist / undlist length of one node, summed lengths */
void cb200h_lists_stats(void *h, long long *out) { memcpy(out, ((Lists *)h)->stats, sizeof(long long) * 8); }

/* out: cells, part buckets, softened cells, expanded particle entries, MAC tests, opened */
void cb200h_lists_sizes(void *h, long long *out) {
  Lists *L = (Lists *)h;
  out[0] = (long long)L->cell.size() / 2;
  out[1] = (long long)L->part.size() / 3;
  out[2] = (long long)L->soft.size() / 2;
  long long e = 0;
  for (size_t i = 2; i < L->part.size(); i += 3) e += L->part[i];
  out[3] = e;
  out[4] = L->tested;
  out[5] = L->opened;
}

/* markers have one entry per bucket of the tree plus one (empty buckets included) */
void cb200h_lists_export(void *h, int *cell, long long *cellMark, int *part, long long *partMark, int *soft,
                         long long *softMark) {
  Lists *L = (Lists *)h;
  if (cell && !L->cell.empty()) memcpy(cell, L->cell.data(), sizeof(int) * L->cell.size());
  if (part && !L->part.empty()) memcpy(part, L->part.data(), sizeof(int) * L->part.size());
  if (soft && !L->soft.empty()) memcpy(soft, L->soft.data(), sizeof(int) * L->soft.size());
  if (cellMark) memcpy(cellMark, L->cellMark.data(), sizeof(long long) * L->cellMark.size());
  if (partMark) memcpy(partMark, L->partMark.data(), sizeof(long long) * L->partMark.size());
  if (softMark) memcpy(softMark, L->softMark.data(), sizeof(long long) * L->softMark.size());
}

/* GenericList<ILPart>::serialize (Compute.cpp:1174-1187): one ILCell{index, off}
 * per source particle; expandedMark[b] = entries before bucket b */
void cb200h_expand_part_list(const int *part, const long long *partMark, int numBuckets, int *expanded,
                             long long *expandedMark) {
  expandedMark[0] = 0;
  for (int b = 0; b < numBuckets; ++b) {
    long long e = 0;
    for (long long j = partMark[b]; j < partMark[b + 1]; ++j) e += part[3 * j + 2];
    expandedMark[b + 1] = expandedMark[b] + e;
  }
  if (!expanded) return;
#pragma omp parallel for schedule(dynamic, 256)
  for (int b = 0; b < numBuckets; ++b) {
    long long o = expandedMark[b];
    for (long long j = partMark[b]; j < partMark[b + 1]; ++j) {
      const int start = part[3 * j], off = part[3 * j + 1], num = part[3 * j + 2];
      for (int k = 0; k < num; ++k, ++o) { expanded[2 * o] = start + k; expanded[2 * o + 1] = off; }
    }
  }
}

/* Ewald set-up of one step (TreePiece::EwaldInit, Ewald.cpp:285-375), C twin of
 * changa_b200/ewald_tables.py: complete root moments (MomcData order, 32 values) and the h-loop
 * table rows {hx, hy, hz, hCfac, hSfac}.  With T2, T3, T4 the complete moment tensors and h an
 * integer wave vector: hCfac = -(g0 M + g2 T2:hh/2 + g4 T4::hhhh/24), hSfac = -(g3 T3:.hhh/6),
 * g_k = g0 (2 pi/L)^k with signs (+,+,-,-,+,+).  Returns the number of rows (<= cap written). */
int cb200h_ewald_tables(const double *root, double L, double dEwhCut, double *momc, double *ewt, int cap) {
  double comp[125];
  cb200::ewald_complete_moments(root, comp, momc); /* ewald_setup.cuh: shared with the device set-up kernel */
  std::vector<int> h(3 * (size_t)(cap > 0 ? cap : 1));
  const int n = cb200::ewald_h_vectors(dEwhCut, h.data(), cap);
  for (int i = 0; i < n && i < cap; ++i)
    cb200::ewald_h_row(comp, root[2], L, h[3 * i], h[3 * i + 1], h[3 * i + 2], ewt + 5 * (size_t)i);
  return n;
}

/* testdata/ppartt.c:34-65: srand(seed); x, y, z = xmin + rand()/RAND_MAX (xmax - xmin) per
 * particle, in that order (glibc rand(): the recipe is only reproducible with it) */
void cb200h_uniform_box(int seed, long long n, double xmin, double xmax, double *pos) {
  srand((unsigned)seed);
  for (long long i = 0; i < 3 * n; ++i) pos[i] = xmin + (double)rand() / (double)RAND_MAX * (xmax - xmin);
}

int cb200h_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

} /* extern "C" */
