/* ewald_ref.cpp -- the reference's OWN Ewald.cpp (QEVAL, TreePiece::EwaldInit, TreePiece::BucketEwald: the CPU
 * Ewald sum), compiled here unmodified, behind a small C interface.
 *
 * TEST INFRASTRUCTURE ONLY (built by oracle/Makefile into oracle/_ref/libgravity_ref.so when /root/reference is
 * present; nothing is copied out of the reference tree).  It pins the oracle's orc_ewald_root_momc / orc_ewald_init
 * / orc_ewald to the code they restate (tests/test_oracle_pins.py; tests/golden/gravity_kat.npz carries the pin to
 * the GPU box).
 *
 * Ewald.cpp includes one header, ParallelGravity.h, which is Charm++ from top to bottom: its include guard is
 * defined below, so the preprocessor skips it.  What the three routines touch of class TreePiece is declared here
 * under the reference's names -- the root node and its complete moments, the particle array, the period, the active
 * rung, the h-loop table -- together with inert stand-ins for the five Charm++ calls at the end of EwaldInit (it
 * posts a message to start the Ewald phase; here the message is dropped).  EwaldGPU / EwaldGPUComplete are empty
 * without SPCUDA (Ewald.cpp:387-552). */
#define PARALLELGRAVITY_H
#define TREENODE_H
#define GENERICTREENODE_H
#define CMK_SSE 0
#ifndef HEXADECAPOLE
#define HEXADECAPOLE 1
#endif

#include <assert.h>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "reference_types.h" /* oracle/shim/gravity */

using namespace Tree;

#define CkAssert(x) assert(x)
#define CK_QUEUEING_IFIFO 0
struct EwaldMsg {
  bool fromInit;
  int priority[8];
  void *operator new(std::size_t bytes, int /* priority bits */) { return ::operator new(bytes); }
  void operator delete(void *p, int) { ::operator delete(p); }
  void operator delete(void *p) { ::operator delete(p); }
};
inline void *CkPriorityPtr(EwaldMsg *m) { return m->priority; }
inline void CkSetQueueing(EwaldMsg *, int) {}
struct TreePieceElementProxy {
  void calculateEwald(EwaldMsg *m) { delete m; }
};
struct TreePieceArrayProxy {
  TreePieceElementProxy operator[](int) { return TreePieceElementProxy(); }
};
unsigned int numTreePieces = 1;
extern cosmoType theta, thetaMono; /* gravity_ref.cpp */

class TreePiece {
 public:
  MOMC momcRoot;
  GenericTreeNode *root;
  GravityParticle *myParticles;
  Vector3D<cosmoType> fPeriod;
  int activeRung;
  EWT *ewt;
  int nEwhLoop, nMaxEwhLoop;
  double dEwhCut;
  bool bBucketsInited;
  int numChunks, thisIndex;
  TreePieceArrayProxy thisProxy;
  void BucketEwald(GenericTreeNode *req, int nReps, double fEwCut);
  void EwaldInit();
  void EwaldGPU();
  void EwaldGPUComplete();
};

#include "Ewald.cpp" /* the reference's, unmodified */
#include "gravity.h" /* the reference's, unmodified (inline: for gref_force_step below; theta lives in gravity_ref.cpp) */

namespace {

struct Piece {
  TreePiece tp;
  GenericTreeNode root;
  Piece(const double *root27, double L, double dEwhCut) {
    cb200_fill_node(root, root27, nullptr, nullptr, 0, 0, -1, 1000);
    std::memset(&tp.momcRoot, 0, sizeof tp.momcRoot);
    tp.root = &root;
    tp.myParticles = nullptr;
    tp.fPeriod = Vector3D<cosmoType>(L, L, L);
    tp.activeRung = 0;
    tp.nMaxEwhLoop = 100; /* ParallelGravity.h: the table starts at 100 rows and doubles */
    tp.ewt = new EWT[tp.nMaxEwhLoop];
    tp.nEwhLoop = 0;
    tp.dEwhCut = dEwhCut;
    tp.bBucketsInited = true;
    tp.numChunks = 1; tp.thisIndex = 0;
    tp.EwaldInit();
  }
  ~Piece() { delete[] tp.ewt; }
};

}  // namespace

extern "C" {

/* EwaldInit on the root cell record (27 doubles, gravity_oracle.c CM_*): the complete root moments (32 doubles in
 * MOMC order, moments.h:40-48) and the h-loop table (rows hx hy hz hCfac hSfac); returns the number of rows */
int eref_init(const double *root27, double L, double dEwhCut, double *momc32, double *ewtRows, int cap) {
  Piece p(root27, L, dEwhCut);
  std::memcpy(momc32, &p.tp.momcRoot, 32 * sizeof(double));
  for (int i = 0; i < p.tp.nEwhLoop && i < cap; ++i) {
    const EWT &e = p.tp.ewt[i];
    double *r = ewtRows + 5 * (size_t)i;
    r[0] = e.hx; r[1] = e.hy; r[2] = e.hz; r[3] = e.hCfac; r[4] = e.hSfac;
  }
  return p.tp.nEwhLoop;
}

/* EwaldInit + BucketEwald on the particle rows [first, last] ({mass, soft, x, y, z}); adds to vars rows
 * {ax, ay, az, pot, dtGrav} as the reference adds to treeAcceleration and potential */
void eref_bucket_ewald(const double *root27, double L, double dEwhCut, double fEwCut, int nReps, const double *part,
                       int first, int last, const unsigned char *rung, int activeRung, double *vars) {
  Piece p(root27, L, dEwhCut);
  std::vector<GravityParticle> g((size_t)last + 1);
  for (int j = first; j <= last; ++j) {
    const double *r = part + (size_t)j * 5, *v = vars + (size_t)j * 5;
    g[j].mass = r[0]; g[j].soft = r[1]; g[j].position = Vector3D<cosmoType>(r[2], r[3], r[4]);
    g[j].treeAcceleration = Vector3D<cosmoType>(v[0], v[1], v[2]);
    g[j].potential = v[3]; g[j].dtGrav = v[4]; g[j].interMass = 0.0;
    g[j].rung = rung ? rung[j] : 0;
  }
  p.tp.myParticles = g.data();
  p.tp.activeRung = activeRung;
  GenericTreeNode req;
  req.type = Tree::Bucket; req.firstParticle = first; req.lastParticle = last; req.particleCount = (unsigned)(last - first + 1);
  p.tp.BucketEwald(&req, nReps, fEwCut);
  for (int j = first; j <= last; ++j) {
    double *v = vars + (size_t)j * 5;
    v[0] = g[j].treeAcceleration.x; v[1] = g[j].treeAcceleration.y; v[2] = g[j].treeAcceleration.z;
    v[3] = g[j].potential;
  }
}

/* One whole force evaluation of buckets [b0, b1) by the reference's own CPU routines, the way its CPU path runs
 * them (ListCompute::stateReady without CUDA, Compute.cpp:1608-1863): for every bucket, nodeBucketForce for each cell
 * of its list (which itself sends a cell whose softening reaches the bucket to the particle routine),
 * partBucketForce for every particle of each source bucket of its particle list, then BucketEwald.  Lists in the
 * per-bucket form the walk produces (tree.py: cell / soft = {node, offsetID}, part = {first particle, offset code,
 * count}, markers over all buckets).  OpenMP over buckets (their particles are disjoint).  This is the CPU arm of
 * bench.py (`--impl reference`, cpu_baseline kind "reference"): gref_step_create holds the particles and the tree
 * records as the reference's objects (and runs EwaldInit once), gref_step_run is the timed part. */
struct gref_step {
  std::vector<GravityParticle> P;
  std::vector<GenericTreeNode> nodes; /* moments, particle range; tight box */
  std::vector<int> bucketNode;
  Piece *piece;
  double period, fEwCut;
  int nReps;
};

gref_step *gref_step_create(const double *part, int n, const double *cells, int numNodes, const int *bucketNode, int numBuckets,
                            const double *boxlo, const double *boxhi, const int *nodeFirst, const int *nodeLast, double period,
                            int ewald, double dEwhCut, double fEwCut, int nReps) {
  gref_step *st = new gref_step;
  st->P.resize((size_t)n);
  for (int j = 0; j < n; ++j) {
    const double *r = part + (size_t)j * 5;
    GravityParticle &p = st->P[j];
    p.mass = r[0]; p.soft = r[1]; p.position = Vector3D<cosmoType>(r[2], r[3], r[4]);
    p.treeAcceleration = Vector3D<cosmoType>(0, 0, 0);
    p.potential = 0; p.dtGrav = 0; p.interMass = 0.0; p.rung = 0;
  }
  st->nodes.resize((size_t)numNodes);
  for (int i = 0; i < numNodes; ++i)
    cb200_fill_node(st->nodes[i], cells + (size_t)i * 27, boxlo + 3 * (size_t)i, boxhi + 3 * (size_t)i, 0, nodeFirst[i], nodeLast[i],
                    (unsigned)(nodeLast[i] - nodeFirst[i] + 1));
  st->bucketNode.assign(bucketNode, bucketNode + numBuckets);
  for (int b = 0; b < numBuckets; ++b) st->nodes[bucketNode[b]].type = Tree::Bucket;
  st->piece = ewald ? new Piece(cells, period, dEwhCut) : nullptr;
  if (st->piece) st->piece->tp.myParticles = st->P.data();
  st->period = period; st->fEwCut = fEwCut; st->nReps = nReps;
  return st;
}

void gref_step_run(gref_step *st, int b0, int b1, const int *cell, const long long *cellMark, const int *plist,
                   const long long *partMark, const int *soft, const long long *softMark, int nThreads) {
  const double period = st->period;
  auto offset_of = [period](int code) { /* decodeOffset, ParallelGravity.h:2064-2073 */
    return Vector3D<cosmoType>((((code >> 22) & 7) - 3) * period, (((code >> 25) & 7) - 3) * period,
                               (((code >> 28) & 7) - 3) * period);
  };
  GravityParticle *P = st->P.data();
  if (nThreads < 1) nThreads = 1;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nThreads)
  for (int b = b0; b < b1; ++b) {
    GenericTreeNode *req = &st->nodes[st->bucketNode[b]];
    for (int j = req->firstParticle; j <= req->lastParticle; ++j) { /* a step starts from zeroed accumulators */
      P[j].treeAcceleration = Vector3D<cosmoType>(0, 0, 0);
      P[j].potential = 0; P[j].dtGrav = 0; P[j].interMass = 0.0;
    }
    for (int pass = 0; pass < 2; ++pass) {
      const int *list = pass ? soft : cell;
      const long long *mark = pass ? softMark : cellMark;
      for (long long e = mark[b]; e < mark[b + 1]; ++e)
        nodeBucketForce(&st->nodes[list[2 * e]], req, P, offset_of(list[2 * e + 1]), 0);
    }
    for (long long e = partMark[b]; e < partMark[b + 1]; ++e) {
      const int first = plist[3 * e], count = plist[3 * e + 2];
      const Vector3D<cosmoType> off = offset_of(plist[3 * e + 1]);
      for (int sidx = first; sidx < first + count; ++sidx) {
        ExternalGravityParticle src = P[sidx]; /* mass, soft, position: never written by the force routines */
        partBucketForce(&src, req, P, off, 0);
      }
    }
    if (st->piece) st->piece->tp.BucketEwald(req, st->nReps, st->fEwCut);
  }
}

/* accumulator rows {ax, ay, az, pot, dtGrav} of particles [p0, p1) */
void gref_step_vars(const gref_step *st, int p0, int p1, double *vars) {
  for (int j = p0; j < p1; ++j) {
    const GravityParticle &p = st->P[j];
    double *v = vars + (size_t)(j - p0) * 5;
    v[0] = p.treeAcceleration.x; v[1] = p.treeAcceleration.y; v[2] = p.treeAcceleration.z; v[3] = p.potential; v[4] = p.dtGrav;
  }
}

void gref_step_destroy(gref_step *st) {
  if (!st) return;
  delete st->piece;
  delete st;
}

} /* extern "C" */
