/* abi_link_test.cu -- the "link inside ChaNGa" surrogate.  TEST INFRASTRUCTURE.
 *
 * A translation unit written the way ChaNGa's own host code is: it includes the REFERENCE's headers
 * (/root/reference/HostCUDA.h, EwaldCUDA.h, cuda_typedef.h, CudaFunctions.h -- where they lie, nothing
 * copied), fills the reference's structs, and calls the reference's C++-linkage entry points.  It is
 * linked against libchanga_b200.so instead of the reference's HostCUDA.o, so every call below resolves
 * -- by the reference's mangled name -- to the product library.
 *
 *   compile time: every field of every boundary record has the same offset and size in the
 *                 reference's declaration (global namespace) and in include/changa_b200_types.h
 *                 (included a second time inside `namespace ours`);
 *   run time (needs a GPU): one Local cell request, one Local particle request and an EwaldHost
 *                 call through the reference-declared functions,
 *                 checked against a double-precision direct sum computed here.
 *
 * Built by oracle/Makefile into oracle/_ref/abi_link_test (only where /root/reference exists); the
 * binary travels to the GPU box, tests/test_abi.py runs it under -m gpu. */
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "hapi.h"          /* oracle/shim: the three HAPI calls CudaFunctions.h expects declared */
#include "HostCUDA.h"      /* the reference's */
#include "EwaldCUDA.h"     /* the reference's */
#include "CudaFunctions.h" /* the reference's */

namespace ours {
#include "../include/changa_b200_types.h"
}

#define SAME_FIELD(T, f)                                                                              \
  static_assert(offsetof(::T, f) == offsetof(ours::T, f) && sizeof(((::T *)0)->f) == sizeof(((ours::T *)0)->f), \
                "layout of " #T "." #f " differs from the reference")
#define SAME_SIZE(T) static_assert(sizeof(::T) == sizeof(ours::T) && alignof(::T) == alignof(ours::T), "size of " #T)

SAME_SIZE(CudaVector3D); SAME_FIELD(CudaVector3D, x); SAME_FIELD(CudaVector3D, y); SAME_FIELD(CudaVector3D, z);
SAME_SIZE(CudaMultipoleMoments);
SAME_FIELD(CudaMultipoleMoments, radius); SAME_FIELD(CudaMultipoleMoments, soft); SAME_FIELD(CudaMultipoleMoments, totalMass);
SAME_FIELD(CudaMultipoleMoments, cm);
SAME_FIELD(CudaMultipoleMoments, xx); SAME_FIELD(CudaMultipoleMoments, xy); SAME_FIELD(CudaMultipoleMoments, xz);
SAME_FIELD(CudaMultipoleMoments, yy); SAME_FIELD(CudaMultipoleMoments, yz);
SAME_FIELD(CudaMultipoleMoments, xxx); SAME_FIELD(CudaMultipoleMoments, xyy); SAME_FIELD(CudaMultipoleMoments, xxy);
SAME_FIELD(CudaMultipoleMoments, yyy); SAME_FIELD(CudaMultipoleMoments, xxz); SAME_FIELD(CudaMultipoleMoments, yyz);
SAME_FIELD(CudaMultipoleMoments, xyz);
SAME_FIELD(CudaMultipoleMoments, xxxx); SAME_FIELD(CudaMultipoleMoments, xyyy); SAME_FIELD(CudaMultipoleMoments, xxxy);
SAME_FIELD(CudaMultipoleMoments, yyyy); SAME_FIELD(CudaMultipoleMoments, xxxz); SAME_FIELD(CudaMultipoleMoments, yyyz);
SAME_FIELD(CudaMultipoleMoments, xxyy); SAME_FIELD(CudaMultipoleMoments, xxyz); SAME_FIELD(CudaMultipoleMoments, xyyz);
SAME_SIZE(ILPart); SAME_FIELD(ILPart, index); SAME_FIELD(ILPart, off); SAME_FIELD(ILPart, num);
SAME_SIZE(ILCell); SAME_FIELD(ILCell, index); SAME_FIELD(ILCell, offsetID);
SAME_SIZE(CompactPartData); SAME_FIELD(CompactPartData, mass); SAME_FIELD(CompactPartData, soft); SAME_FIELD(CompactPartData, position);
SAME_SIZE(VariablePartData); SAME_FIELD(VariablePartData, a); SAME_FIELD(VariablePartData, potential); SAME_FIELD(VariablePartData, dtGrav);
SAME_SIZE(CudaRequest);
SAME_FIELD(CudaRequest, stream); SAME_FIELD(CudaRequest, d_localMoments); SAME_FIELD(CudaRequest, d_remoteMoments);
SAME_FIELD(CudaRequest, d_localParts); SAME_FIELD(CudaRequest, d_remoteParts); SAME_FIELD(CudaRequest, d_localVars);
SAME_FIELD(CudaRequest, sMoments); SAME_FIELD(CudaRequest, sCompactParts); SAME_FIELD(CudaRequest, sVarParts);
SAME_FIELD(CudaRequest, list); SAME_FIELD(CudaRequest, bucketMarkers); SAME_FIELD(CudaRequest, bucketStarts);
SAME_FIELD(CudaRequest, bucketSizes); SAME_FIELD(CudaRequest, numInteractions); SAME_FIELD(CudaRequest, numBucketsPlusOne);
SAME_FIELD(CudaRequest, tp); SAME_FIELD(CudaRequest, missedNodes); SAME_FIELD(CudaRequest, missedParts);
SAME_FIELD(CudaRequest, sMissed); SAME_FIELD(CudaRequest, affectedBuckets); SAME_FIELD(CudaRequest, cb);
SAME_FIELD(CudaRequest, state); SAME_FIELD(CudaRequest, fperiod); SAME_FIELD(CudaRequest, node); SAME_FIELD(CudaRequest, remote);
SAME_SIZE(CudaDevPtr);
SAME_FIELD(CudaDevPtr, d_list); SAME_FIELD(CudaDevPtr, d_bucketMarkers); SAME_FIELD(CudaDevPtr, d_bucketStarts);
SAME_FIELD(CudaDevPtr, d_bucketSizes);
SAME_SIZE(EwtData);
SAME_FIELD(EwtData, hx); SAME_FIELD(EwtData, hy); SAME_FIELD(EwtData, hz); SAME_FIELD(EwtData, hCfac); SAME_FIELD(EwtData, hSfac);
SAME_SIZE(MultipoleMomentsData);
SAME_FIELD(MultipoleMomentsData, totalMass); SAME_FIELD(MultipoleMomentsData, cmx); SAME_FIELD(MultipoleMomentsData, cmy);
SAME_FIELD(MultipoleMomentsData, cmz);
SAME_SIZE(MomcData);
SAME_FIELD(MomcData, m); SAME_FIELD(MomcData, xx); SAME_FIELD(MomcData, yy); SAME_FIELD(MomcData, xy); SAME_FIELD(MomcData, xz);
SAME_FIELD(MomcData, yz); SAME_FIELD(MomcData, xxx); SAME_FIELD(MomcData, xyy); SAME_FIELD(MomcData, xxy); SAME_FIELD(MomcData, yyy);
SAME_FIELD(MomcData, xxz); SAME_FIELD(MomcData, yyz); SAME_FIELD(MomcData, xyz); SAME_FIELD(MomcData, xxxx); SAME_FIELD(MomcData, xyyy);
SAME_FIELD(MomcData, xxxy); SAME_FIELD(MomcData, yyyy); SAME_FIELD(MomcData, xxxz); SAME_FIELD(MomcData, yyyz); SAME_FIELD(MomcData, xxyy);
SAME_FIELD(MomcData, xxyz); SAME_FIELD(MomcData, xyyz); SAME_FIELD(MomcData, zz); SAME_FIELD(MomcData, xzz); SAME_FIELD(MomcData, yzz);
SAME_FIELD(MomcData, zzz); SAME_FIELD(MomcData, xxzz); SAME_FIELD(MomcData, xyzz); SAME_FIELD(MomcData, xzzz); SAME_FIELD(MomcData, yyzz);
SAME_FIELD(MomcData, yzzz); SAME_FIELD(MomcData, zzzz);
SAME_SIZE(EwaldReadOnlyData);
SAME_FIELD(EwaldReadOnlyData, mm); SAME_FIELD(EwaldReadOnlyData, momcRoot); SAME_FIELD(EwaldReadOnlyData, n);
SAME_FIELD(EwaldReadOnlyData, nReps); SAME_FIELD(EwaldReadOnlyData, nEwReps); SAME_FIELD(EwaldReadOnlyData, nEwhLoop);
SAME_FIELD(EwaldReadOnlyData, L); SAME_FIELD(EwaldReadOnlyData, fEwCut); SAME_FIELD(EwaldReadOnlyData, alpha);
SAME_FIELD(EwaldReadOnlyData, alpha2); SAME_FIELD(EwaldReadOnlyData, k1); SAME_FIELD(EwaldReadOnlyData, ka);
SAME_FIELD(EwaldReadOnlyData, fEwCut2); SAME_FIELD(EwaldReadOnlyData, fInner2);
SAME_SIZE(EwaldData);
SAME_FIELD(EwaldData, EwaldRange); SAME_FIELD(EwaldData, EwaldMarkers); SAME_FIELD(EwaldData, ewt); SAME_FIELD(EwaldData, cachedData);
static_assert(NEWH == 80, "h-table capacity (EwaldCUDA.h:6)");

/* standalone builds of the product library route hapiAddCallback to this handler; the host program
 * registers it through the one symbol that is not the reference's */
extern "C" void cb200_set_callback_handler(void (*handler)(void *));
static int g_fired = 0;
static void on_complete(void *cb) { __sync_fetch_and_add(&g_fired, *(int *)cb); }
/* the shim's hapi.h declares these; the product library carries its own stand-in, nothing here calls them */
void hapiAddCallback(cudaStream_t, void *) {}
void hapiMallocHost(void **p, size_t n, bool) { cudaHostAlloc(p, n, cudaHostAllocDefault); }
void hapiFreeHost(void *p, bool) { cudaFreeHost(p); }

static int offset_code(int x, int y, int z) { return ((x + 3) | ((y + 3) << 3) | ((z + 3) << 6)) << 22; } /* TreePiece.cpp:3631-3644 */

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    printf("abi_link_test: layouts verified at compile time; no CUDA device, run part skipped\n");
    return 0;
  }
  cudaStream_t stream;
  cudaStreamCreate(&stream);
  cb200_set_callback_handler(on_complete);

  /* 3 buckets of 5, 12 and 1 particles; 40 monopole-only cells; 64 further source particles */
  const int sizes[3] = {5, 12, 1}, starts[3] = {0, 5, 17};
  const int nTargets = 18, nParts = nTargets + 64, nCells = 40;
  CompactPartData *parts; VariablePartData *vars, *back; CudaMultipoleMoments *mom;
  allocatePinnedHostMemory((void **)&parts, nParts * sizeof(CompactPartData));
  allocatePinnedHostMemory((void **)&vars, nParts * sizeof(VariablePartData));
  allocatePinnedHostMemory((void **)&back, nParts * sizeof(VariablePartData));
  allocatePinnedHostMemory((void **)&mom, nCells * sizeof(CudaMultipoleMoments));
  srand(12345);
  auto u = []() { return rand() / (double)RAND_MAX - 0.5; };
  for (int i = 0; i < nParts; ++i) {
    parts[i].mass = (float)(1.0 / nParts * (1.0 + 0.5 * u()));
    parts[i].soft = 1e-5f;
    parts[i].position.x = (float)(0.2 * u()); parts[i].position.y = (float)(0.2 * u()); parts[i].position.z = (float)(0.2 * u());
    vars[i].a.x = vars[i].a.y = vars[i].a.z = vars[i].potential = vars[i].dtGrav = 777.0f; /* must be zeroed by the upload */
  }
  memset(mom, 0, nCells * sizeof(CudaMultipoleMoments));
  for (int c = 0; c < nCells; ++c) { /* far away, monopole only: a point mass */
    mom[c].radius = 0.01f; mom[c].soft = 1e-3f; mom[c].totalMass = (float)(0.02 * (1.5 + u()));
    mom[c].cm.x = (float)(u() > 0 ? 0.45 : -0.45); mom[c].cm.y = (float)u(); mom[c].cm.z = (float)u();
  }

  void *d_mom, *d_parts, *d_vars;
  int one = 1;
  DataManagerTransferLocalTree(mom, nCells * sizeof(CudaMultipoleMoments), parts, nParts * sizeof(CompactPartData), vars,
                               nParts * sizeof(VariablePartData), &d_mom, &d_parts, &d_vars, stream, nParts, &one);

  /* lists: every bucket sees all cells (bucket 1 with the +x replica) and all particles */
  std::vector<ILCell> cl, pl;
  int cmarks[4] = {0, 0, 0, 0}, pmarks[4] = {0, 0, 0, 0};
  for (int b = 0; b < 3; ++b) {
    for (int c = 0; c < nCells; ++c) { ILCell e; e.index = c; e.offsetID = offset_code(b == 1 ? 1 : 0, 0, 0) | b; cl.push_back(e); }
    cmarks[b + 1] = (int)cl.size();
    for (int i = 0; i < nParts; ++i) { ILCell e; e.index = i; e.offsetID = offset_code(0, b == 2 ? -1 : 0, 0); pl.push_back(e); }
    pmarks[b + 1] = (int)pl.size();
  }
  auto pinned_copy = [](const void *src, size_t n) { void *p; allocatePinnedHostMemory(&p, n); memcpy(p, src, n); return p; };
  auto request = [&](std::vector<ILCell> &il, int *marks, bool node) {
    CudaRequest r;
    memset(&r, 0, sizeof r);
    r.stream = stream;
    r.d_localMoments = (CudaMultipoleMoments *)d_mom; r.d_localParts = (CompactPartData *)d_parts;
    r.d_localVars = (VariablePartData *)d_vars;
    r.sMoments = nCells * sizeof(CudaMultipoleMoments); r.sCompactParts = nParts * sizeof(CompactPartData);
    r.sVarParts = nParts * sizeof(VariablePartData);
    r.list = pinned_copy(il.data(), il.size() * sizeof(ILCell));
    r.bucketMarkers = (int *)pinned_copy(marks, 4 * sizeof(int));
    r.bucketStarts = (int *)pinned_copy(starts, 3 * sizeof(int));
    r.bucketSizes = (int *)pinned_copy(sizes, 3 * sizeof(int));
    r.numInteractions = (int)il.size(); r.numBucketsPlusOne = 4;
    r.cb = &one; r.fperiod = 1.0f; r.node = node; r.remote = false;
    return r;
  };
  CudaRequest rc = request(cl, cmarks, true), rp = request(pl, pmarks, false);
  TreePieceCellListDataTransferLocal(&rc);
  TreePiecePartListDataTransferLocal(&rp);
  TransferParticleVarsBack(back, nParts * sizeof(VariablePartData), d_vars, stream, &one);
  cudaStreamSynchronize(stream);
  cudaDeviceSynchronize();

  /* double direct sum (all pairs here are outside 2 x softening, or are the self pair) */
  double worst = 0.0;
  for (int b = 0; b < 3; ++b)
    for (int t = starts[b]; t < starts[b] + sizes[b]; ++t) {
      double a[3] = {0, 0, 0}, pot = 0;
      const double px = parts[t].position.x, py = parts[t].position.y, pz = parts[t].position.z;
      for (int c = 0; c < nCells; ++c) {
        const double dx = mom[c].cm.x + (b == 1 ? 1.0 : 0.0) - px, dy = mom[c].cm.y - py, dz = mom[c].cm.z - pz;
        const double r2 = dx * dx + dy * dy + dz * dz, ir = 1.0 / sqrt(r2), m = mom[c].totalMass;
        a[0] += m * dx * ir * ir * ir; a[1] += m * dy * ir * ir * ir; a[2] += m * dz * ir * ir * ir; pot -= m * ir;
      }
      for (int i = 0; i < nParts; ++i) {
        const double dx = parts[i].position.x - px, dy = parts[i].position.y + (b == 2 ? -1.0 : 0.0) - py, dz = parts[i].position.z - pz;
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 == 0) continue;
        const double ir = 1.0 / sqrt(r2), m = parts[i].mass;
        if (sqrt(r2) < 2e-5) continue; /* inside the spline radius: not part of this check */
        a[0] += m * dx * ir * ir * ir; a[1] += m * dy * ir * ir * ir; a[2] += m * dz * ir * ir * ir; pot -= m * ir;
      }
      const double an = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
      const double da = sqrt(pow(back[t].a.x - a[0], 2) + pow(back[t].a.y - a[1], 2) + pow(back[t].a.z - a[2], 2));
      if (da / an > worst) worst = da / an;
      if (fabs(back[t].potential - pot) / fabs(pot) > worst) worst = fabs(back[t].potential - pot) / fabs(pot);
      if (!(back[t].dtGrav > 0)) worst = 1.0;
    }
  bool untouched = true; /* rows outside every bucket were zeroed by the upload and never written */
  for (int i = nTargets; i < nParts; ++i)
    untouched = untouched && back[i].a.x == 0 && back[i].potential == 0 && back[i].dtGrav == 0;

  /* Ewald entry points: memory setup / free and one launch over a 2-particle range with an empty h-table */
  EwaldData ew;
  EwaldHostMemorySetup(&ew, nParts, 1, 0);
  memset(ew.cachedData, 0, sizeof(EwaldReadOnlyData));
  memset(ew.ewt, 0, sizeof(EwtData));
  ew.EwaldRange[0] = 0; ew.EwaldRange[1] = 1;
  ew.cachedData->n = 2; ew.cachedData->nReps = 1; ew.cachedData->nEwReps = 3; ew.cachedData->nEwhLoop = 0;
  ew.cachedData->L = 1.0f; ew.cachedData->fEwCut = 2.6f; ew.cachedData->alpha = 2.0f; ew.cachedData->alpha2 = 4.0f;
  ew.cachedData->k1 = (float)(M_PI / 4.0); ew.cachedData->ka = (float)(4.0 / sqrt(M_PI));
  ew.cachedData->fEwCut2 = 2.6f * 2.6f; ew.cachedData->fInner2 = 1.1e-2f;
  ew.cachedData->mm.totalMass = 1.0f; ew.cachedData->momcRoot.m = 1.0f;
  EwaldHost((CompactPartData *)d_parts, (VariablePartData *)d_vars, &ew, stream, &one, 0, 0);
  cudaStreamSynchronize(stream);
  EwaldHostMemoryFree(&ew, 0);

  cudaFree(d_mom); cudaFree(d_parts); cudaFree(d_vars); /* DataManager.cpp:992-996 frees them with cudaFree */
  cudaDeviceSynchronize();
  const bool ok = worst < 1e-4 && untouched && g_fired == 5 && cudaGetLastError() == cudaSuccess;
  printf("abi_link_test: worst relative error %.3g, untouched rows %s, callbacks fired %d of 5 -> %s\n", worst,
         untouched ? "zero" : "WRITTEN", g_fired, ok ? "ok" : "FAILED");
  return ok ? 0 : 1;
}
