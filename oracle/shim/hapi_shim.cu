/* Minimal HAPI behaviour for the reference build under oracle/_ref:
 * callbacks are counted when the stream reaches them; pinned memory is plain
 * cudaHostAlloc.  TEST INFRASTRUCTURE. */
#include "hapi.h"
#include <atomic>
#include <cstdio>
#include <cstdlib>

static std::atomic<long long> g_fired{0};

static void CUDART_CB fire(void *) { g_fired.fetch_add(1); }

void hapiAddCallback(cudaStream_t stream, void *cb) {
  (void)cb;
  cudaLaunchHostFunc(stream, fire, nullptr);
}
void hapiMallocHost(void **ptr, size_t size, bool) {
  if (cudaHostAlloc(ptr, size, cudaHostAllocDefault) != cudaSuccess) {
    fprintf(stderr, "hapi shim: cudaHostAlloc(%zu) failed\n", size);
    abort();
  }
}
void hapiFreeHost(void *ptr, bool) { cudaFreeHost(ptr); }

extern "C" long long refshim_callbacks_fired() { return g_fired.load(); }
extern "C" void *refshim_stream_create() { cudaStream_t s; cudaStreamCreate(&s); return s; }
extern "C" void refshim_stream_sync(void *s) { cudaStreamSynchronize((cudaStream_t)s); }
extern "C" void refshim_free(void *p) { cudaFree(p); }

/* C handles onto the reference's C++-linkage entry points */
#include "HostCUDA.h"
#include "EwaldCUDA.h"
extern "C" {
void ref_allocatePinnedHostMemory(void **p, size_t n) { allocatePinnedHostMemory(p, n); }
void ref_freePinnedHostMemory(void *p) { freePinnedHostMemory(p); }
void ref_DataManagerTransferLocalTree(void *m, size_t sm, void *p, size_t sp, void *v, size_t sv,
                                      void **dm, void **dp, void **dv, void *s, int n, void *cb) {
  DataManagerTransferLocalTree(m, sm, p, sp, v, sv, dm, dp, dv, (cudaStream_t)s, n, cb);
}
void ref_TransferParticleVarsBack(void *h, size_t n, void *d, void *s, void *cb) {
  TransferParticleVarsBack((VariablePartData *)h, n, d, (cudaStream_t)s, cb);
}
void ref_TreePieceCellListDataTransferLocal(CudaRequest *r) { TreePieceCellListDataTransferLocal(r); }
void ref_TreePiecePartListDataTransferLocal(CudaRequest *r) { TreePiecePartListDataTransferLocal(r); }
void ref_EwaldHostMemorySetup(EwaldData *e, int n, int nh, int lp) { EwaldHostMemorySetup(e, n, nh, lp); }
void ref_EwaldHostMemoryFree(EwaldData *e, int lp) { EwaldHostMemoryFree(e, lp); }
void ref_EwaldHost(void *p, void *v, EwaldData *e, void *s, void *cb, int idx, int lp) {
  EwaldHost((CompactPartData *)p, (VariablePartData *)v, e, (cudaStream_t)s, cb, idx, lp);
}
}
