/* Stand-in for Charm++'s hapi.h so that the UNMODIFIED reference HostCUDA.cu
 * compiles outside a Charm++ tree.  Only the three
 * calls HostCUDA.cu makes are provided.  TEST INFRASTRUCTURE. */
#ifndef ORACLE_SHIM_HAPI_H
#define ORACLE_SHIM_HAPI_H
#include <cuda_runtime.h>
#include <cstddef>
void hapiAddCallback(cudaStream_t stream, void *cb);
void hapiMallocHost(void **ptr, size_t size, bool pooled);
void hapiFreeHost(void *ptr, bool pooled);
#endif
