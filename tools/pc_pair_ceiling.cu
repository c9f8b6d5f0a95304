// tools/pc_pair_ceiling.cu -- compute ceiling of the packed p-c pair evaluation (pc_pair2 of
// changa_b200/csrc/gravity_kernels.cuh) with no list / gather / reduction around it: every lane
// keeps one cell in registers and re-evaluates it against NP target pairs from shared memory.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr \
//        -Iinclude -o gpurun_out/pc_pair_ceiling tools/pc_pair_ceiling.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../changa_b200/csrc/gravity_kernels.cuh"
using namespace cb200;

template <int NP, int MINB>
__global__ void __launch_bounds__(128, MINB) ceiling_kernel(const PackedCell *cells, const float4 *targets, float *out, int iters) {
  __shared__ TargetPair sp[NP];
  if (threadIdx.x < 2 * NP) {
    const float4 q = targets[threadIdx.x];
    float *dst = reinterpret_cast<float *>(sp + (threadIdx.x >> 1)) + (threadIdx.x & 1);
    dst[0] = q.x; dst[2] = q.y; dst[4] = q.z; dst[6] = q.w;
  }
  __syncthreads();
  float c[kCellReals];
  const float *src = reinterpret_cast<const float *>(cells + (blockIdx.x * blockDim.x + threadIdx.x) % 4096);
#pragma unroll
  for (int i = 0; i < kCellReals; ++i) c[i] = src[i];
  f32x2 ax[NP], ay[NP], az[NP], pot[NP];
  float idt[2 * NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) { ax[j] = ay[j] = az[j] = pot[j] = 0ull; idt[2 * j] = idt[2 * j + 1] = 0.f; }
  for (int it = 0; it < iters; ++it) {
    const float ccx = c[PK_CX] + it * 1e-6f, ccy = c[PK_CY], ccz = c[PK_CZ];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const TargetPair p = sp[j];
      pc_pair2(c, ccx, ccy, ccz, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1]);
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    float a, b;
    unpk2(ax[j], a, b); s += a + b; unpk2(ay[j], a, b); s += a + b; unpk2(az[j], a, b); s += a + b;
    unpk2(pot[j], a, b); s += a + b + idt[2 * j] + idt[2 * j + 1];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NP, int MINB>
void run(const PackedCell *cells, const float4 *t, float *out, int sms, int clk) {
  const int iters = 2000, blocks = sms * MINB;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  ceiling_kernel<NP, MINB><<<blocks, 128>>>(cells, t, out, 10);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(a); ceiling_kernel<NP, MINB><<<blocks, 128>>>(cells, t, out, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  const double pairs = (double)blocks * 128 * iters * NP * 2;
  const double cyc_per_eval = best * 1e-3 * clk * 1e3 / ((double)iters * NP * MINB);  // per SMSP: MINB warps each doing iters*NP evals
  printf("{\"NP\": %d, \"warps_per_smsp\": %d, \"pairs_per_s\": %.4g, \"tflops_198\": %.2f, \"cycles_per_packed_eval_per_smsp\": %.1f}\n", NP, MINB,
         pairs / (best * 1e-3), pairs * 198 / (best * 1e-3) * 1e-12, cyc_per_eval);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  PackedCell *cells; float4 *t; float *out;
  cudaMalloc(&cells, 4096 * sizeof(PackedCell)); cudaMalloc(&t, 64 * sizeof(float4)); cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 128 * 4);
  std::vector<float> h(4096 * 32); for (size_t i = 0; i < h.size(); ++i) h[i] = 0.01f + 1e-3f * (i % 97);
  for (int i = 0; i < 4096; ++i) { h[i * 32 + PK_CX] = 0.3f + 1e-4f * i; h[i * 32 + PK_CY] = -0.2f; h[i * 32 + PK_CZ] = 0.1f; h[i * 32 + PK_RADIUS] = 0.02f; }
  cudaMemcpy(cells, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  std::vector<float> ht(64 * 4); for (int i = 0; i < 64; ++i) { ht[4 * i] = 0.001f * i; ht[4 * i + 1] = 0.002f * i; ht[4 * i + 2] = -0.001f * i; ht[4 * i + 3] = 1e-5f; }
  cudaMemcpy(t, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
  run<6, 3>(cells, t, out, p.multiProcessorCount, clk);
  run<6, 2>(cells, t, out, p.multiProcessorCount, clk);
  run<6, 1>(cells, t, out, p.multiProcessorCount, clk);
  run<4, 3>(cells, t, out, p.multiProcessorCount, clk);
  run<4, 4>(cells, t, out, p.multiProcessorCount, clk);
  return 0;
}
