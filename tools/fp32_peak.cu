// tools/fp32_peak.cu -- measures the FP32 FMA-pipe roofline denominators on the box:
// scalar FFMA, packed FFMA2 (fma.rn.f32x2, sm_100+), FP64 DFMA, and MUFU.RSQ issue rates.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fp32_peak tools/fp32_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
#define ITERS 8192
__global__ void k_ffma(float *out, float a, float b) {
  float c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma3(float *out, float a, float b) {  // three distinct register sources
  float c[CHAINS], d[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c[i] = threadIdx.x * 1e-3f + i; d[i] = a + i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fmaf(c[i], d[i], d[(i + 1) % CHAINS]);
  }
  float s = b;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float2 *out, float2 a, float2 b) {
  unsigned long long c[CHAINS];
  unsigned long long A = *reinterpret_cast<unsigned long long *>(&a), B = *reinterpret_cast<unsigned long long *>(&b);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { float2 t = make_float2(threadIdx.x * 1e-3f + i, i); c[i] = *reinterpret_cast<unsigned long long *>(&t); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(A), "l"(B));
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { float2 t = *reinterpret_cast<float2 *>(&c[i]); s.x += t.x; s.y += t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double *out, double a, double b) {
  double c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rsq(float *out, float a) {
  float c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x * 1e-3f + i + a;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(c[i]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount, clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int threads = 256, blocks = sms * 8;
  void *out; cudaMalloc(&out, (size_t)blocks * threads * 16);
  double n = (double)blocks * threads * CHAINS * ITERS;
  float t1 = timeit([&] { k_ffma<<<blocks, threads>>>((float *)out, 1.0001f, 0.5f); });
  float t1b = timeit([&] { k_ffma3<<<blocks, threads>>>((float *)out, 1.0001f, 0.5f); });
  float t2 = timeit([&] { k_ffma2<<<blocks, threads>>>((float2 *)out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); });
  float t3 = timeit([&] { k_dfma<<<blocks, threads>>>((double *)out, 1.0001, 0.5); });
  float t4 = timeit([&] { k_rsq<<<blocks, threads>>>((float *)out, 1.5f); });
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"ffma_tflops\": %.2f, \"ffma_3reg_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"dfma_tflops\": %.2f, \"mufu_rsq_gops\": %.1f, "
         "\"ffma_per_clk_per_sm_at_max_clock\": %.1f, \"ffma2_fma_per_clk_per_sm_at_max_clock\": %.1f}\n",
         p.name, sms, clk, 2 * n / t1 * 1e-9, 2 * n / t1b * 1e-9, 4 * n / t2 * 1e-9, 2 * n / t3 * 1e-9, n / t4 * 1e-6,
         n / (t1 * 1e-3) / sms / (clk * 1e3), 2 * n / (t2 * 1e-3) / sms / (clk * 1e3));
  return 0;
}
