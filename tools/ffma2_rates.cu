// tools/ffma2_rates.cu -- issue rates of the packed-FP32 instruction forms the p-c kernel uses.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ffma2_rates tools/ffma2_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CH 8
#define IT 4096
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
#define KERNEL(name, BODY)                                                     \
  __global__ void name(u64 *out, float s0, float s1, u64 A, u64 B) {          \
    u64 c[CH];                                                                 \
    float sc[CH];                                                              \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) { c[i] = pk(threadIdx.x * 1e-3f + i, i); sc[i] = s0 + i * s1 + threadIdx.x * 1e-6f; } \
    A ^= (u64)(threadIdx.x & 1) << 3; B ^= (u64)(threadIdx.x & 2) << 5; /* per-lane values: regular registers, not uniform ones */ \
    for (int it = 0; it < IT; ++it) {                                          \
      _Pragma("unroll") for (int i = 0; i < CH; ++i) { BODY; }                \
    }                                                                          \
    u64 s = 0;                                                                 \
    _Pragma("unroll") for (int i = 0; i < CH; ++i) s ^= c[i];                  \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;                            \
  }
// packed regs only
KERNEL(k_ffma2_rrr, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(A), "l"(B)))
// scalar broadcast multiplier (per-chain distinct scalar register)
KERNEL(k_ffma2_srr, { u64 b = pk(sc[i], sc[i]); asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(c[i]) : "l"(b), "l"(B)); })
// scalar broadcast multiplier and distinct packed third operand (3 different registers + scalar)
KERNEL(k_ffma2_srr3, { u64 b = pk(sc[i], sc[i]); asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c[i]) : "l"(b), "l"(c[(i + 1) % CH])); })
__global__ void k_ffma2_uniform(u64 *out, float s0, float s1, u64 A, u64 B) {
  u64 c[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i] = pk(threadIdx.x * 1e-3f + i, i);
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(A), "l"(B));
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s ^= c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// scalar 3-register FFMA for reference
__global__ void k_ffma_rrr(u64 *out, float s0, float s1, u64 A, u64 B) {
  float c[CH], d[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i] = threadIdx.x * 1e-3f + i; d[i] = s0 + i * s1 + threadIdx.x * 1e-6f; }
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(c[i]) : "f"(d[i]), "f"(d[(i + 3) % CH]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(s);
}
KERNEL(k_fmul2, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(c[i]) : "l"(A)))
KERNEL(k_fadd2, asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(c[i]) : "l"(A)))
KERNEL(k_fmul2_s, { u64 b = pk(sc[i], sc[i]); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(c[i]) : "l"(b)); })
// dependent chain of length 1 (latency probe): single chain
__global__ void k_lat(u64 *out, u64 A, u64 B) {
  u64 c = pk(threadIdx.x, 1.f);
  A ^= (u64)(threadIdx.x & 1) << 3; B ^= (u64)(threadIdx.x & 2) << 5;
  for (int it = 0; it < IT * CH; ++it) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c) : "l"(A), "l"(B));
  out[blockIdx.x * blockDim.x + threadIdx.x] = c;
}
__global__ void k_lat_scalar(float *out, float a, float b) {
  float c = threadIdx.x;
  for (int it = 0; it < IT * CH; ++it) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(c) : "f"(a), "f"(b));
  out[blockIdx.x * blockDim.x + threadIdx.x] = c;
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
  void *out; cudaMalloc(&out, (size_t)sms * 16 * 1024 * 16);
  u64 A = 0x3f8000103f800010ull, B = 0x3a0000003a000000ull;
  printf("{\"sms\": %d, \"clock_khz\": %d", sms, clk);
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
    int threads = 128 * warps_per_smsp / 1, blocks = sms;  // one CTA per SM, warps spread over the 4 SMSPs
    threads = 32 * 4 * warps_per_smsp;
    double n = (double)blocks * threads * CH * IT;  // packed instructions x lanes
    auto rate = [&](float ms) { return n / (ms * 1e-3) / sms / (clk * 1e3) / 32.0; };  // warp-instr per clk per SM
    float t;
    t = timeit([&] { k_ffma2_rrr<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_ffma2_rrr\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_ffma2_uniform<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_ffma2_uniform_ops\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_ffma_rrr<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_ffma_scalar_rrr\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_ffma2_srr<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_ffma2_srr\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_ffma2_srr3<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_ffma2_srr3\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_fmul2<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_fmul2\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_fmul2_s<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_fmul2_s\": %.3f", warps_per_smsp, rate(t));
    t = timeit([&] { k_fadd2<<<blocks, threads>>>((u64 *)out, 1.f, .1f, A, B); }); printf(", \"w%d_fadd2\": %.3f", warps_per_smsp, rate(t));
  }
  {
    int threads = 32, blocks = sms;
    double n = (double)CH * IT;
    float t = timeit([&] { k_lat<<<blocks, threads>>>((u64 *)out, A, B); });
    printf(", \"ffma2_dep_latency_clk\": %.2f", t * 1e-3 * clk * 1e3 / n);
    t = timeit([&] { k_lat_scalar<<<blocks, threads>>>((float *)out, 1.0001f, 0.5f); });
    printf(", \"ffma_dep_latency_clk\": %.2f", t * 1e-3 * clk * 1e3 / n);
  }
  printf(", \"unit\": \"warp-instructions per clock per SM (4 = one per SMSP per clock)\"}\n");
  return 0;
}
