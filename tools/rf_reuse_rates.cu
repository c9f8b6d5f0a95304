// tools/rf_reuse_rates.cu -- does the operand-reuse latch lift the register-file limit of
//   t = coef(.F32 scalar) * mono(pair) + t(pair)      (5 registers, 3 in one bank)?
// Patterns over 6 accumulator pairs t0..t5, scalars s*, monomial pairs m*:
//   A  all operands distinct every instruction
//   B  the monomial is shared by 3 consecutive instructions
//   C  B, and the last coefficient of a triple is the first of the next
//   D  the monomial is shared by 6 consecutive instructions
//   E  B with an unrelated FMUL2 between the triples
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/rf_reuse_rates tools/rf_reuse_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define IT 2048
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
#define F(t, s, m) asm volatile("{.reg .b64 q; mov.b64 q, {%1, %1}; fma.rn.f32x2 %0, q, %2, %0;}" : "+l"(t) : "f"(s), "l"(m))
#define M(d, a, b) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define PRO                                                                                   \
  u64 t0, t1, t2, t3, t4, t5, m0, m1, m2, m3, m4, m5, x = 0;                                  \
  float s0, s1, s2, s3, s4, s5;                                                               \
  { float b = threadIdx.x * 1e-3f;                                                            \
    t0 = pk(b, 1); t1 = pk(b, 2); t2 = pk(b, 3); t3 = pk(b, 4); t4 = pk(b, 5); t5 = pk(b, 6); \
    m0 = pk(1e-3f + b, .1f); m1 = pk(2e-3f + b, .2f); m2 = pk(3e-3f + b, .3f);                \
    m3 = pk(4e-3f + b, .4f); m4 = pk(5e-3f + b, .5f); m5 = pk(6e-3f + b, .6f);                \
    s0 = a0 + b; s1 = a0 * 2 + b; s2 = a0 * 3 + b; s3 = a0 * 4 + b; s4 = a0 * 5 + b; s5 = a0 * 6 + b; }
#define EPI out[blockIdx.x * blockDim.x + threadIdx.x] = t0 ^ t1 ^ t2 ^ t3 ^ t4 ^ t5 ^ x;
__global__ void kA(u64 *out, float a0) { PRO _Pragma("unroll 1") for (int i = 0; i < IT; ++i) {
  F(t0, s0, m0); F(t1, s1, m1); F(t2, s2, m2); F(t3, s3, m3); F(t4, s4, m4); F(t5, s5, m5);
  F(t0, s1, m2); F(t1, s2, m3); F(t2, s3, m4); F(t3, s4, m5); F(t4, s5, m0); F(t5, s0, m1); } EPI }
__global__ void kB(u64 *out, float a0) { PRO _Pragma("unroll 1") for (int i = 0; i < IT; ++i) {
  F(t0, s0, m0); F(t1, s1, m0); F(t2, s2, m0); F(t3, s3, m1); F(t4, s4, m1); F(t5, s5, m1);
  F(t0, s1, m2); F(t1, s2, m2); F(t2, s3, m2); F(t3, s4, m3); F(t4, s5, m3); F(t5, s0, m3); } EPI }
__global__ void kC(u64 *out, float a0) { PRO _Pragma("unroll 1") for (int i = 0; i < IT; ++i) {
  F(t0, s0, m0); F(t1, s1, m0); F(t2, s2, m0); F(t3, s2, m1); F(t4, s4, m1); F(t5, s5, m1);
  F(t0, s5, m2); F(t1, s3, m2); F(t2, s1, m2); F(t3, s1, m3); F(t4, s3, m3); F(t5, s0, m3); } EPI }
__global__ void kD(u64 *out, float a0) { PRO _Pragma("unroll 1") for (int i = 0; i < IT; ++i) {
  F(t0, s0, m0); F(t1, s1, m0); F(t2, s2, m0); F(t3, s3, m0); F(t4, s4, m0); F(t5, s5, m0);
  F(t0, s1, m2); F(t1, s2, m2); F(t2, s3, m2); F(t3, s4, m2); F(t4, s5, m2); F(t5, s0, m2); } EPI }
__global__ void kE(u64 *out, float a0) { PRO _Pragma("unroll 1") for (int i = 0; i < IT; ++i) {
  F(t0, s0, m0); F(t1, s1, m0); F(t2, s2, m0); M(x, m4, m5); F(t3, s3, m1); F(t4, s4, m1); F(t5, s5, m1); M(m4, x, m5);
  F(t0, s1, m2); F(t1, s2, m2); F(t2, s3, m2); M(x, m4, m5); F(t3, s4, m3); F(t4, s5, m3); F(t5, s0, m3); M(m4, x, m5); } EPI }
template <typename K> void run(const char *name, K k, int sms, int clk, u64 *out, int nops) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 1; w <= 4; ++w) {
    if (w == 3) continue;
    k<<<sms, 128 * w>>>(out, 0.25f); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(a); k<<<sms, 128 * w>>>(out, 0.25f); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    double cyc = best * 1e-3 * clk * 1e3 / ((double)IT * nops * w);   // per SMSP: w warps x IT x nops instructions
    printf("%s\"%s_w%d_cycles_per_packed_op\": %.3f", (name[0] == 'A' && w == 1) ? "" : ", ", name, w, cyc);
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  u64 *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 512 * 8);
  printf("{");
  run("A_distinct", kA, p.multiProcessorCount, clk, out, 12);
  run("B_mono3", kB, p.multiProcessorCount, clk, out, 12);
  run("C_mono3_coefchain", kC, p.multiProcessorCount, clk, out, 12);
  run("D_mono6", kD, p.multiProcessorCount, clk, out, 12);
  run("E_mono3_fmul_between", kE, p.multiProcessorCount, clk, out, 16);
  printf("}\n");
  return 0;
}
