/* gravity_kernels.cuh -- the sm_100a force kernels.
 *
 *   cell_list_kernel   bucket x cell-list, hexadecapole      (replaces nodeGravityComputation,
 *                                                             HostCUDA.cu:1008-1205 + CUDAMoments.cu:14-109)
 *   part_list_kernel   bucket x particle-list, spline p-p    (replaces particleGravityComputation,
 *                                                             HostCUDA.cu:1565-1751)
 *   ewald_kernel       periodic correction per particle      (replaces EwaldKernel, HostCUDA.cu:1958-2192)
 *   repack_* / zero    layout conversion at upload time      (replaces ZeroVars, HostCUDA.cu:2195-2205)
 *
 * Work decomposition (both list kernels): ONE WARP OWNS ONE BUCKET.  The 32
 * lanes spread over the bucket's interaction list (lane = list entry), each
 * lane keeps its cell / source particle in registers and walks the bucket's
 * <= PB target particles, which sit in shared memory and are read as warp
 * broadcasts.  Per-target partial sums stay in registers (5 x PB per lane)
 * and are combined once per bucket with an xor-butterfly, so every lane is
 * busy whatever the bucket size is (the reference's 16x8 thread tile idles
 * (16 - bucketSize)/16 of its lanes), there is no __syncthreads anywhere, and
 * the summation order is fixed -> bitwise reproducible results.  Warps pull
 * buckets from a global counter (persistent CTAs, dynamic load balance over
 * ragged lists).
 *
 * Data movement: the list is streamed from HBM with coalesced 8-byte loads one
 * tile ahead; the 128-byte PackedCell rows it points at are gathered with
 * 16-byte cp.async (LDGSTS) into a double-buffered, XOR-swizzled shared tile
 * (8 lanes cover one row = one L2 line, 4 rows per instruction), overlapping
 * the gather of tile t+1 with the arithmetic of tile t.
 */
#ifndef CB200_GRAVITY_KERNELS_CUH
#define CB200_GRAVITY_KERNELS_CUH

#include <cuda_runtime.h>
#include <math_constants.h>
#include "device_layout.cuh"

namespace cb200 {

constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) real4 { real x, y, z, w; };

/* ------------------------------------------------------------------ helpers */
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
/* same, allocating in L1 (records that neighbouring warps gather again) */
__device__ __forceinline__ void cp_async16_ca(void *smem, const void *gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float rsqrt_dev(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double rsqrt_dev(double x) { return rsqrt(x); }

__device__ __forceinline__ void unpack_piece(const uint4 &v, float *o) {
  o[0] = __uint_as_float(v.x); o[1] = __uint_as_float(v.y);
  o[2] = __uint_as_float(v.z); o[3] = __uint_as_float(v.w);
}
__device__ __forceinline__ void unpack_piece(const uint4 &v, double *o) {
  o[0] = __hiloint2double(v.y, v.x);
  o[1] = __hiloint2double(v.w, v.z);
}

__device__ __forceinline__ float rmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double rmax(double a, double b) { return fmax(a, b); }

/* -------------------------------------------------------- layout conversion */
/* CudaMultipoleMoments (27 reals, cuda_typedef.h:104-128) -> PackedCell.
 * The 3 / 15 / 105 factors are the (2l-1)!! of g2, g3, g4
 * (CUDAMoments.cu:40-44) folded into the components once per upload. */
__global__ void repack_cells_kernel(const real *__restrict__ raw, PackedCell *__restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const real *m = raw + (size_t)i * 27;
  real radius = m[0], soft = m[1], mass = m[2], cx = m[3], cy = m[4], cz = m[5];
  real xx = 3 * m[6], xy = 3 * m[7], xz = 3 * m[8], yy = 3 * m[9], yz = 3 * m[10];
  real xxx = 15 * m[11], xyy = 15 * m[12], xxy = 15 * m[13], yyy = 15 * m[14], xxz = 15 * m[15],
       yyz = 15 * m[16], xyz = 15 * m[17];
  real xxxx = 105 * m[18], xyyy = 105 * m[19], xxxy = 105 * m[20], yyyy = 105 * m[21],
       xxxz = 105 * m[22], yyyz = 105 * m[23], xxyy = 105 * m[24], xxyz = 105 * m[25],
       xyyz = 105 * m[26];
  PackedCell c;
  c.v[PK_CX] = cx; c.v[PK_CY] = cy; c.v[PK_CZ] = cz; c.v[PK_RADIUS] = radius;
  c.v[PK_MASS] = mass; c.v[PK_XX] = xx; c.v[PK_XY] = xy; c.v[PK_XZ] = xz;
  c.v[PK_YY] = yy; c.v[PK_YZ] = yz; c.v[PK_ZZ] = -(xx + yy); c.v[PK_XXX] = xxx;
  c.v[PK_XYY] = xyy; c.v[PK_XXY] = xxy; c.v[PK_YYY] = yyy; c.v[PK_XXZ] = xxz;
  c.v[PK_YYZ] = yyz; c.v[PK_XYZ] = xyz; c.v[PK_XZZ] = -(xxx + xyy); c.v[PK_YZZ] = -(xxy + yyy);
  c.v[PK_XXXX] = xxxx; c.v[PK_XYYY] = xyyy; c.v[PK_XXXY] = xxxy; c.v[PK_YYYY] = yyyy;
  c.v[PK_XXXZ] = xxxz; c.v[PK_YYYZ] = yyyz; c.v[PK_XXYY] = xxyy; c.v[PK_XXYZ] = xxyz;
  c.v[PK_XYYZ] = xyyz; c.v[PK_XY3S] = xyyy + xxxy; c.v[PK_SOFT] = soft; c.v[PK_PAD] = 0;
  uint4 *dst = reinterpret_cast<uint4 *>(out + i);
  const uint4 *src = reinterpret_cast<const uint4 *>(&c);
#pragma unroll
  for (int j = 0; j < kCellPieces; ++j) dst[j] = src[j];
}

/* CompactPartData {mass, soft, x, y, z} (cuda_typedef.h:240-243) -> PackedPart */
__global__ void repack_parts_kernel(const real *__restrict__ raw, PackedPart *__restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const real *p = raw + (size_t)i * 5;
  PackedPart q;
  q.x = p[2]; q.y = p[3]; q.z = p[4]; q.mass = p[0];
  q.soft = p[1]; q.pad0 = q.pad1 = q.pad2 = 0;
  uint4 *dst = reinterpret_cast<uint4 *>(out + i);
  const uint4 *src = reinterpret_cast<const uint4 *>(&q);
#pragma unroll
  for (int j = 0; j < kPartBytes / 16; ++j) dst[j] = src[j];
}

/* ------------------------------------------------- shared tile of PackedCells */
/* piece j of the row in slot s lives at s*kCellPieces + (j ^ (s & 7)): both the
 * row-wise cp.async writes (consecutive lanes = consecutive pieces of one row)
 * and the column-wise register loads (lane s reads piece j of its own row) are
 * bank-conflict free. */
__device__ __forceinline__ void stage_cell_tile(uint4 *buf, const PackedCell *__restrict__ cells,
                                                int myIndex, int lane) {
  constexpr int kRowsPerInst = 32 / kCellPieces;
  const int q = lane % kCellPieces, sub = lane / kCellPieces;
#pragma unroll
  for (int i = 0; i < kCellPieces; ++i) {
    const int slot = i * kRowsPerInst + sub;
    const int idx = __shfl_sync(kFull, myIndex, slot);
    if (idx >= 0)
      cp_async16(&buf[slot * kCellPieces + (q ^ (slot & 7))],
                 reinterpret_cast<const uint4 *>(cells + idx) + q);
  }
}

__device__ __forceinline__ void load_cell_row(const uint4 *buf, int lane, real *c) {
#pragma unroll
  for (int j = 0; j < kCellPieces; ++j) unpack_piece(buf[lane * kCellPieces + (j ^ (lane & 7))], c + j * kPieceReals);
}

/* ------------------------------------------------------------ p-c evaluation */
/* One (target particle, cell) pair: potential and acceleration of the cell's
 * multipole expansion to hexadecapole order.  Same series as momEvalFmomrcm
 * (moments.c:1469-1525) / CUDA_momEvalFmomrcm (CUDAMoments.cu:14-109), written
 * in the scaled displacement xi = r * radius / |r|^2 so that every g_l factor
 * collapses into a power of xi and the pre-multiplied components:
 *   T_l   = (l-1)-fold contraction of the order-l moment with xi   (vectors)
 *   S_l   = T_l . xi
 *   pot  -= d  * (M + S2/2 + S3/3 + S4/4)
 *   acc  += d^3 * (radius*(T2+T3+T4) - r*(M + 5/2 S2 + 7/3 S3 + 9/4 S4))
 *   idt2  = max(idt2, (m_p + M) d^3)                      (gravity.h:446)
 * with d = 1/|r| (0 when r == 0: the pair is skipped, HostCUDA.cu:1103). */
struct PairOut { real ax, ay, az, pot, idt; };

__device__ __forceinline__ void pc_pair(const real *__restrict__ c, real ccx, real ccy, real ccz,
                                        const real4 &p, real &ax, real &ay, real &az, real &pot,
                                        real &idt) {
  const real third = real(1.0 / 3.0);
  const real rx = p.x - ccx, ry = p.y - ccy, rz = p.z - ccz;
  const real rsq = fma(rz, rz, fma(ry, ry, rx * rx));
  real d = rsqrt_dev(rsq);
  d = (rsq != real(0)) ? d : real(0);
  const real d2 = d * d;
  const real s = c[PK_RADIUS] * d2;
  const real X = rx * s, Y = ry * s, Z = rz * s;

  real xx = (real(0.5) * X) * X, yy = (real(0.5) * Y) * Y;
  const real zz = (real(0.5) * Z) * Z;
  const real xy = X * Y, xz = X * Z, yz = Y * Z;
  const real xxx = X * fma(third, xx, -zz);
  const real xxz = Z * fma(-third, zz, xx);
  const real yyy = Y * fma(third, yy, -zz);
  const real yyz = Z * fma(-third, zz, yy);
  xx -= zz;
  yy -= zz;
  const real xxy = Y * xx, xyy = X * yy, xyz = xy * Z;

  /* hexadecapole */
  real t4x = c[PK_XXXX] * xxx;
  t4x = fma(c[PK_XYYY], yyy, t4x); t4x = fma(c[PK_XXXY], xxy, t4x); t4x = fma(c[PK_XXXZ], xxz, t4x);
  t4x = fma(c[PK_XXYY], xyy, t4x); t4x = fma(c[PK_XXYZ], xyz, t4x); t4x = fma(c[PK_XYYZ], yyz, t4x);
  real t4y = c[PK_XYYY] * xyy;
  t4y = fma(c[PK_XXXY], xxx, t4y); t4y = fma(c[PK_YYYY], yyy, t4y); t4y = fma(c[PK_YYYZ], yyz, t4y);
  t4y = fma(c[PK_XXYY], xxy, t4y); t4y = fma(c[PK_XXYZ], xxz, t4y); t4y = fma(c[PK_XYYZ], xyz, t4y);
  real t4z = c[PK_XXXZ] * xxx;
  t4z = fma(c[PK_YYYZ], yyy, t4z); t4z = fma(c[PK_XXYZ], xxy, t4z); t4z = fma(c[PK_XYYZ], xyy, t4z);
  t4z = fma(-c[PK_XXXX], xxz, t4z); t4z = fma(-c[PK_XY3S], xyz, t4z); t4z = fma(-c[PK_YYYY], yyz, t4z);
  t4z = fma(-c[PK_XXYY], xxz + yyz, t4z);
  const real s4 = fma(t4z, Z, fma(t4y, Y, t4x * X));

  /* octupole */
  real t3x = c[PK_XXX] * xx;
  t3x = fma(c[PK_XYY], yy, t3x); t3x = fma(c[PK_XXY], xy, t3x); t3x = fma(c[PK_XXZ], xz, t3x);
  t3x = fma(c[PK_XYZ], yz, t3x);
  real t3y = c[PK_XYY] * xy;
  t3y = fma(c[PK_XXY], xx, t3y); t3y = fma(c[PK_YYY], yy, t3y); t3y = fma(c[PK_YYZ], yz, t3y);
  t3y = fma(c[PK_XYZ], xz, t3y);
  real t3z = c[PK_XZZ] * xz;
  t3z = fma(c[PK_YZZ], yz, t3z); t3z = fma(c[PK_XXZ], xx, t3z); t3z = fma(c[PK_YYZ], yy, t3z);
  t3z = fma(c[PK_XYZ], xy, t3z);
  const real s3 = fma(t3z, Z, fma(t3y, Y, t3x * X));

  /* quadrupole */
  const real t2x = fma(c[PK_XZ], Z, fma(c[PK_XY], Y, c[PK_XX] * X));
  const real t2y = fma(c[PK_YZ], Z, fma(c[PK_XY], X, c[PK_YY] * Y));
  const real t2z = fma(c[PK_YZ], Y, fma(c[PK_XZ], X, c[PK_ZZ] * Z));
  const real s2 = fma(t2z, Z, fma(t2y, Y, t2x * X));

  const real M = c[PK_MASS];
  const real phi = fma(real(0.25), s4, fma(third, s3, fma(real(0.5), s2, M)));
  const real G = fma(real(2.25), s4, fma(real(7.0 / 3.0), s3, fma(real(2.5), s2, M)));
  const real d3 = d2 * d;
  const real R = c[PK_RADIUS];
  pot = fma(-d, phi, pot);
  ax = fma(d3, fma(-rx, G, R * (t2x + t3x + t4x)), ax);
  ay = fma(d3, fma(-ry, G, R * (t2y + t3y + t4y)), ay);
  az = fma(d3, fma(-rz, G, R * (t2z + t3z + t4z)), az);
  idt = rmax(idt, (p.w + M) * d3);
}

/* The same evaluation for N targets at once, statement by statement: the N dependency chains are
 * interleaved in program order, which is what a latency-bound schedule needs (pc_pair called N
 * times in a row is scheduled mostly one after the other).  Identical arithmetic to pc_pair. */
template <int N>
__device__ __forceinline__ void pc_pairN(const real *__restrict__ c, real ccx, real ccy, real ccz,
                                         const real4 *p, real *ax, real *ay, real *az, real *pot, real *idt) {
  const real third = real(1.0 / 3.0);
  real rx[N], ry[N], rz[N], d[N], d2[N], X[N], Y[N], Z[N];
#pragma unroll
  for (int t = 0; t < N; ++t) {
    rx[t] = p[t].x - ccx; ry[t] = p[t].y - ccy; rz[t] = p[t].z - ccz;
  }
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const real rsq = fma(rz[t], rz[t], fma(ry[t], ry[t], rx[t] * rx[t]));
    const real r = rsqrt_dev(rsq);
    d[t] = (rsq != real(0)) ? r : real(0);
  }
#pragma unroll
  for (int t = 0; t < N; ++t) {
    d2[t] = d[t] * d[t];
    const real s = c[PK_RADIUS] * d2[t];
    X[t] = rx[t] * s; Y[t] = ry[t] * s; Z[t] = rz[t] * s;
  }
  real xx[N], yy[N], zz[N], xy[N], xz[N], yz[N], xxx[N], xxz[N], yyy[N], yyz[N], xxy[N], xyy[N], xyz[N];
#pragma unroll
  for (int t = 0; t < N; ++t) {
    xx[t] = (real(0.5) * X[t]) * X[t]; yy[t] = (real(0.5) * Y[t]) * Y[t]; zz[t] = (real(0.5) * Z[t]) * Z[t];
    xy[t] = X[t] * Y[t]; xz[t] = X[t] * Z[t]; yz[t] = Y[t] * Z[t];
  }
#pragma unroll
  for (int t = 0; t < N; ++t) {
    xxx[t] = X[t] * fma(third, xx[t], -zz[t]);
    xxz[t] = Z[t] * fma(-third, zz[t], xx[t]);
    yyy[t] = Y[t] * fma(third, yy[t], -zz[t]);
    yyz[t] = Z[t] * fma(-third, zz[t], yy[t]);
  }
#pragma unroll
  for (int t = 0; t < N; ++t) {
    xx[t] -= zz[t]; yy[t] -= zz[t];
    xxy[t] = Y[t] * xx[t]; xyy[t] = X[t] * yy[t]; xyz[t] = xy[t] * Z[t];
  }
  real t4x[N], t4y[N], t4z[N], t3x[N], t3y[N], t3z[N];
#define CB200_ALL(stmt) _Pragma("unroll") for (int t = 0; t < N; ++t) { stmt; }
  /* hexadecapole */
  CB200_ALL(t4x[t] = c[PK_XXXX] * xxx[t])
  CB200_ALL(t4y[t] = c[PK_XYYY] * xyy[t])
  CB200_ALL(t4z[t] = c[PK_XXXZ] * xxx[t])
  CB200_ALL(t4x[t] = fma(c[PK_XYYY], yyy[t], t4x[t]))
  CB200_ALL(t4y[t] = fma(c[PK_XXXY], xxx[t], t4y[t]))
  CB200_ALL(t4z[t] = fma(c[PK_YYYZ], yyy[t], t4z[t]))
  CB200_ALL(t4x[t] = fma(c[PK_XXXY], xxy[t], t4x[t]))
  CB200_ALL(t4y[t] = fma(c[PK_YYYY], yyy[t], t4y[t]))
  CB200_ALL(t4z[t] = fma(c[PK_XXYZ], xxy[t], t4z[t]))
  CB200_ALL(t4x[t] = fma(c[PK_XXXZ], xxz[t], t4x[t]))
  CB200_ALL(t4y[t] = fma(c[PK_YYYZ], yyz[t], t4y[t]))
  CB200_ALL(t4z[t] = fma(c[PK_XYYZ], xyy[t], t4z[t]))
  CB200_ALL(t4x[t] = fma(c[PK_XXYY], xyy[t], t4x[t]))
  CB200_ALL(t4y[t] = fma(c[PK_XXYY], xxy[t], t4y[t]))
  CB200_ALL(t4z[t] = fma(-c[PK_XXXX], xxz[t], t4z[t]))
  CB200_ALL(t4x[t] = fma(c[PK_XXYZ], xyz[t], t4x[t]))
  CB200_ALL(t4y[t] = fma(c[PK_XXYZ], xxz[t], t4y[t]))
  CB200_ALL(t4z[t] = fma(-c[PK_XY3S], xyz[t], t4z[t]))
  CB200_ALL(t4x[t] = fma(c[PK_XYYZ], yyz[t], t4x[t]))
  CB200_ALL(t4y[t] = fma(c[PK_XYYZ], xyz[t], t4y[t]))
  CB200_ALL(t4z[t] = fma(-c[PK_YYYY], yyz[t], t4z[t]))
  CB200_ALL(t4z[t] = fma(-c[PK_XXYY], xxz[t] + yyz[t], t4z[t]))
  /* octupole */
  CB200_ALL(t3x[t] = c[PK_XXX] * xx[t])
  CB200_ALL(t3y[t] = c[PK_XYY] * xy[t])
  CB200_ALL(t3z[t] = c[PK_XZZ] * xz[t])
  CB200_ALL(t3x[t] = fma(c[PK_XYY], yy[t], t3x[t]))
  CB200_ALL(t3y[t] = fma(c[PK_XXY], xx[t], t3y[t]))
  CB200_ALL(t3z[t] = fma(c[PK_YZZ], yz[t], t3z[t]))
  CB200_ALL(t3x[t] = fma(c[PK_XXY], xy[t], t3x[t]))
  CB200_ALL(t3y[t] = fma(c[PK_YYY], yy[t], t3y[t]))
  CB200_ALL(t3z[t] = fma(c[PK_XXZ], xx[t], t3z[t]))
  CB200_ALL(t3x[t] = fma(c[PK_XXZ], xz[t], t3x[t]))
  CB200_ALL(t3y[t] = fma(c[PK_YYZ], yz[t], t3y[t]))
  CB200_ALL(t3z[t] = fma(c[PK_YYZ], yy[t], t3z[t]))
  CB200_ALL(t3x[t] = fma(c[PK_XYZ], yz[t], t3x[t]))
  CB200_ALL(t3y[t] = fma(c[PK_XYZ], xz[t], t3y[t]))
  CB200_ALL(t3z[t] = fma(c[PK_XYZ], xy[t], t3z[t]))
#undef CB200_ALL
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const real s4 = fma(t4z[t], Z[t], fma(t4y[t], Y[t], t4x[t] * X[t]));
    const real s3 = fma(t3z[t], Z[t], fma(t3y[t], Y[t], t3x[t] * X[t]));
    /* quadrupole */
    const real t2x = fma(c[PK_XZ], Z[t], fma(c[PK_XY], Y[t], c[PK_XX] * X[t]));
    const real t2y = fma(c[PK_YZ], Z[t], fma(c[PK_XY], X[t], c[PK_YY] * Y[t]));
    const real t2z = fma(c[PK_YZ], Y[t], fma(c[PK_XZ], X[t], c[PK_ZZ] * Z[t]));
    const real s2 = fma(t2z, Z[t], fma(t2y, Y[t], t2x * X[t]));
    const real M = c[PK_MASS];
    const real phi = fma(real(0.25), s4, fma(third, s3, fma(real(0.5), s2, M)));
    const real G = fma(real(2.25), s4, fma(real(7.0 / 3.0), s3, fma(real(2.5), s2, M)));
    const real d3 = d2[t] * d[t];
    const real R = c[PK_RADIUS];
    pot[t] = fma(-d[t], phi, pot[t]);
    ax[t] = fma(d3, fma(-rx[t], G, R * (t2x + t3x[t] + t4x[t])), ax[t]);
    ay[t] = fma(d3, fma(-ry[t], G, R * (t2y + t3y[t] + t4y[t])), ay[t]);
    az[t] = fma(d3, fma(-rz[t], G, R * (t2z + t3z[t] + t4z[t])), az[t]);
    idt[t] = rmax(idt[t], (p[t].w + M) * d3);
  }
}

/* xor-butterfly over the warp; every lane ends with the same totals */
__device__ __forceinline__ void warp_reduce5(real &a0, real &a1, real &a2, real &a3, real &a4) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(kFull, a0, o);
    a1 += __shfl_xor_sync(kFull, a1, o);
    a2 += __shfl_xor_sync(kFull, a2, o);
    a3 += __shfl_xor_sync(kFull, a3, o);
    a4 = rmax(a4, __shfl_xor_sync(kFull, a4, o));
  }
}

constexpr int kListWarps = 4; /* warps (= buckets in flight) per CTA */

/* Next bucket from the grid-wide counter.  Every warp stops at its first index >= nBuckets, so a
 * launch performs exactly nBuckets + (warps of the grid) increments: the warp that draws the last
 * value puts the counter back to zero, and the next launch on the stream needs no memset node. */
__device__ __forceinline__ int grab_bucket(unsigned int *nextBucket, int nBuckets, int lane) {
  unsigned int k = 0;
  if (lane == 0) {
    k = atomicAdd(nextBucket, 1u);
    if (k == (unsigned int)nBuckets + gridDim.x * kListWarps - 1u) *nextBucket = 0u;
  }
  return (int)__shfl_sync(0xffffffffu, k, 0);
}

template <int PB>
constexpr size_t cell_list_smem_bytes() {
  return (size_t)kListWarps * (2 * 32 * kCellBytes + PB * sizeof(real4));
}

/* ---------------------------------------------------------- particle-cell */
/* PAIR: two targets per bounds check, so that their (independent) evaluations sit in one basic block and
 * ptxas can interleave them -- the FP64 build is latency-bound at 2 resident warps per scheduler
 * (DESIGN 4.4).  With an odd particle count the second evaluation of the last pair meets a stale target;
 * its sums are never reduced.  Experimental: selected by CB200_PC64_VARIANT, not the default. */
template <int PB, int MINB, bool PAIR = false>
__global__ void __launch_bounds__(kListWarps * 32, MINB)
cell_list_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                 const PackedCell *__restrict__ cells, const ILCell *__restrict__ list,
                 const int *__restrict__ markers, const int *__restrict__ starts,
                 const int *__restrict__ sizes, int nBuckets, real fperiod,
                 unsigned int *__restrict__ nextBucket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4 *tiles = reinterpret_cast<uint4 *>(smem_raw) + (size_t)warp * (2 * 32 * kCellPieces);
  real4 *sp = reinterpret_cast<real4 *>(smem_raw + (size_t)kListWarps * 2 * 32 * kCellBytes) + warp * PB;

  for (;;) {
    const int k = grab_bucket(nextBucket, nBuckets, lane);
    if (k >= nBuckets) break;
    const int begin = markers[k], len = markers[k + 1] - begin;
    const int first = starts[k], count = sizes[k];
    if (len <= 0) continue;
    const ILCell *__restrict__ mylist = list + begin;
    const int ntiles = (len + 31) >> 5;

    for (int p0 = 0; p0 < count; p0 += PB) { /* one pass unless the bucket outgrows PB */
      const int np = min(PB, count - p0);
      __syncwarp();
      if (lane < np) sp[lane] = *reinterpret_cast<const real4 *>(parts + first + p0 + lane);
      __syncwarp();

      real ax[PB], ay[PB], az[PB], pot[PB], idt[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) ax[j] = ay[j] = az[j] = pot[j] = idt[j] = real(0);

      /* software pipeline: list entries two tiles ahead, cell rows one tile ahead */
      ILCell cur, nxt;
      cur.index = -1; cur.offsetID = 0; nxt = cur;
      if (lane < len) cur = mylist[lane];
      stage_cell_tile(tiles, cells, cur.index, lane);
      cp_async_commit();
      if (32 + lane < len) nxt = mylist[32 + lane];

      for (int t = 0; t < ntiles; ++t) {
        uint4 *buf = tiles + (t & 1) * (32 * kCellPieces);
        if (t + 1 < ntiles) stage_cell_tile(tiles + ((t + 1) & 1) * (32 * kCellPieces), cells, nxt.index, lane);
        cp_async_commit();
        ILCell nn;
        nn.index = -1; nn.offsetID = 0;
        if ((t + 2) * 32 + lane < len) nn = mylist[(t + 2) * 32 + lane];
        cp_async_wait<1>();
        __syncwarp();

        if (cur.index >= 0) {
          real c[kCellReals];
          load_cell_row(buf, lane, c);
          const real ccx = fma(real(replica_x(cur.offsetID)), fperiod, c[PK_CX]);
          const real ccy = fma(real(replica_y(cur.offsetID)), fperiod, c[PK_CY]);
          const real ccz = fma(real(replica_z(cur.offsetID)), fperiod, c[PK_CZ]);
          if constexpr (PAIR) {
            static_assert(!PAIR || PB % 2 == 0, "pairs of targets");
#pragma unroll
            for (int j = 0; j < PB; j += 2) {
              if (j < np) {
                const real4 pq[2] = {sp[j], sp[j + 1]};
                pc_pairN<2>(c, ccx, ccy, ccz, pq, ax + j, ay + j, az + j, pot + j, idt + j);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < PB; ++j) {
              if (j < np) {
                const real4 p = sp[j];
                pc_pair(c, ccx, ccy, ccz, p, ax[j], ay[j], az[j], pot[j], idt[j]);
              }
            }
          }
        }
        __syncwarp();
        cur = nxt;
        nxt = nn;
      }
      cp_async_wait<0>();

      real m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        if (j < np) {
          real a0 = ax[j], a1 = ay[j], a2 = az[j], a3 = pot[j], a4 = idt[j];
          warp_reduce5(a0, a1, a2, a3, a4);
          if (lane == j) { m0 = a0; m1 = a1; m2 = a2; m3 = a3; m4 = a4; }
        }
      }
      if (lane < np) { /* accumulate, never overwrite (HostCUDA.cu:1196-1200) */
        VariablePartData *v = vars + first + p0 + lane;
        v->a.x += m0; v->a.y += m1; v->a.z += m2;
        v->potential += m3;
        v->dtGrav = rmax(v->dtGrav, m4);
      }
    }
  }
}

#ifndef CUDA_USE_DOUBLE
/* ------------------------------------------- particle-cell, packed f32x2 math */
/* Blackwell's FMA pipe issues a 3-register FFMA every other cycle per SM
 * sub-partition but a packed FFMA2 (fma.rn.f32x2: two lanes of a 64-bit
 * register pair) at the same cadence -- twice the flops per issue slot, and the
 * only way to the FP32 peak with register operands (tools/fp32_peak.cu:
 * 3-register FFMA 49 TFLOP/s, FFMA2 74 TFLOP/s).  FFMA2 also takes a 32-bit
 * operand broadcast to both halves (SASS `Rn.F32`), so a lane keeps ONE cell's
 * coefficients as scalars and evaluates it against TWO target particles per
 * instruction.  Targets sit in shared memory as pairs {x0,x1 | y0,y1 | z0,z1 |
 * m0,m1}; accumulators are packed pairs too.  Same series as pc_pair above. */
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 bc2(float c) { return pk2(c, c); } /* ptxas folds this into an .F32 operand */
__device__ __forceinline__ void unpk2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
/* sign flips fold into the consuming instruction's operand modifier */
__device__ __forceinline__ f32x2 neg2(f32x2 a) {
  float lo, hi;
  unpk2(a, lo, hi);
  return pk2(-lo, -hi);
}
/* a*s + c and a*s with a scalar s */
__device__ __forceinline__ f32x2 fma2s(float s, f32x2 a, f32x2 c) { return fma2(bc2(s), a, c); }
__device__ __forceinline__ f32x2 mul2s(float s, f32x2 a) { return mul2(bc2(s), a); }

struct TargetPair { f32x2 x, y, z, m; }; /* 32 bytes: two LDS.128 */

__device__ __forceinline__ void pc_pair2(const float *__restrict__ c, float ccx, float ccy, float ccz,
                                         const TargetPair &p, f32x2 &ax, f32x2 &ay, f32x2 &az, f32x2 &pot,
                                         float &idt0, float &idt1) {
  const float third = 1.0f / 3.0f, sixth = 1.0f / 6.0f;
  const f32x2 rx = sub2(p.x, bc2(ccx)), ry = sub2(p.y, bc2(ccy)), rz = sub2(p.z, bc2(ccz));
  const f32x2 rsq = fma2(rz, rz, fma2(ry, ry, mul2(rx, rx)));
  float q0, q1;
  unpk2(rsq, q0, q1);
  float d0 = rsqrt_dev(q0), d1 = rsqrt_dev(q1);
  d0 = (q0 != 0.0f) ? d0 : 0.0f; /* r == 0: the pair is skipped (HostCUDA.cu:1103) */
  d1 = (q1 != 0.0f) ? d1 : 0.0f;
  const f32x2 d = pk2(d0, d1);
  const f32x2 d2 = mul2(d, d);
  const f32x2 s = mul2s(c[PK_RADIUS], d2);
  const f32x2 X = mul2(rx, s), Y = mul2(ry, s), Z = mul2(rz, s);

  /* monomials xi^alpha / alpha!, with the z-traces already removed (20 ops):
   *   xxm = (X^2 - Z^2)/2            xxx = X (X^2/6 - Z^2/2)
   *   xxz = Z (X^2/2 - Z^2/6) = Z (xxm + Z^2/3)        (same with y) */
  const f32x2 a = mul2(X, X), b = mul2(Y, Y), cz = mul2(Z, Z);
  const f32x2 nhc = mul2s(-0.5f, cz);
  const f32x2 xxm = fma2s(0.5f, a, nhc), yym = fma2s(0.5f, b, nhc);
  const f32x2 xxx = mul2(X, fma2s(sixth, a, nhc)), yyy = mul2(Y, fma2s(sixth, b, nhc));
  const f32x2 xxz = mul2(Z, fma2s(third, cz, xxm)), yyz = mul2(Z, fma2s(third, cz, yym));
  const f32x2 xy = mul2(X, Y), xz = mul2(X, Z), yz = mul2(Y, Z);
  const f32x2 xxy = mul2(Y, xxm), xyy = mul2(X, yym), xyz = mul2(xy, Z);

  /* T accumulates the contracted vectors order by order (4, then 4+3, then
   * 4+3+2); A4, A34, A234 are its projections on xi at each stage, so that
   *   s4 = A4, s3 = A34 - A4, s2 = A234 - A34.
   *
   * Instruction order is chosen for the register file, not for reading: a
   * packed FMA  t = coef * mono + t  reads five registers (one scalar, two
   * pairs) and the file delivers two per bank (even/odd) in the two cycles the
   * FMA pipe needs, so it issues at the pipe rate only when one operand comes
   * from the operand-reuse latch of the instruction before it
   * (tools/sass_rf_model.py, tools/ffma2_rates.cu: 2 vs 3 cycles).  The terms
   * are therefore grouped by monomial (3 consecutive uses, one per component),
   * the three components rotate in a fixed order (no accumulator is touched
   * twice within three instructions), and consecutive monomials are chosen so
   * that the z-coefficient of one is the x-coefficient of the next. */
  /* hexadecapole: 7 + 7 + 8 terms */
  f32x2 tx = mul2s(c[PK_XXXX], xxx), ty = mul2s(c[PK_XXXY], xxx), tz = mul2s(c[PK_XXXZ], xxx);
  tx = fma2s(c[PK_XXXZ], xxz, tx); ty = fma2s(c[PK_XXYZ], xxz, ty); tz = fma2s(-c[PK_XXXX], xxz, tz);
  tx = fma2s(c[PK_XXXY], xxy, tx); ty = fma2s(c[PK_XXYY], xxy, ty); tz = fma2s(c[PK_XXYZ], xxy, tz);
  tx = fma2s(c[PK_XXYZ], xyz, tx); ty = fma2s(c[PK_XYYZ], xyz, ty); tz = fma2s(-c[PK_XY3S], xyz, tz);
  tx = fma2s(c[PK_XXYY], xyy, tx); ty = fma2s(c[PK_XYYY], xyy, ty); tz = fma2s(c[PK_XYYZ], xyy, tz);
  tx = fma2s(c[PK_XYYZ], yyz, tx); ty = fma2s(c[PK_YYYZ], yyz, ty); tz = fma2s(-c[PK_YYYY], yyz, tz);
  tx = fma2s(c[PK_XYYY], yyy, tx); ty = fma2s(c[PK_YYYY], yyy, ty); tz = fma2s(c[PK_YYYZ], yyy, tz);
  tz = fma2s(-c[PK_XXYY], add2(xxz, yyz), tz);
  const f32x2 t4x = tx, t4y = ty, t4z = tz;

  /* octupole: 5 + 5 + 5 terms */
  tx = fma2s(c[PK_XXX], xxm, tx); ty = fma2s(c[PK_XXY], xxm, ty); tz = fma2s(c[PK_XXZ], xxm, tz);
  tx = fma2s(c[PK_XXZ], xz, tx); ty = fma2s(c[PK_XYZ], xz, ty); tz = fma2s(c[PK_XZZ], xz, tz);
  tx = fma2s(c[PK_XXY], xy, tx); ty = fma2s(c[PK_XYY], xy, ty); tz = fma2s(c[PK_XYZ], xy, tz);
  tx = fma2s(c[PK_XYZ], yz, tx); ty = fma2s(c[PK_YYZ], yz, ty); tz = fma2s(c[PK_YZZ], yz, tz);
  tx = fma2s(c[PK_XYY], yym, tx); ty = fma2s(c[PK_YYY], yym, ty); tz = fma2s(c[PK_YYZ], yym, tz);
  const f32x2 t3x = tx, t3y = ty, t3z = tz;

  /* quadrupole: 3 + 3 + 3 terms */
  tx = fma2s(c[PK_XX], X, tx); ty = fma2s(c[PK_XY], X, ty); tz = fma2s(c[PK_XZ], X, tz);
  tx = fma2s(c[PK_XZ], Z, tx); ty = fma2s(c[PK_YZ], Z, ty); tz = fma2s(c[PK_ZZ], Z, tz);
  tx = fma2s(c[PK_XY], Y, tx); ty = fma2s(c[PK_YY], Y, ty); tz = fma2s(c[PK_YZ], Y, tz);
  /* the three projections side by side: X, then Y, then Z stays latched in its slot */
  f32x2 A4 = mul2(t4x, X), A34 = mul2(t3x, X), A234 = mul2(tx, X);
  A4 = fma2(t4y, Y, A4); A34 = fma2(t3y, Y, A34); A234 = fma2(ty, Y, A234);
  A4 = fma2(t4z, Z, A4); A34 = fma2(t3z, Z, A34); A234 = fma2(tz, Z, A234);

  /* phi = M + s2/2 + s3/3 + s4/4;  G = M + 5/2 s2 + 7/3 s3 + 9/4 s4 = phi + 2 (s2+s3+s4) */
  const float M = c[PK_MASS];
  const f32x2 phi = fma2s(-1.0f / 12.0f, A4, fma2s(-sixth, A34, fma2s(0.5f, A234, bc2(M))));
  const f32x2 G = fma2s(2.0f, A234, phi);
  const f32x2 d3 = mul2(d2, d);
  pot = fma2(neg2(d), phi, pot);
  const f32x2 e = mul2s(c[PK_RADIUS], d3), g = neg2(mul2(G, d3));
  ax = fma2(e, tx, ax); ay = fma2(e, ty, ay); az = fma2(e, tz, az);
  ax = fma2(g, rx, ax); ay = fma2(g, ry, ay); az = fma2(g, rz, az);
  float i0, i1;
  unpk2(mul2(add2(p.m, bc2(M)), d3), i0, i1);
  idt0 = fmaxf(idt0, i0);
  idt1 = fmaxf(idt1, i1);
}

/* Shared memory of cell_list_x2_kernel: first the cell tiles of all warps (each warp's
 * pair of 4 KB tiles is 1 KB-aligned: rows must be 128-byte aligned, an LDGSTS whose 128-byte
 * row straddles two shared-memory lines is split into extra L2 requests -- measured 3x), then
 * the small per-warp areas.  Inside a tile piece p of row s lives at position p ^ (s & 7), so the
 * row-wise cp.async writes and the lane-wise 16-byte reads are both free of bank conflicts; with
 * the tile 1 KB-aligned both swizzles are ONE xor of a lane-constant address with an immediate. */
constexpr int kTileBytes = 32 * kCellBytes;       /* 4096 */
constexpr int kRedPitch = 33 * 4;                 /* reduction scratch rows: 32 floats + 1 pad */
template <int PB>
struct CellWarpSmem {
  static constexpr int tilesAll = kListWarps * 2 * kTileBytes;
  static constexpr int targets = 0;                            /* PB/2 TargetPair                              */
  static constexpr int preList = targets + (PB / 2) * 32;      /* first 64 list entries of the NEXT bucket     */
  static constexpr int preTargets = preList + 64 * 8;          /* raw {x,y,z,m} of the next bucket's targets   */
  static constexpr int bytes = preTargets + PB * 16;
};
template <int PB>
constexpr size_t cell_list_x2_smem_bytes() {
  return (size_t)CellWarpSmem<PB>::tilesAll + (size_t)kListWarps * CellWarpSmem<PB>::bytes;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
/* a value ptxas must keep in a register instead of recomputing it at every use */
__device__ __forceinline__ unsigned pin_u32(unsigned v) { unsigned o; asm volatile("mov.u32 %0, %1;" : "=r"(o) : "r"(v)); return o; }
__device__ __forceinline__ void cp_async16_s(unsigned saddr, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8_s(unsigned saddr, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gmem) : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ TargetPair lds_target_pair(unsigned saddr) {
  TargetPair p;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(p.x), "=l"(p.y) : "r"(saddr));
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(p.z), "=l"(p.m) : "r"(saddr));
  return p;
}

/* Work decomposition as cell_list_kernel (one warp owns one bucket, lane =
 * list entry, double-buffered cp.async tile of PackedCell rows), with
 *   - the inner loop over target PAIRS in packed f32x2 math (pc_pair2);
 *   - everything around the pair evaluations stripped to `register + immediate`
 *     addressing: per tile 8 x (SHFL with an immediate source lane, one IMAD.WIDE,
 *     LDGSTS) to gather the 32 rows, 8 LDS.128 to read the lane's own row;
 *   - the next bucket's first 64 list entries and its targets prefetched with
 *     cp.async into a small staging area under the last tile of the current bucket
 *     (no registers held across the bucket);
 *   - the per-bucket reduction through shared memory: every lane parks its 5*PB
 *     partial sums as column `lane` of a [value][32] array (row pitch 144 B) laid
 *     over the idle cell tiles, then lane v adds up row v with 8 LDS.128.  Rows are
 *     particle-major, so row v IS float v of the bucket's contiguous
 *     VariablePartData block and the += is one coalesced read-modify-write.
 *     Fixed summation order -> bitwise reproducible. */
struct BucketMeta { int begin, len, first, count; };

__device__ __forceinline__ BucketMeta load_bucket_meta(const int *__restrict__ markers,
                                                       const int *__restrict__ starts,
                                                       const int *__restrict__ sizes, int k) {
  BucketMeta m;
  m.begin = markers[k];
  m.len = markers[k + 1] - m.begin;
  m.first = starts[k];
  m.count = sizes[k];
  return m;
}

/* ---- the bucket pipeline of the packed list kernels -------------------------------------------
 * A warp works on bucket k while it already holds k1 (the next one: its first list entries and targets
 * are prefetched under k's last tile) and, on launches with many buckets per warp, while the atomic
 * that draws k2 is in flight.  ncu of the 256^3 step (profiles/r02j_ncu_part_list_stream_256.json):
 * 13.6 % of the p-p kernel's stall samples sat on the shuffle that broadcasts the atomic's result and
 * on the first use of the bucket's markers -- both at the top of a bucket, with nothing to overlap. */
__device__ __forceinline__ unsigned grab_bucket_raw(unsigned int *nextBucket, int nBuckets, int lane) {
  unsigned int k = 0;
  /* inline PTX, and nothing looks at k here: nvcc turns a plain atomicAdd under `lane == 0` into its
   * warp-aggregated form, which broadcasts the result with a SHFL right behind the atomic -- the round trip this
   * pipeline exists to hide (6.2 % of the p-p kernel's stall samples, profiles/r02aa_ncu_pp_256.json).  The counter
   * is put back to zero by the launch's last warp to leave (list_kernel_exit) instead of by whoever draws the last
   * value (grab_bucket): that test needs the value at once. */
  if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(k) : "l"(nextBucket) : "memory");
  (void)nBuckets;
  return k;
}
/* every warp of a launch calls this once, behind its last draw: nextBucket[1] counts the warps that are done, the
 * last one zeroes both words for the next launch on the stream */
__device__ __forceinline__ void list_kernel_exit(unsigned int *nextBucket, int lane) {
  if (lane == 0) {
    const unsigned int d = atomicAdd(nextBucket + 1, 1u);
    if (d == gridDim.x * kListWarps - 1u) { nextBucket[0] = 0u; nextBucket[1] = 0u; }
  }
}
/* markers / start / size of a bucket, loaded through asm so that ptxas does not treat the values as
 * warp-uniform: a uniform value is moved to the uniform register file at once (R2UR right behind the
 * load = a full memory round trip at the top of every bucket); these stay in flight until the bucket
 * is started, when bucket_meta_uniform broadcasts them */
struct RawMeta { int begin, end, first, count; };
__device__ __forceinline__ RawMeta load_bucket_meta_raw(const int *__restrict__ markers, const int *__restrict__ starts,
                                                        const int *__restrict__ sizes, int k) {
  RawMeta r;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r.begin) : "l"(markers + k));
  asm volatile("ld.global.nc.s32 %0, [%1+4];" : "=r"(r.end) : "l"(markers + k));
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r.first) : "l"(starts + k));
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(r.count) : "l"(sizes + k));
  return r;
}
__device__ __forceinline__ BucketMeta bucket_meta_uniform(const RawMeta &r) {
  BucketMeta m;
  m.begin = __shfl_sync(kFull, r.begin, 0);
  m.len = __shfl_sync(kFull, r.end, 0) - m.begin;
  m.first = __shfl_sync(kFull, r.first, 0);
  m.count = __shfl_sync(kFull, r.count, 0);
  return m;
}
/* launches with fewer buckets per warp than this keep one bucket in reserve instead of two: what a warp
 * holds and has not started is idle at the tail of the launch (13 k buckets over 1776 warps on cube300) */
constexpr int kDeepPipeBuckets = 32;

/* ---- per-bucket reduction of the packed list kernels ---------------------------------------------
 * Every lane holds, per target PAIR j, the packed partial sums {ax, ay, az, pot} (two targets per
 * register pair) and two dtGrav maxima.  Sums: the lane parks its 4 * npairs packed values as 8-byte
 * column `lane` of a [4 * npairs][32] array (row pitch 272 B: STS.64 and LDS.128 both conflict-free),
 * lane v then adds up row v with 16 LDS.128 and packed adds -- one pass for a bucket of up to 16
 * particles -- and adds both halves to the bucket's VariablePartData rows with RED.ADD.F32 (no load of
 * the old value, nothing to wait for; one warp owns a bucket, so the order of additions to an address
 * is the launch order on the stream: bitwise reproducible).  dtGrav: a non-negative float orders like
 * its bit pattern, so the maximum over the lanes is one REDUX.MAX.U32 per target and the store one
 * RED.MAX.U32.  Replaces parking 5 * np scalar rows and two divergent 32-element row sums (sum rows
 * and max rows interleaved): 290 -> ~130 instructions per bucket, 18 % of the p-p kernel's stall
 * samples at 256^3 (profiles/r02j_ncu_part_list_stream_256.json). */
constexpr int kRed2Pitch = 272;
__device__ __forceinline__ void sts64(unsigned a, f32x2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ f32x2 lds64(unsigned a) { f32x2 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void red_add_f32(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_max_u32(float *p, unsigned v) { asm volatile("red.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int NP>
__device__ __forceinline__ void bucket_reduce_store(const f32x2 (&ax)[NP], const f32x2 (&ay)[NP], const f32x2 (&az)[NP],
                                                    const f32x2 (&pot)[NP], const float (&idt)[2 * NP], int np, int npairs,
                                                    unsigned redBase, int lane, float *__restrict__ out) {
  const unsigned col = redBase + lane * 8;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (j < npairs) {
      const unsigned r0 = col + (4 * j) * kRed2Pitch;
      sts64(r0, ax[j]); sts64(r0 + kRed2Pitch, ay[j]); sts64(r0 + 2 * kRed2Pitch, az[j]); sts64(r0 + 3 * kRed2Pitch, pot[j]);
    }
  }
  /* dtGrav while the stores land */
  unsigned myMax = 0u;
#pragma unroll
  for (int i = 0; i < 2 * NP; ++i) { /* unguarded: slots past np hold zeros (or an odd bucket's unused half) and are not stored */
    const unsigned mx = __reduce_max_sync(kFull, __float_as_uint(idt[i]));
    if (lane == i) myMax = mx;
  }
  __syncwarp();
  if (lane < 4 * npairs) {
    const unsigned row = redBase + lane * kRed2Pitch;
    f32x2 acc0 = 0ull, acc1 = 0ull;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      f32x2 e0, e1;
      asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(e0), "=l"(e1) : "r"(row + i * 16));
      acc0 = add2(acc0, e0);
      acc1 = add2(acc1, e1);
    }
    float s0, s1;
    unpk2(add2(acc0, acc1), s0, s1);
    /* accumulate, never overwrite (HostCUDA.cu:1196-1200, 1749-1751) */
    const int t0 = 2 * (lane >> 2), comp = lane & 3;
    red_add_f32(out + t0 * 5 + comp, s0);
    if (t0 + 1 < np) red_add_f32(out + (t0 + 1) * 5 + comp, s1);
  }
  if (lane < np) red_max_u32(out + lane * 5 + 4, myMax); /* dtGrav is a running max */
  __syncwarp();
}

template <int PB, int MINB>
__global__ void __launch_bounds__(kListWarps * 32, MINB)
cell_list_x2_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                    const PackedCell *__restrict__ cells, const ILCell *__restrict__ list,
                    const int *__restrict__ markers, const int *__restrict__ starts,
                    const int *__restrict__ sizes, int nBuckets, float fperiod,
                    unsigned int *__restrict__ nextBucket) {
  static_assert(PB % 2 == 0 && 2 * PB <= 32, "one packed sum row per lane (bucket_reduce_store)");
  static_assert(4 * (PB / 2) * kRed2Pitch <= 2 * kTileBytes, "reduction scratch fits in the cell tiles");
  static_assert(kCellPieces == 8, "float build");
  constexpr int NP = PB / 2;
  typedef CellWarpSmem<PB> S;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned tbase = smem_u32(smem_raw) + warp * (2 * kTileBytes);
  if (tbase & 1023u) __trap(); /* the xor swizzles below need it */
  unsigned char *wsm = smem_raw + S::tilesAll + warp * S::bytes;
  const unsigned wbase = smem_u32(wsm);
  /* lane-constant addresses */
  const unsigned rowAddr = pin_u32(tbase + lane * kCellBytes + (lane & 7) * 16);          /* my row of tile 0: piece j at ^ (j << 4) */
  const unsigned dstAddr = pin_u32(tbase + (lane & ~7) * kCellBytes + (lane & 7) * 16);   /* row 8g + i, piece q at ^ (i * 144)      */
  const unsigned tgtAddr = pin_u32(wbase + S::targets);
  const char *srcBase; /* piece q of row 0, pinned: one IMAD.WIDE per gathered row */
  asm volatile("mov.u64 %0, %1;" : "=l"(srcBase) : "l"(reinterpret_cast<const char *>(cells) + (lane & 7) * 16));

  /* rows 8g .. 8g+7 of a tile are fetched by lane group g (8 lanes = 8 pieces of one row) */
  auto stage = [&](unsigned dst, int index) {
    /* lanes without an entry re-fetch the tile's first row (never read): no predicate per row,
     * and no single row of the array that every warp of the grid would hammer */
#ifdef CB200_V_PRED
    const int idx = index;
#else
    const int first = __shfl_sync(kFull, index, 0);
    const int idx = index >= 0 ? index : first;
#endif
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = __shfl_sync(kFull, idx, i, 8);
#ifdef CB200_V_PRED
      if (r >= 0)
#endif
      cp_async16_s(dst ^ (i * (kCellBytes + 16)), srcBase + (size_t)(unsigned)r * kCellBytes);
    }
  };
  /* first two list tiles and the targets of bucket b -> staging area */
  auto prefetch_bucket = [&](const RawMeta &b) { /* per-lane values: nothing here needs them uniform */
    const ILCell *nl = list + b.begin;
    const int blen = b.end - b.begin;
    if (lane < blen) cp_async8_s(wbase + S::preList + lane * 8, nl + lane);
    if (32 + lane < blen) cp_async8_s(wbase + S::preList + 256 + lane * 8, nl + 32 + lane);
    if (lane < min(PB, b.count)) cp_async16_s(wbase + S::preTargets + lane * 16, parts + b.first + lane);
  };
  const ILCell none = {-1, 0};

  /* the bucket pipeline (see grab_bucket_raw): k in work, k1 held, k2's atomic in flight */
  const bool deep = nBuckets >= kDeepPipeBuckets * (int)(gridDim.x * kListWarps);
  int k = __shfl_sync(kFull, grab_bucket_raw(nextBucket, nBuckets, lane), 0);
  int k1 = k;
  if (k < nBuckets) k1 = __shfl_sync(kFull, grab_bucket_raw(nextBucket, nBuckets, lane), 0);
  BucketMeta m = {0, 0, 0, 0};
  if (k < nBuckets) {
    const RawMeta r = load_bucket_meta_raw(markers, starts, sizes, k);
    prefetch_bucket(r);
    m = bucket_meta_uniform(r);
  }
  cp_async_commit();

  while (k < nBuckets) {
    /* a warp draws exactly one index >= nBuckets (the self-resetting counter counts on it) */
    unsigned k2raw = (unsigned)k1;
    if (deep && k1 < nBuckets) k2raw = grab_bucket_raw(nextBucket, nBuckets, lane);
    RawMeta mn = {0, 0, 0, 0};
    if (k1 < nBuckets) mn = load_bucket_meta_raw(markers, starts, sizes, k1);
    bool prefetched = false;

    const ILCell *__restrict__ mylist = list + m.begin;
    const int len = m.len, ntiles = (len + 31) >> 5;
    for (int p0 = 0; p0 < m.count && len > 0; p0 += PB) { /* one pass unless the bucket outgrows PB */
      const int np = min(PB, m.count - p0);
      const int npairs = (np + 1) >> 1;
      const bool lastPass = p0 + PB >= m.count;
      ILCell cur = none, nxt = none;
      cp_async_wait<0>();
      __syncwarp();
      if (p0 == 0) { /* from the staging area */
        if (lane < len) cur = *reinterpret_cast<const ILCell *>(wsm + S::preList + lane * 8);
        if (32 + lane < len) nxt = *reinterpret_cast<const ILCell *>(wsm + S::preList + 256 + lane * 8);
      } else {
        if (lane < len) cur = mylist[lane];
        if (32 + lane < len) nxt = mylist[32 + lane];
      }
      if (lane < 2 * npairs) { /* an odd bucket's last slot repeats its last particle; that half is never stored */
        const int src = min(lane, np - 1);
        float4 q;
        if (p0 == 0) q = *reinterpret_cast<const float4 *>(wsm + S::preTargets + src * 16);
        else q = *reinterpret_cast<const float4 *>(parts + m.first + p0 + src);
        float *dst = reinterpret_cast<float *>(wsm + S::targets) + (lane >> 1) * 8 + (lane & 1);
        dst[0] = q.x; dst[2] = q.y; dst[4] = q.z; dst[6] = q.w;
      }
      __syncwarp();

      f32x2 ax[NP], ay[NP], az[NP], pot[NP];
      float idt[PB];
#pragma unroll
      for (int j = 0; j < NP; ++j) { ax[j] = ay[j] = az[j] = pot[j] = 0ull; idt[2 * j] = idt[2 * j + 1] = 0.0f; }

      stage(dstAddr, cur.index);
      cp_async_commit();

      for (int t = 0; t < ntiles; ++t) {
        const unsigned boff = (t & 1) * kTileBytes; /* a multiple of 4 KB: does not disturb the xor swizzles */
        if (t + 1 < ntiles) stage(dstAddr + (kTileBytes - boff), nxt.index);
        if (lastPass && t == ntiles - 1) { /* the next bucket's first loads ride under this tile */
          prefetch_bucket(mn);
          prefetched = true;
        }
        cp_async_commit();
        ILCell nn = none;
#ifdef CB200_V_LDG
        if ((t + 2) * 32 + lane < len) { const int2 e = __ldg(reinterpret_cast<const int2 *>(mylist + (t + 2) * 32 + lane)); nn.index = e.x; nn.offsetID = e.y; }
#else
        if ((t + 2) * 32 + lane < len) nn = mylist[(t + 2) * 32 + lane];
#endif
        cp_async_wait<1>();
        __syncwarp();

        if (cur.index >= 0) {
          float c[kCellReals];
          const unsigned row = rowAddr + boff;
#pragma unroll
          for (int j = 0; j < 8; ++j) unpack_piece(lds128(row ^ (j * 16)), c + j * 4);
          const float ccx = fmaf(float(replica_x(cur.offsetID)), fperiod, c[PK_CX]);
          const float ccy = fmaf(float(replica_y(cur.offsetID)), fperiod, c[PK_CY]);
          const float ccz = fmaf(float(replica_z(cur.offsetID)), fperiod, c[PK_CZ]);
#ifdef CB200_V_DOUBLE
          /* two pair evaluations per basic block: the dependent head of one (displacement ->
           * rsqrt -> scaled displacement) overlaps the FMA-dense tail of the other */
#pragma unroll
          for (int j = 0; j < NP; j += 2) {
            if (j + 1 < npairs) {
              const TargetPair pa = lds_target_pair(tgtAddr + j * 32);
              const TargetPair pb = lds_target_pair(tgtAddr + (j + 1) * 32);
              pc_pair2(c, ccx, ccy, ccz, pa, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1]);
              pc_pair2(c, ccx, ccy, ccz, pb, ax[j + 1], ay[j + 1], az[j + 1], pot[j + 1], idt[2 * j + 2], idt[2 * j + 3]);
            } else if (j < npairs) {
              const TargetPair p = lds_target_pair(tgtAddr + j * 32);
              pc_pair2(c, ccx, ccy, ccz, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1]);
            }
          }
#else
          /* (requesting the next pair's targets before this pair is evaluated -- the first FADD2 of a body
           * waits on its own LDS, 9 % of the stall samples -- was measured: 27.61 ms against 27.19 at 256^3,
           * the 12 extra registers cost more than the wait; profiles/r02k) */
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            if (j < npairs) {
              const TargetPair p = lds_target_pair(tgtAddr + j * 32);
              pc_pair2(c, ccx, ccy, ccz, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1]);
            }
          }
#endif
        }
        __syncwarp();
        cur = nxt;
        nxt = nn;
      }
      /* every tile has landed (wait<1> of the last iteration); only the staging-area
       * prefetch may still be in flight, and it does not touch the tiles */

      bucket_reduce_store<NP>(ax, ay, az, pot, idt, np, npairs, tbase, lane,
                              reinterpret_cast<float *>(vars + m.first + p0));
    }
    if (!prefetched) { /* empty list: nothing rode under a tile, and nothing waited for this bucket's own prefetch */
      cp_async_wait<0>(); /* two copies in flight to the same staging bytes may land in either order (racecheck, r02ag) */
      prefetch_bucket(mn);
    }
    cp_async_commit();
    if (!deep && k1 < nBuckets) k2raw = grab_bucket_raw(nextBucket, nBuckets, lane);
    k = k1;
    k1 = (int)__shfl_sync(kFull, k2raw, 0);
    m = bucket_meta_uniform(mn);
  }
  cp_async_wait<0>();
  list_kernel_exit(nextBucket, lane);
}
#endif /* !CUDA_USE_DOUBLE */

/* ------------------------------------------------------ particle-particle */
/* Hernquist-Katz spline-softened monopole, SPLINE of gravity.h:147-182.
 * r = (shift + source) - target (HostCUDA.cu:1655-1663). */
__device__ __forceinline__ void pp_pair(real sx, real sy, real sz, real sm, real ssoft,
                                        const real4 &p, real psoft, real &ax, real &ay, real &az,
                                        real &pot, real &idt) {
  const real rx = sx - p.x, ry = sy - p.y, rz = sz - p.z;
  const real rsq = fma(rz, rz, fma(ry, ry, rx * rx));
  const real twoh = ssoft + psoft;
  real a, b;
  real d = rsqrt_dev(rsq);
  d = (rsq != real(0)) ? d : real(0);
  if (rsq >= twoh * twoh) {
    a = d;
    b = d * d * d;
  } else if (rsq == real(0)) { /* self / coincident: skipped (HostCUDA.cu:1665) */
    a = b = real(0);
  } else {
    const real r = rsq * d;
    const real dih = real(2) / twoh;
    const real u = r * dih, u2 = u * u;
    const real dih3 = dih * dih * dih;
    if (u < real(1)) {
      a = dih * (real(7.0 / 5.0) + u2 * (real(-2.0 / 3.0) + u2 * (real(3.0 / 10.0) - real(1.0 / 10.0) * u)));
      b = dih3 * (real(4.0 / 3.0) + u2 * (real(-6.0 / 5.0) + real(0.5) * u));
    } else {
      a = real(-1.0 / 15.0) * d +
          dih * (real(8.0 / 5.0) + u2 * (real(-4.0 / 3.0) + u * (real(1) + u * (real(-3.0 / 10.0) + real(1.0 / 30.0) * u))));
      b = real(-1.0 / 15.0) * d * d * d +
          dih3 * (real(8.0 / 3.0) + u * (real(-3) + u * (real(6.0 / 5.0) - real(1.0 / 6.0) * u)));
    }
  }
  const real bm = b * sm;
  ax = fma(rx, bm, ax);
  ay = fma(ry, bm, ay);
  az = fma(rz, bm, az);
  pot = fma(-sm, a, pot);
  idt = rmax(idt, (p.w + sm) * b);
}

template <int PB>
constexpr size_t part_list_smem_bytes() {
  return (size_t)kListWarps * PB * (sizeof(real4) + sizeof(real));
}

template <int PB, int MINB>
__global__ void __launch_bounds__(kListWarps * 32, MINB)
part_list_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                 const PackedPart *__restrict__ sources, const ILCell *__restrict__ list,
                 const int *__restrict__ markers, const int *__restrict__ starts,
                 const int *__restrict__ sizes, int nBuckets, real fperiod,
                 unsigned int *__restrict__ nextBucket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  real4 *sp = reinterpret_cast<real4 *>(smem_raw) + warp * PB;
  real *ssoft = reinterpret_cast<real *>(smem_raw + (size_t)kListWarps * PB * sizeof(real4)) + warp * PB;

  for (;;) {
    const int k = grab_bucket(nextBucket, nBuckets, lane);
    if (k >= nBuckets) break;
    const int begin = markers[k], len = markers[k + 1] - begin;
    const int first = starts[k], count = sizes[k];
    if (len <= 0) continue;
    const ILCell *__restrict__ mylist = list + begin;
    const int ntiles = (len + 31) >> 5;

    for (int p0 = 0; p0 < count; p0 += PB) {
      const int np = min(PB, count - p0);
      __syncwarp();
      if (lane < np) {
        const PackedPart *q = parts + first + p0 + lane;
        sp[lane] = *reinterpret_cast<const real4 *>(q);
        ssoft[lane] = q->soft;
      }
      __syncwarp();

      real ax[PB], ay[PB], az[PB], pot[PB], idt[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) ax[j] = ay[j] = az[j] = pot[j] = idt[j] = real(0);

      /* source rows one tile ahead in registers (8 reals), list entries two ahead */
      ILCell cur, nxt;
      cur.index = -1; cur.offsetID = 0; nxt = cur;
      if (lane < len) cur = mylist[lane];
      if (32 + lane < len) nxt = mylist[32 + lane];
      real4 s_pos = {0, 0, 0, 0};
      real s_soft = 0;
      if (cur.index >= 0) {
        const PackedPart *q = sources + cur.index;
        s_pos = *reinterpret_cast<const real4 *>(q);
        s_soft = q->soft;
      }

      for (int t = 0; t < ntiles; ++t) {
        real4 n_pos = {0, 0, 0, 0};
        real n_soft = 0;
        if (nxt.index >= 0) {
          const PackedPart *q = sources + nxt.index;
          n_pos = *reinterpret_cast<const real4 *>(q);
          n_soft = q->soft;
        }
        ILCell nn;
        nn.index = -1; nn.offsetID = 0;
        if ((t + 2) * 32 + lane < len) nn = mylist[(t + 2) * 32 + lane];

        if (cur.index >= 0) {
          const real sx = fma(real(replica_x(cur.offsetID)), fperiod, s_pos.x);
          const real sy = fma(real(replica_y(cur.offsetID)), fperiod, s_pos.y);
          const real sz = fma(real(replica_z(cur.offsetID)), fperiod, s_pos.z);
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            if (j < np) {
              const real4 p = sp[j];
              pp_pair(sx, sy, sz, s_pos.w, s_soft, p, ssoft[j], ax[j], ay[j], az[j], pot[j], idt[j]);
            }
          }
        }
        cur = nxt; nxt = nn;
        s_pos = n_pos; s_soft = n_soft;
      }

      real m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        if (j < np) {
          real a0 = ax[j], a1 = ay[j], a2 = az[j], a3 = pot[j], a4 = idt[j];
          warp_reduce5(a0, a1, a2, a3, a4);
          if (lane == j) { m0 = a0; m1 = a1; m2 = a2; m3 = a3; m4 = a4; }
        }
      }
      if (lane < np) {
        VariablePartData *v = vars + first + p0 + lane;
        v->a.x += m0; v->a.y += m1; v->a.z += m2;
        v->potential += m3;
        v->dtGrav = rmax(v->dtGrav, m4);
      }
    }
  }
}

#ifndef CUDA_USE_DOUBLE
/* ---------------------------------------- particle-particle, packed f32x2 math */
/* Same idea as cell_list_x2_kernel: the lane keeps ONE source particle and meets
 * TWO targets per instruction.  The unsoftened branch (r >= soft_s + soft_t for
 * both targets, by far the common case) is 17 packed FP instructions + 2 MUFU.RSQ
 * per two pairs; a pair that needs the spline falls back to the scalar pp_pair
 * for both halves (gravity.h:147-182). */
struct TargetSoftPair { f32x2 x, y, z, m, soft; float pad[2]; }; /* 48 bytes */
__device__ __forceinline__ TargetSoftPair lds_target_soft_pair(unsigned saddr) {
  TargetSoftPair p;
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(p.x), "=l"(p.y) : "r"(saddr));
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(p.z), "=l"(p.m) : "r"(saddr));
  asm volatile("ld.shared.u64 %0, [%1+32];" : "=l"(p.soft) : "r"(saddr));
  p.pad[0] = p.pad[1] = 0.0f;
  return p;
}

/* Branch-free Newtonian part of two pairs.  A half that needs the spline (0 < r < soft_s + soft_t)
 * or is exactly coincident contributes nothing here (d = 0); the return value says whether some half
 * still needs the spline, which pp_pair2_soft adds afterwards -- rare, so the bodies of a tile carry
 * no divergent region and the compiler can overlap them. */
__device__ __forceinline__ bool pp_pair2(float sx, float sy, float sz, float sm, float ssoft,
                                         const TargetSoftPair &p, f32x2 &ax, f32x2 &ay, f32x2 &az,
                                         f32x2 &pot, float &idt0, float &idt1) {
  const f32x2 rx = sub2(bc2(sx), p.x), ry = sub2(bc2(sy), p.y), rz = sub2(bc2(sz), p.z);
  const f32x2 rsq = fma2(rz, rz, fma2(ry, ry, mul2(rx, rx)));
  const f32x2 twoh = add2(p.soft, bc2(ssoft));
  float q0, q1, h0, h1;
  unpk2(rsq, q0, q1);
  unpk2(mul2(twoh, twoh), h0, h1);
  const bool far0 = q0 >= h0, far1 = q1 >= h1, some0 = q0 != 0.0f, some1 = q1 != 0.0f;
  float d0 = rsqrt_dev(q0), d1 = rsqrt_dev(q1);
  d0 = (far0 && some0) ? d0 : 0.0f;
  d1 = (far1 && some1) ? d1 : 0.0f;
  const f32x2 d = pk2(d0, d1);
  const f32x2 b = mul2(mul2(d, d), d);
  const f32x2 bm = mul2s(sm, b);
  ax = fma2(rx, bm, ax);
  ay = fma2(ry, bm, ay);
  az = fma2(rz, bm, az);
  pot = fma2(bc2(-sm), d, pot);
  float i0, i1;
  unpk2(mul2(add2(p.m, bc2(sm)), b), i0, i1);
  idt0 = fmaxf(idt0, i0);
  idt1 = fmaxf(idt1, i1);
  return (!far0 && some0) || (!far1 && some1);
}

/* the spline halves pp_pair2 left out: the scalar pp_pair (gravity.h:147-182) for exactly those */
__device__ __forceinline__ void pp_pair2_soft(float sx, float sy, float sz, float sm, float ssoft,
                                              const TargetSoftPair &p, f32x2 &ax, f32x2 &ay, f32x2 &az,
                                              f32x2 &pot, float &idt0, float &idt1) {
  float x0, x1, y0, y1, z0, z1, m0, m1, s0, s1;
  unpk2(p.x, x0, x1); unpk2(p.y, y0, y1); unpk2(p.z, z0, z1); unpk2(p.m, m0, m1); unpk2(p.soft, s0, s1);
  float a0, a1, b0, b1, c0, c1, e0, e1;
  unpk2(ax, a0, a1); unpk2(ay, b0, b1); unpk2(az, c0, c1); unpk2(pot, e0, e1);
  const float rx0 = sx - x0, ry0 = sy - y0, rz0 = sz - z0, rx1 = sx - x1, ry1 = sy - y1, rz1 = sz - z1;
  const float q0 = fmaf(rz0, rz0, fmaf(ry0, ry0, rx0 * rx0)), q1 = fmaf(rz1, rz1, fmaf(ry1, ry1, rx1 * rx1));
  const float t0 = s0 + ssoft, t1 = s1 + ssoft;
  const real4 p0 = {x0, y0, z0, m0}, p1 = {x1, y1, z1, m1};
  if (!(q0 >= t0 * t0) && q0 != 0.0f) pp_pair(sx, sy, sz, sm, ssoft, p0, s0, a0, b0, c0, e0, idt0);
  if (!(q1 >= t1 * t1) && q1 != 0.0f) pp_pair(sx, sy, sz, sm, ssoft, p1, s1, a1, b1, c1, e1, idt1);
  ax = pk2(a0, a1); ay = pk2(b0, b1); az = pk2(c0, c1); pot = pk2(e0, e1);
}

template <int PB>
constexpr size_t part_list_x2_smem_bytes() {
  return (size_t)kListWarps * (5 * PB * 32 * sizeof(float) + (PB / 2) * sizeof(TargetSoftPair));
}

template <int PB, int MINB>
__global__ void __launch_bounds__(kListWarps * 32, MINB)
part_list_x2_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                    const PackedPart *__restrict__ sources, const ILCell *__restrict__ list,
                    const int *__restrict__ markers, const int *__restrict__ starts,
                    const int *__restrict__ sizes, int nBuckets, float fperiod,
                    unsigned int *__restrict__ nextBucket) {
  static_assert(PB % 2 == 0 && 5 * PB <= 64, "two reduction rows per lane at most");
  constexpr int NP = PB / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *red = reinterpret_cast<float *>(smem_raw) + (size_t)warp * (5 * PB * 32);
  TargetSoftPair *sp = reinterpret_cast<TargetSoftPair *>(smem_raw + (size_t)kListWarps * 5 * PB * 32 * sizeof(float)) + warp * NP;
  const unsigned spAddr = pin_u32(smem_u32(sp)); /* one register + immediates: no per-body address rebuild */
  const ILCell none = {-1, 0};
  auto grab = [&]() { return grab_bucket(nextBucket, nBuckets, lane); };
  auto load_source = [&](const ILCell &e, float4 &pos, float &soft) {
    pos = make_float4(0.f, 0.f, 0.f, 0.f);
    soft = 0.f;
    if (e.index >= 0) {
      const PackedPart *q = sources + e.index;
      pos = *reinterpret_cast<const float4 *>(q);
      soft = q->soft;
    }
  };

  /* one bucket ahead, as in cell_list_x2_kernel: markers of bucket k+1 are loaded while k
   * runs; its first two list tiles, its first source rows and its targets are requested
   * during the last tile of k */
  int k = grab();
  BucketMeta m = {0, 0, 0, 0};
  if (k < nBuckets) m = load_bucket_meta(markers, starts, sizes, k);
  bool havePre = false;
  ILCell pre0 = none, pre1 = none;
  float4 preq = {0.f, 0.f, 0.f, 0.f}, pres = {0.f, 0.f, 0.f, 0.f};
  float preqSoft = 0.f, presSoft = 0.f;

  while (k < nBuckets) {
    const int kn = grab();
    BucketMeta mn = {0, 0, 0, 0};
    if (kn < nBuckets) mn = load_bucket_meta(markers, starts, sizes, kn);
    bool nextPre = false;
    ILCell npre0 = none, npre1 = none;
    float4 npreq = {0.f, 0.f, 0.f, 0.f}, npres = {0.f, 0.f, 0.f, 0.f};
    float npreqSoft = 0.f, npresSoft = 0.f;

    const ILCell *__restrict__ mylist = list + m.begin;
    const int len = m.len, ntiles = (len + 31) >> 5;

    for (int p0 = 0; p0 < m.count && len > 0; p0 += PB) {
      const int np = min(PB, m.count - p0);
      const int npairs = (np + 1) >> 1;
      const bool lastPass = p0 + PB >= m.count;
      const bool usePre = havePre && p0 == 0;
      __syncwarp();
      if (lane < 2 * npairs) {
        float4 v = preq;
        float vs = preqSoft;
        if (!usePre) {
          const PackedPart *q = parts + m.first + p0 + min(lane, np - 1);
          v = *reinterpret_cast<const float4 *>(q);
          vs = q->soft;
        }
        float *dst = reinterpret_cast<float *>(sp + (lane >> 1)) + (lane & 1);
        dst[0] = v.x; dst[2] = v.y; dst[4] = v.z; dst[6] = v.w; dst[8] = vs;
      }
      __syncwarp();

      f32x2 ax[NP], ay[NP], az[NP], pot[NP];
      float idt[PB];
#pragma unroll
      for (int j = 0; j < NP; ++j) { ax[j] = ay[j] = az[j] = pot[j] = 0ull; idt[2 * j] = idt[2 * j + 1] = 0.0f; }

      /* source rows one tile ahead in registers, list entries two ahead */
      ILCell cur = none, nxt = none;
      float4 s_pos;
      float s_soft;
      if (usePre) {
        cur = pre0; nxt = pre1; s_pos = pres; s_soft = presSoft;
      } else {
        if (lane < len) cur = mylist[lane];
        if (32 + lane < len) nxt = mylist[32 + lane];
        load_source(cur, s_pos, s_soft);
      }
      for (int t = 0; t < ntiles; ++t) {
        float4 n_pos;
        float n_soft;
        load_source(nxt, n_pos, n_soft);
        ILCell nn = none;
        if ((t + 2) * 32 + lane < len) nn = mylist[(t + 2) * 32 + lane];
        if (lastPass && t == ntiles - 1 && mn.len > 0) {
          const ILCell *nl = list + mn.begin;
          if (lane < mn.len) npre0 = nl[lane];
          if (32 + lane < mn.len) npre1 = nl[32 + lane];
          const int npn = min(PB, mn.count);
          if (lane < 2 * ((npn + 1) >> 1)) {
            const PackedPart *q = parts + mn.first + min(lane, npn - 1);
            npreq = *reinterpret_cast<const float4 *>(q);
            npreqSoft = q->soft;
          }
          nextPre = true;
        }
        if (cur.index >= 0) {
          const float sx = fmaf(float(replica_x(cur.offsetID)), fperiod, s_pos.x);
          const float sy = fmaf(float(replica_y(cur.offsetID)), fperiod, s_pos.y);
          const float sz = fmaf(float(replica_z(cur.offsetID)), fperiod, s_pos.z);
          unsigned needSoft = 0;
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            if (j < npairs) {
              const TargetSoftPair p = lds_target_soft_pair(spAddr + j * (unsigned)sizeof(TargetSoftPair));
              if (pp_pair2(sx, sy, sz, s_pos.w, s_soft, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1]))
                needSoft |= 1u << j;
            }
          }
          if (needSoft) { /* rare: some pair of this lane is inside the softening length */
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              if ((needSoft >> j) & 1u) {
                const TargetSoftPair p = lds_target_soft_pair(spAddr + j * (unsigned)sizeof(TargetSoftPair));
                pp_pair2_soft(sx, sy, sz, s_pos.w, s_soft, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1]);
              }
            }
          }
        }
        if (nextPre && lastPass && t == ntiles - 1) load_source(npre0, npres, npresSoft); /* after the math: its index has landed */
        cur = nxt; nxt = nn;
        s_pos = n_pos; s_soft = n_soft;
      }

      __syncwarp();
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (j < npairs) {
          float a0, a1, b0, b1, c0, c1, e0, e1;
          unpk2(ax[j], a0, a1); unpk2(ay[j], b0, b1); unpk2(az[j], c0, c1); unpk2(pot[j], e0, e1);
          float *r0 = red + (size_t)(2 * j) * 5 * 32 + lane;
          r0[0] = a0; r0[32] = b0; r0[64] = c0; r0[96] = e0; r0[128] = idt[2 * j];
          r0[160] = a1; r0[192] = b1; r0[224] = c1; r0[256] = e1; r0[288] = idt[2 * j + 1];
        }
      }
      float *out = reinterpret_cast<float *>(vars + m.first + p0);
      float old[2]; /* the accumulators' current values: loaded under the shared-memory reduction */
#pragma unroll
      for (int h = 0; h < 2; ++h) old[h] = (lane + 32 * h < 5 * np) ? out[lane + 32 * h] : 0.0f;
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int v = lane + 32 * h;
        if (v < 5 * np) {
          const float *row = red + v * 32;
          const bool isMax = (v % 5) == 4;
          float acc = 0.0f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = row[(i + lane) & 31];
            acc = isMax ? fmaxf(acc, x) : acc + x;
          }
          out[v] = isMax ? fmaxf(old[h], acc) : old[h] + acc;
        }
      }
      __syncwarp();
    }
    k = kn; m = mn;
    havePre = nextPre; pre0 = npre0; pre1 = npre1; preq = npreq; preqSoft = npreqSoft; pres = npres; presSoft = npresSoft;
  }
}
#endif /* !CUDA_USE_DOUBLE */

/* ------------------------------------------------------------------ Ewald */
struct EwaldParams {
  EwaldReadOnlyData ro;
  EwtData ewt[NEWH];
};

__device__ __forceinline__ float erfc_dev(float x) { return erfcf(x); }
__device__ __forceinline__ double erfc_dev(double x) { return erfc(x); }
__device__ __forceinline__ float erf_dev(float x) { return erff(x); }
__device__ __forceinline__ double erf_dev(double x) { return erf(x); }
__device__ __forceinline__ float exp_dev(float x) { return expf(x); }
__device__ __forceinline__ double exp_dev(double x) { return exp(x); }
__device__ __forceinline__ void sincos_dev(float x, float *s, float *c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_dev(double x, double *s, double *c) { sincos(x, s, c); }

/* g0 = (erfc(x) or, inside the hole, -erf(x) = erfc(x) - 1) * dir and a = exp(-x^2) * scale.
 * Float build: erfc(x) = t P(u) exp(-x^2) with t = 1/(1 + x/2), u = A t + B and P of degree 8
 * fitted to erfc(x) exp(x^2)/t on [0.75, 6] (1.2e-7 relative in float arithmetic, the accuracy
 * of erfcf), sharing its one exp(-x^2) = ex2(-x^2 log2 e) with the Gaussian term: ~18
 * instructions instead of ~50 for erfcf + expf.  x >= 0.8 here (smaller x takes the series) and
 * x <= 2 fEwCut; erfc(6) = 2e-17, so t is clamped there.  Double build: the library functions. */
__device__ __forceinline__ void ewald_erfc_gauss(float x, float x2, bool hole, float scale, float dir,
                                                 float &g0, float &a) {
  float ex, t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * x2));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.5f, fminf(x, 6.0f), 1.0f)));
  const float u = fmaf(4.190476190e+00f, t, -2.047619048e+00f);
  float p = -7.590534245e-07f;
  p = fmaf(p, u, 2.597246448e-06f); p = fmaf(p, u, 1.730812692e-05f); p = fmaf(p, u, -8.525551993e-05f);
  p = fmaf(p, u, -5.383136886e-04f); p = fmaf(p, u, 2.085378610e-03f); p = fmaf(p, u, 3.153986530e-02f);
  p = fmaf(p, u, 1.609637792e-01f); p = fmaf(p, u, 5.030546696e-01f);
  const float e = t * p * ex;
  g0 = (hole ? e - 1.0f : e) * dir;
  a = ex * scale;
}
__device__ __forceinline__ void ewald_erfc_gauss(double x, double x2, bool hole, double scale, double dir,
                                                 double &g0, double &a) {
  a = exp(-x2) * scale;
  g0 = (hole ? -erf(x) : erfc(x)) * dir;
}
/* sin and cos of h.x (|h.x| < ~20): one Cody-Waite reduction to [-pi, pi], then the SFU
 * (absolute error 4e-7, which is what the reference's -use_fast_math sincosf gives) */
__device__ __forceinline__ void ewald_sincos(float x, float &s, float &c) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);
  r = fmaf(-k, -1.7484555314695172e-07f, r);
  s = __sinf(r);
  c = __cosf(r);
}
__device__ __forceinline__ void ewald_sincos(double x, double &s, double &c) { sincos(x, &s, &c); }

constexpr int kEwaldThreads = 128;

/* sum_j (-x)^j / (j! (2j + 2n + 1)), Horner from the highest term; enough terms
 * for rounding-level accuracy up to x = kEwaldSeriesX in the build's precision */
constexpr double kEwaldSeriesX = 0.64;
constexpr int kEwaldSeriesTerms = sizeof(real) == 4 ? 10 : 19;
template <int N>
__device__ __forceinline__ real ewald_series(real x) {
  double fact = 1.0;
  for (int j = 1; j < kEwaldSeriesTerms; ++j) fact *= j;
  real s = real(1.0 / (fact * (2 * (kEwaldSeriesTerms - 1) + 2 * N + 1)));
#pragma unroll
  for (int j = kEwaldSeriesTerms - 2; j >= 0; --j) {
    fact /= (j + 1);
    s = fma(-x, s, real(1.0 / (fact * (2 * j + 2 * N + 1))));
  }
  return s;
}

/* One thread per active particle: real-space sum over the (2 nEwReps+1)^3
 * replicas of the root cell's complete hexadecapole expansion, then the
 * reciprocal-space sum over the h-table (BucketEwald, Ewald.cpp:72-281;
 * EwaldKernel, HostCUDA.cu:1958-2192).  Constants and the h-table arrive as a
 * __grid_constant__ argument (constant bank, uniform reads) instead of the
 * reference's process-global __constant__ symbols, so concurrent requests on
 * different streams cannot race.  Adds to acc/pot, never touches dtGrav. */
__device__ __forceinline__ void ewald_particle(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                                               const int *__restrict__ markers, int first, int last,
                                               const EwaldParams &P) {
  int id = blockIdx.x * kEwaldThreads + threadIdx.x;
  if (markers) {
    if (id > last) return;
    id = markers[id];
  } else {
    id += first;
    if (id > last) return;
  }
  const EwaldReadOnlyData &ro = P.ro;
  const MomcData &q = ro.momcRoot;
  const real third = real(1.0 / 3.0), half = real(0.5);

  const real Q4xx = half * (q.xxxx + q.xxyy + q.xxzz);
  const real Q4xy = half * (q.xxxy + q.xyyy + q.xyzz);
  const real Q4xz = half * (q.xxxz + q.xyyz + q.xzzz);
  const real Q4yy = half * (q.xxyy + q.yyyy + q.yyzz);
  const real Q4yz = half * (q.xxyz + q.yyyz + q.yzzz);
  const real Q4zz = half * (q.xxzz + q.yyzz + q.zzzz);
  const real Q4 = real(0.25) * (Q4xx + Q4yy + Q4zz);
  const real Q3x = half * (q.xxx + q.xyy + q.xzz);
  const real Q3y = half * (q.xxy + q.yyy + q.yzz);
  const real Q3z = half * (q.xxz + q.yyz + q.zzz);
  const real Q2 = half * (q.xx + q.yy + q.zz);

  const real4 p = *reinterpret_cast<const real4 *>(parts + id);
  const real dx = p.x - ro.mm.cmx, dy = p.y - ro.mm.cmy, dz = p.z - ro.mm.cmz;
  real fPot = ro.mm.totalMass * ro.k1, ax = 0, ay = 0, az = 0;
  const int nE = ro.nEwReps, nR = ro.nReps;
  const real L = ro.L, alpha = ro.alpha, alpha2 = ro.alpha2, ka = ro.ka;
  const real twoa2 = 2 * alpha2;

  for (int ix = -nE; ix <= nE; ++ix) {
    const bool hx = (ix >= -nR && ix <= nR);
    const real x = dx + ix * L;
    for (int iy = -nE; iy <= nE; ++iy) {
      const bool hxy = hx && (iy >= -nR && iy <= nR);
      const real y = dy + iy * L;
      /* a column whose axis already misses the cut sphere holds no term (x^2 + y^2 <= r^2 in
       * floating point too: the sum is monotone); 28 of the 49 columns end here */
      if (x * x + y * y > ro.fEwCut2 && !hxy) continue;
      /* Along a column only z changes, and every contraction of the root moments with the
       * displacement is a polynomial in z whose coefficients belong to the column:
       *   Q2m_i = P2_i + q_iz z,   Q3m_i = P3_i0 + P3_i1 z + q_izz z^2/2,
       *   Q4m_i = P4_i0 + P4_i1 z + P4_i2 z^2/2 + q_izzz z^3/6,   Q4_i = P4a_i + Q4_iz z.
       * 63 operations per column buy 24 instead of 88 per replica (same terms as
       * HostCUDA.cu:2098-2136, regrouped). */
      const real xx = half * x * x, xxx = third * xx * x, xxy = xx * y;
      const real yy = half * y * y, yyy = third * yy * y, xyy = yy * x;
      const real xy = x * y;
      const real P2x = q.xx * x + q.xy * y, P2y = q.xy * x + q.yy * y, P2z = q.xz * x + q.yz * y;
      const real P3x0 = q.xxx * xx + q.xxy * xy + q.xyy * yy, P3x1 = q.xxz * x + q.xyz * y;
      const real P3y0 = q.xxy * xx + q.xyy * xy + q.yyy * yy, P3y1 = q.xyz * x + q.yyz * y;
      const real P3z0 = q.xxz * xx + q.xyz * xy + q.yyz * yy, P3z1 = q.xzz * x + q.yzz * y;
      const real P4x0 = q.xxxx * xxx + q.xxxy * xxy + q.xxyy * xyy + q.xyyy * yyy;
      const real P4y0 = q.xxxy * xxx + q.xxyy * xxy + q.xyyy * xyy + q.yyyy * yyy;
      const real P4z0 = q.xxxz * xxx + q.xxyz * xxy + q.xyyz * xyy + q.yyyz * yyy;
      const real P4x1 = q.xxxz * xx + q.xxyz * xy + q.xyyz * yy;
      const real P4y1 = q.xxyz * xx + q.xyyz * xy + q.yyyz * yy;
      const real P4z1 = q.xxzz * xx + q.xyzz * xy + q.yyzz * yy;
      const real P4x2 = q.xxzz * x + q.xyzz * y, P4y2 = q.xyzz * x + q.yyzz * y, P4z2 = q.xzzz * x + q.yzzz * y;
      const real P4ax = Q4xx * x + Q4xy * y, P4ay = Q4xy * x + Q4yy * y, P4az = Q4xz * x + Q4yz * y;
      const real xy2 = x * x + y * y;
      for (int iz = -nE; iz <= nE; ++iz) {
        const bool hole = hxy && (iz >= -nR && iz <= nR);
        const real z = dz + iz * L;
        const real r2 = xy2 + z * z;
        if (r2 > ro.fEwCut2 && !hole) continue;
        real g0, g1, g2, g3, g4, g5;
        const real xa = r2 * alpha2;
        if (hole && xa < real(kEwaldSeriesX)) {
          /* -erf(alpha r)/r and its (1/r d/dr)^n derivatives as a power series in
           * x = alpha^2 r^2:  g_n = -ka (2 alpha^2)^n sum_j (-x)^j / (j! (2j+2n+1)).
           * The reference switches to the first two terms of this series only for
           * r^2 < fInner2 (Ewald.cpp:141-152; 1.2e-3 L^2 on the CPU, 1.1e-2 L^2 on its
           * GPU, Ewald.cpp:516) and uses the erf/exp recurrences beyond, which in
           * single precision cancel catastrophically just outside that radius
           * (measured: 2.5e-3 relative force error at r = 0.11 L).  Summing the series
           * to rounding over the whole well-conditioned range removes both the
           * truncation and the cancellation error; ro.fInner2 is not used. */
          real an = -ka;
          g0 = an * ewald_series<0>(xa);
          an *= twoa2; g1 = an * ewald_series<1>(xa);
          an *= twoa2; g2 = an * ewald_series<2>(xa);
          an *= twoa2; g3 = an * ewald_series<3>(xa);
          an *= twoa2; g4 = an * ewald_series<4>(xa);
          an *= twoa2; g5 = an * ewald_series<5>(xa);
        } else {
          const real dir = rsqrt_dev(r2), dir2 = dir * dir;
          const real r = r2 * dir;
          real a;
          ewald_erfc_gauss(alpha * r, xa, hole, ka * dir2, dir, g0, a);
          g1 = g0 * dir2 + a;
          real an = twoa2;
          g2 = 3 * g1 * dir2 + an * a;
          an *= twoa2; g3 = 5 * g2 * dir2 + an * a;
          an *= twoa2; g4 = 7 * g3 * dir2 + an * a;
          an *= twoa2; g5 = 9 * g4 * dir2 + an * a;
        }
        const real zz = half * z * z, zzz = third * zz * z;
        const real Q2mx = fma(q.xz, z, P2x), Q2my = fma(q.yz, z, P2y), Q2mz = fma(q.zz, z, P2z);
        const real Q3mx = fma(q.xzz, zz, fma(P3x1, z, P3x0));
        const real Q3my = fma(q.yzz, zz, fma(P3y1, z, P3y0));
        const real Q3mz = fma(q.zzz, zz, fma(P3z1, z, P3z0));
        const real Q4mx = fma(q.xzzz, zzz, fma(P4x2, zz, fma(P4x1, z, P4x0)));
        const real Q4my = fma(q.yzzz, zzz, fma(P4y2, zz, fma(P4y1, z, P4y0)));
        const real Q4mz = fma(q.zzzz, zzz, fma(P4z2, zz, fma(P4z1, z, P4z0)));
        const real Q4x = fma(Q4xz, z, P4ax), Q4y = fma(Q4yz, z, P4ay), Q4z = fma(Q4zz, z, P4az);
        const real Q2m = half * (Q2mx * x + Q2my * y + Q2mz * z) - (Q3x * x + Q3y * y + Q3z * z) + Q4;
        const real Q3m = third * (Q3mx * x + Q3my * y + Q3mz * z) - half * (Q4x * x + Q4y * y + Q4z * z);
        const real Q4m = real(0.25) * (Q4mx * x + Q4my * y + Q4mz * z);
        const real Qta = g1 * q.m - g2 * Q2 + g3 * Q2m + g4 * Q3m + g5 * Q4m;
        fPot -= g0 * q.m - g1 * Q2 + g2 * Q2m + g3 * Q3m + g4 * Q4m;
        ax += g2 * (Q2mx - Q3x) + g3 * (Q3mx - Q4x) + g4 * Q4mx - x * Qta;
        ay += g2 * (Q2my - Q3y) + g3 * (Q3my - Q4y) + g4 * Q4my - y * Qta;
        az += g2 * (Q2mz - Q3z) + g3 * (Q3mz - Q4z) + g4 * Q4mz - z * Qta;
      }
    }
  }

  /* reciprocal space (Ewald.cpp:265-273) */
  const int nh = ro.nEwhLoop;
  for (int i = 0; i < nh; ++i) {
    const EwtData &e = P.ewt[i];
    const real hdotx = e.hx * dx + e.hy * dy + e.hz * dz;
    real s, c;
    ewald_sincos(hdotx, s, c);
    fPot += e.hCfac * c + e.hSfac * s;
    const real w = e.hCfac * s - e.hSfac * c;
    ax += e.hx * w; ay += e.hy * w; az += e.hz * w;
  }

  VariablePartData *v = vars + id;
  v->a.x += ax; v->a.y += ay; v->a.z += az;
  v->potential += fPot;
}

__global__ void __launch_bounds__(kEwaldThreads)
ewald_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
             const int *__restrict__ markers, int first, int last,
             const __grid_constant__ EwaldParams P) {
  ewald_particle(parts, vars, markers, first, last, P);
}

/* The same kernel for a force step that builds its Ewald tables ON THE DEVICE (ewald_setup_kernel,
 * hostcuda.cu): the parameters sit in one of a few __constant__ slots, filled by a stream-ordered
 * device-to-device copy, so the root moments never travel to the host and the step needs no
 * synchronisation in front of the Ewald launch.  A slot belongs to one step object at a time. */
constexpr int kEwaldSlots = 4;
__constant__ EwaldParams c_ewaldSlot[kEwaldSlots];
__global__ void __launch_bounds__(kEwaldThreads)
ewald_slot_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                  const int *__restrict__ markers, int first, int last, int slot) {
  ewald_particle(parts, vars, markers, first, last, c_ewaldSlot[slot]);
}


}  // namespace cb200
#endif
