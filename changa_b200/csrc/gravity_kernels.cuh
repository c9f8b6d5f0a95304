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

template <int PB>
constexpr size_t cell_list_smem_bytes() {
  return (size_t)kListWarps * (2 * 32 * kCellBytes + PB * sizeof(real4));
}

/* ---------------------------------------------------------- particle-cell */
template <int PB, int MINB>
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
    int k = 0;
    if (lane == 0) k = (int)atomicAdd(nextBucket, 1u);
    k = __shfl_sync(kFull, k, 0);
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
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            if (j < np) {
              const real4 p = sp[j];
              pc_pair(c, ccx, ccy, ccz, p, ax[j], ay[j], az[j], pot[j], idt[j]);
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
    int k = 0;
    if (lane == 0) k = (int)atomicAdd(nextBucket, 1u);
    k = __shfl_sync(kFull, k, 0);
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

constexpr int kEwaldThreads = 128;

/* One thread per active particle: real-space sum over the (2 nEwReps+1)^3
 * replicas of the root cell's complete hexadecapole expansion, then the
 * reciprocal-space sum over the h-table (BucketEwald, Ewald.cpp:72-281;
 * EwaldKernel, HostCUDA.cu:1958-2192).  Constants and the h-table arrive as a
 * __grid_constant__ argument (constant bank, uniform reads) instead of the
 * reference's process-global __constant__ symbols, so concurrent requests on
 * different streams cannot race.  Adds to acc/pot, never touches dtGrav. */
__global__ void __launch_bounds__(kEwaldThreads)
ewald_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
             const int *__restrict__ markers, int first, int last,
             const __grid_constant__ EwaldParams P) {
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
      for (int iz = -nE; iz <= nE; ++iz) {
        const bool hole = hxy && (iz >= -nR && iz <= nR);
        const real z = dz + iz * L;
        real r2 = x * x + y * y + z * z;
        if (r2 > ro.fEwCut2 && !hole) continue;
        real g0, g1, g2, g3, g4, g5;
        if (r2 < ro.fInner2) { /* series about r = 0 (Ewald.cpp:141-152) */
          real an = ka;
          r2 *= alpha2;
          g0 = an * (third * r2 - real(1));
          an *= twoa2; g1 = an * (real(1.0 / 5.0) * r2 - third);
          an *= twoa2; g2 = an * (real(1.0 / 7.0) * r2 - real(1.0 / 5.0));
          an *= twoa2; g3 = an * (real(1.0 / 9.0) * r2 - real(1.0 / 7.0));
          an *= twoa2; g4 = an * (real(1.0 / 11.0) * r2 - real(1.0 / 9.0));
          an *= twoa2; g5 = an * (real(1.0 / 13.0) * r2 - real(1.0 / 11.0));
        } else {
          const real dir = rsqrt_dev(r2), dir2 = dir * dir;
          const real r = r2 * dir;
          real a = exp_dev(-r2 * alpha2) * ka * dir2;
          g0 = (hole ? -erf_dev(alpha * r) : erfc_dev(alpha * r)) * dir;
          g1 = g0 * dir2 + a;
          real an = twoa2;
          g2 = 3 * g1 * dir2 + an * a;
          an *= twoa2; g3 = 5 * g2 * dir2 + an * a;
          an *= twoa2; g4 = 7 * g3 * dir2 + an * a;
          an *= twoa2; g5 = 9 * g4 * dir2 + an * a;
        }
        const real xx = half * x * x, xxx = third * xx * x, xxy = xx * y, xxz = xx * z;
        const real yy = half * y * y, yyy = third * yy * y, xyy = yy * x, yyz = yy * z;
        const real zz = half * z * z, zzz = third * zz * z, xzz = zz * x, yzz = zz * y;
        const real xy = x * y, xyz = xy * z, xz = x * z, yz = y * z;
        const real Q2mx = q.xx * x + q.xy * y + q.xz * z;
        const real Q2my = q.xy * x + q.yy * y + q.yz * z;
        const real Q2mz = q.xz * x + q.yz * y + q.zz * z;
        const real Q3mx = q.xxx * xx + q.xxy * xy + q.xxz * xz + q.xyy * yy + q.xyz * yz + q.xzz * zz;
        const real Q3my = q.xxy * xx + q.xyy * xy + q.xyz * xz + q.yyy * yy + q.yyz * yz + q.yzz * zz;
        const real Q3mz = q.xxz * xx + q.xyz * xy + q.xzz * xz + q.yyz * yy + q.yzz * yz + q.zzz * zz;
        const real Q4mx = q.xxxx * xxx + q.xxxy * xxy + q.xxxz * xxz + q.xxyy * xyy + q.xxyz * xyz +
                          q.xxzz * xzz + q.xyyy * yyy + q.xyyz * yyz + q.xyzz * yzz + q.xzzz * zzz;
        const real Q4my = q.xxxy * xxx + q.xxyy * xxy + q.xxyz * xxz + q.xyyy * xyy + q.xyyz * xyz +
                          q.xyzz * xzz + q.yyyy * yyy + q.yyyz * yyz + q.yyzz * yzz + q.yzzz * zzz;
        const real Q4mz = q.xxxz * xxx + q.xxyz * xxy + q.xxzz * xxz + q.xyyz * xyy + q.xyzz * xyz +
                          q.xzzz * xzz + q.yyyz * yyy + q.yyzz * yyz + q.yzzz * yzz + q.zzzz * zzz;
        const real Q4x = Q4xx * x + Q4xy * y + Q4xz * z;
        const real Q4y = Q4xy * x + Q4yy * y + Q4yz * z;
        const real Q4z = Q4xz * x + Q4yz * y + Q4zz * z;
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
    sincos_dev(hdotx, &s, &c);
    fPot += e.hCfac * c + e.hSfac * s;
    const real w = e.hCfac * s - e.hSfac * c;
    ax += e.hx * w; ay += e.hy * w; az += e.hz * w;
  }

  VariablePartData *v = vars + id;
  v->a.x += ax; v->a.y += ay; v->a.z += az;
  v->potential += fPot;
}

}  // namespace cb200
#endif
