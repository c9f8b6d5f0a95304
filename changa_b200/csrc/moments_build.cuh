/* moments_build.cuh -- tree-moment build on the device.
 *
 * The reference builds moments on the host only (GenericTreeNode.h:219-256,
 * MultipoleMoments.h:207-352,460-526 on top of moments.c); this is the same
 * arithmetic, in double, one thread per tree node, one launch per tree level
 * from the leaves up:
 *   bucket   : particles added one at a time in index order -- shift the running
 *              expansion to the new centre of mass (momShiftFmomr,
 *              moments.c:1002-1056), add the particle's own moments
 *              (momMakeFmomr :772-835) -- then radius := farthest particle;
 *   internal : children combined in order 0,1 (shift both to the joint centre,
 *              momScaledAddFmomr :202-232), radius := farthest box corner.
 * Components are stored scaled by the node radius (FMOMR convention).  The
 * single-precision literals (0.2f, 1.0f/7.0f ...) are the reference's own and
 * are kept so that results agree with the host build to rounding.
 */
#ifndef CB200_MOMENTS_BUILD_CUH
#define CB200_MOMENTS_BUILD_CUH

#include "device_layout.cuh"

namespace cb200 {

/* reduced (trace-free) scaled moment components, order of moments.h:68-73 */
enum { F_M, F_XX, F_YY, F_XY, F_XZ, F_YZ,
       F_XXX, F_XYY, F_XXY, F_YYY, F_XXZ, F_YYZ, F_XYZ,
       F_XXXX, F_XYYY, F_XXXY, F_YYYY, F_XXXZ, F_YYYZ, F_XXYY, F_XXYZ, F_XYYZ, F_N };

struct MomentNode {
  double radius, soft, mass, cm[3];
  double f[F_N];
};

__host__ __device__ inline void fm_point(double *r, double m, double u, double x, double y, double z) {
  const double iu = 1.0f / u;
  x *= iu; y *= iu; z *= iu;
  const double x2 = x * x, y2 = y * y, d2 = x2 + y2 + z * z;
  double tx = m * x, ty = m * y;
  r[F_M] = m;
  r[F_XY] = tx * y; r[F_XZ] = tx * z; r[F_YZ] = ty * z;
  tx *= x; ty *= y;
  m *= d2;
  double t = (1.0f / 3.0f) * m;
  r[F_XX] = tx - t; r[F_YY] = ty - t;
  t = 0.2f * m;
  double dx = tx - t, dy = ty - t;
  r[F_XXY] = dx * y; r[F_XXZ] = dx * z; r[F_YYZ] = dy * z; r[F_XYY] = dy * x;
  r[F_XYZ] = r[F_XY] * z;
  t *= 3.0f;
  r[F_XXX] = (tx - t) * x; r[F_YYY] = (ty - t) * y;
  t = (1.0f / 7.0f) * m;
  r[F_XXYZ] = (tx - t) * y * z; r[F_XYYZ] = (ty - t) * x * z;
  dx = (tx - 3.0f * t) * x; dy = (ty - 3.0f * t) * y;
  r[F_XXXY] = dx * y; r[F_XXXZ] = dx * z; r[F_XYYY] = dy * x; r[F_YYYZ] = dy * z;
  dx = t * (x2 - 0.1f * d2); dy = t * (y2 - 0.1f * d2);
  r[F_XXXX] = tx * x2 - 6.0f * dx; r[F_YYYY] = ty * y2 - 6.0f * dy;
  r[F_XXYY] = tx * y2 - dx - dy;
}

/* components of order 2, 3, 4 follow F_M in runs of 5, 7, 9 */
__host__ __device__ inline void fm_scale_orders(double *r, const double *a, double ratio, double w, bool accumulate) {
  /* three runs with constant bounds: fully unrolled on the device, so the expansions stay in
   * registers (a run-length table and a running index put them in local memory) */
  const double s2 = ratio * w * ratio, s3 = s2 * ratio, s4 = s3 * ratio;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = F_XX; k < F_XXX; ++k) r[k] = accumulate ? r[k] + s2 * a[k] : s2 * a[k];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = F_XXX; k < F_XXXX; ++k) r[k] = accumulate ? r[k] + s3 * a[k] : s3 * a[k];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = F_XXXX; k < F_N; ++k) r[k] = accumulate ? r[k] + s4 * a[k] : s4 * a[k];
}

__host__ __device__ inline void fm_rescale(double *r, double unew, double uold) {
  fm_scale_orders(r, r, uold / unew, 1.0, false);
}

__host__ __device__ inline void fm_scaled_add(double *r, double ur, const double *a, double ua) {
  r[F_M] += a[F_M];
  fm_scale_orders(r, a, ua / ur, 1.0, true);
}

/* move the expansion centre by -(x,y,z): (x,y,z) = old centre - new centre */
__host__ __device__ inline void fm_shift(double *m, double u, double x, double y, double z) {
  double f[F_N];
  const double c27 = 2.0f / 7.0f;
  fm_point(f, 1.0f, u, x, y, z);
  const double iu = 1.0f / u;
  x *= iu; y *= iu; z *= iu;
  const double tx = 0.4f * (m[F_XX] * x + m[F_XY] * y + m[F_XZ] * z);
  const double ty = 0.4f * (m[F_XY] * x + m[F_YY] * y + m[F_YZ] * z);
  const double tz = 0.4f * (m[F_XZ] * x + m[F_YZ] * y - (m[F_XX] + m[F_YY]) * z);
  const double t = tx * x + ty * y + tz * z;
  const double txx = c27 * (m[F_XXX] * x + m[F_XXY] * y + m[F_XXZ] * z +
                            2.0f * (m[F_XX] * f[F_XX] + m[F_XY] * f[F_XY] + m[F_XZ] * f[F_XZ]) - 0.5f * t);
  const double tyy = c27 * (m[F_XYY] * x + m[F_YYY] * y + m[F_YYZ] * z +
                            2.0f * (m[F_XY] * f[F_XY] + m[F_YY] * f[F_YY] + m[F_YZ] * f[F_YZ]) - 0.5f * t);
  const double txy = c27 * (m[F_XXY] * x + m[F_XYY] * y + m[F_XYZ] * z + m[F_XY] * (f[F_XX] + f[F_YY]) +
                            (m[F_XX] + m[F_YY]) * f[F_XY] + m[F_YZ] * f[F_XZ] + m[F_XZ] * f[F_YZ]);
  const double tyz = c27 * (m[F_XYZ] * x + m[F_YYZ] * y - (m[F_XXY] + m[F_YYY]) * z - m[F_YZ] * f[F_XX] -
                            m[F_XX] * f[F_YZ] + m[F_XZ] * f[F_XY] + m[F_XY] * f[F_XZ]);
  const double txz = c27 * (m[F_XXZ] * x + m[F_XYZ] * y - (m[F_XXX] + m[F_XYY]) * z - m[F_XZ] * f[F_YY] -
                            m[F_YY] * f[F_XZ] + m[F_YZ] * f[F_XY] + m[F_XY] * f[F_YZ]);
  /* order 4 first: it reads the not-yet-shifted orders 2 and 3 */
  m[F_XXXX] += 4.0f * m[F_XXX] * x + 6.0f * (m[F_XX] * f[F_XX] - txx);
  m[F_YYYY] += 4.0f * m[F_YYY] * y + 6.0f * (m[F_YY] * f[F_YY] - tyy);
  m[F_XYYY] += m[F_YYY] * x + 3.0f * (m[F_XYY] * y + m[F_YY] * f[F_XY] + m[F_XY] * f[F_YY] - txy);
  m[F_XXXY] += m[F_XXX] * y + 3.0f * (m[F_XXY] * x + m[F_XX] * f[F_XY] + m[F_XY] * f[F_XX] - txy);
  m[F_XXXZ] += m[F_XXX] * z + 3.0f * (m[F_XXZ] * x + m[F_XX] * f[F_XZ] + m[F_XZ] * f[F_XX] - txz);
  m[F_YYYZ] += m[F_YYY] * z + 3.0f * (m[F_YYZ] * y + m[F_YY] * f[F_YZ] + m[F_YZ] * f[F_YY] - tyz);
  m[F_XXYY] += 2.0f * (m[F_XXY] * y + m[F_XYY] * x) + m[F_XX] * f[F_YY] + m[F_YY] * f[F_XX] +
               4.0f * m[F_XY] * f[F_XY] - txx - tyy;
  m[F_XXYZ] += m[F_XXY] * z + m[F_XXZ] * y + m[F_XX] * f[F_YZ] + m[F_YZ] * f[F_XX] +
               2.0f * (m[F_XYZ] * x + m[F_XY] * f[F_XZ] + m[F_XZ] * f[F_XY]) - tyz;
  m[F_XYYZ] += m[F_XYY] * z + m[F_YYZ] * x + m[F_YY] * f[F_XZ] + m[F_XZ] * f[F_YY] +
               2.0f * (m[F_XYZ] * y + m[F_XY] * f[F_YZ] + m[F_YZ] * f[F_XY]) - txz;
  m[F_XXX] += 3.0f * (m[F_XX] * x - tx);
  m[F_XYY] += 2.0f * m[F_XY] * y + m[F_YY] * x - tx;
  m[F_YYY] += 3.0f * (m[F_YY] * y - ty);
  m[F_XXY] += 2.0f * m[F_XY] * x + m[F_XX] * y - ty;
  m[F_XXZ] += 2.0f * m[F_XZ] * x + m[F_XX] * z - tz;
  m[F_YYZ] += 2.0f * m[F_YZ] * y + m[F_YY] * z - tz;
  m[F_XYZ] += m[F_XY] * z + m[F_XZ] * y + m[F_YZ] * x;
  /* the monopole, displaced from the new centre, feeds every higher order */
  const double M = m[F_M];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = F_XX; k < F_N; ++k) m[k] += M * f[k];
}

/* MultipoleMoments::operator+=(particle), MultipoleMoments.h:276-329 */
__host__ __device__ inline void node_add_particle(MomentNode &n, double pm, double ps, const double *px) {
  const double m1 = n.mass;
  n.mass += pm;
  if (n.mass == 0.0) {
    n.soft = 0.5 * (n.soft + ps);
    for (int d = 0; d < 3; ++d) n.cm[d] = 0.5 * (n.cm[d] + px[d]);
    return;
  }
  n.soft = (m1 * n.soft + pm * ps) / n.mass;
  double old[3] = {n.cm[0], n.cm[1], n.cm[2]};
  for (int d = 0; d < 3; ++d) n.cm[d] = (m1 * n.cm[d] + pm * px[d]) / n.mass;
  fm_shift(n.f, n.radius, old[0] - n.cm[0], old[1] - n.cm[1], old[2] - n.cm[2]);
  double one[F_N];
  fm_point(one, pm, n.radius, px[0] - n.cm[0], px[1] - n.cm[1], px[2] - n.cm[2]);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 0; k < F_N; ++k) n.f[k] += one[k];
}

/* MultipoleMoments::operator+=(moments), MultipoleMoments.h:207-253 */
__host__ __device__ inline void node_add_node(MomentNode &n, const MomentNode &o) {
  const double m1 = n.mass;
  n.mass += o.mass;
  if (n.mass == 0.0) {
    n.soft = 0.5 * (n.soft + o.soft);
    for (int d = 0; d < 3; ++d) n.cm[d] = 0.5 * (n.cm[d] + o.cm[d]);
    return;
  }
  if (m1 == 0.0) { n = o; return; }
  if (o.mass == 0.0) return;
  n.soft = (m1 * n.soft + o.mass * o.soft) / n.mass;
  double old[3] = {n.cm[0], n.cm[1], n.cm[2]};
  for (int d = 0; d < 3; ++d) n.cm[d] = (m1 * n.cm[d] + o.mass * o.cm[d]) / n.mass;
  fm_shift(n.f, n.radius, old[0] - n.cm[0], old[1] - n.cm[1], old[2] - n.cm[2]);
  double g[F_N];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 0; k < F_N; ++k) g[k] = o.f[k];
  fm_shift(g, o.radius, o.cm[0] - n.cm[0], o.cm[1] - n.cm[1], o.cm[2] - n.cm[2]);
  fm_scaled_add(n.f, n.radius, g, o.radius);
}

__host__ __device__ inline void node_clear(MomentNode &n) {
  n.radius = n.soft = n.mass = 0.0;
  n.cm[0] = n.cm[1] = n.cm[2] = 0.0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 0; k < F_N; ++k) n.f[k] = 0.0;
}

/* makeBucket, GenericTreeNode.h:219-256 */
__host__ __device__ inline void node_make_bucket(MomentNode &n, const double *pos, const double *mass,
                                                 const double *soft, int first, int last,
                                                 const double *geolo, const double *geohi) {
  node_clear(n);
  const double e0 = geohi[0] - geolo[0], e1 = geohi[1] - geolo[1], e2 = geohi[2] - geolo[2];
  n.radius = 0.5 * sqrt(e0 * e0 + e1 * e1 + e2 * e2);
  const int count = last - first + 1;
  if (n.radius <= 0.0) {
    if (count > 1) {
      double best = 0.0;
      for (int i = first + 1; i <= last; ++i) {
        const double a = pos[3 * first] - pos[3 * i], b = pos[3 * first + 1] - pos[3 * i + 1],
                     c = pos[3 * first + 2] - pos[3 * i + 2];
        const double d = a * a + b * b + c * c;
        if (d > best) best = d;
      }
      if (best > 0.0) n.radius = sqrt(best);
    } else {
      n.radius = 1.0;
    }
  }
  for (int i = first; i <= last; ++i) node_add_particle(n, mass[i], soft[i], pos + 3 * i);
  if (count > 1) {
    double best = 0.0;
    for (int i = first; i <= last; ++i) {
      const double a = n.cm[0] - pos[3 * i], b = n.cm[1] - pos[3 * i + 1], c = n.cm[2] - pos[3 * i + 2];
      const double d = a * a + b * b + c * c;
      if (d > best) best = d;
    }
    if (best > 0.0) {
      best = sqrt(best);
      fm_rescale(n.f, best, n.radius);
      n.radius = best;
    }
  }
}

/* calculateRadiusFarthestCorner, MultipoleMoments.h:460-471 */
__host__ __device__ inline void node_radius_from_box(MomentNode &n, const double *lo, const double *hi) {
  double s = 0.0;
  for (int d = 0; d < 3; ++d) {
    const double a = n.cm[d] - lo[d], b = hi[d] - n.cm[d];
    const double w = a > b ? a : b;
    s += w * w;
  }
  const double r = sqrt(s);
  fm_rescale(n.f, r, n.radius);
  n.radius = r;
}

/* MomentNode -> 27 values in CudaMultipoleMoments order (cuda_typedef.h:137-175) */
template <typename T>
__host__ __device__ inline void node_export(const MomentNode &n, T *o) {
  o[0] = (T)n.radius; o[1] = (T)n.soft; o[2] = (T)n.mass;
  o[3] = (T)n.cm[0]; o[4] = (T)n.cm[1]; o[5] = (T)n.cm[2];
  o[6] = (T)n.f[F_XX]; o[7] = (T)n.f[F_XY]; o[8] = (T)n.f[F_XZ]; o[9] = (T)n.f[F_YY]; o[10] = (T)n.f[F_YZ];
  o[11] = (T)n.f[F_XXX]; o[12] = (T)n.f[F_XYY]; o[13] = (T)n.f[F_XXY]; o[14] = (T)n.f[F_YYY];
  o[15] = (T)n.f[F_XXZ]; o[16] = (T)n.f[F_YYZ]; o[17] = (T)n.f[F_XYZ];
  o[18] = (T)n.f[F_XXXX]; o[19] = (T)n.f[F_XYYY]; o[20] = (T)n.f[F_XXXY]; o[21] = (T)n.f[F_YYYY];
  o[22] = (T)n.f[F_XXXZ]; o[23] = (T)n.f[F_YYYZ]; o[24] = (T)n.f[F_XXYY]; o[25] = (T)n.f[F_XXYZ];
  o[26] = (T)n.f[F_XYYZ];
}

#ifdef __CUDACC__
/* CudaMultipoleMoments record (27 reals) -> the 32 slots of a PackedCell: what repack_cells_kernel
 * (gravity_kernels.cuh) does at upload time, on the values already rounded to `real` */
__device__ __forceinline__ void pack_cell_slots(const real *m, real *v) {
  const real xx = 3 * m[6], xy = 3 * m[7], xz = 3 * m[8], yy = 3 * m[9], yz = 3 * m[10];
  const real xxx = 15 * m[11], xyy = 15 * m[12], xxy = 15 * m[13], yyy = 15 * m[14], xxz = 15 * m[15],
             yyz = 15 * m[16], xyz = 15 * m[17];
  const real xxxx = 105 * m[18], xyyy = 105 * m[19], xxxy = 105 * m[20], yyyy = 105 * m[21],
             xxxz = 105 * m[22], yyyz = 105 * m[23], xxyy = 105 * m[24], xxyz = 105 * m[25], xyyz = 105 * m[26];
  v[PK_CX] = m[3]; v[PK_CY] = m[4]; v[PK_CZ] = m[5]; v[PK_RADIUS] = m[0];
  v[PK_MASS] = m[2]; v[PK_XX] = xx; v[PK_XY] = xy; v[PK_XZ] = xz;
  v[PK_YY] = yy; v[PK_YZ] = yz; v[PK_ZZ] = -(xx + yy); v[PK_XXX] = xxx;
  v[PK_XYY] = xyy; v[PK_XXY] = xxy; v[PK_YYY] = yyy; v[PK_XXZ] = xxz;
  v[PK_YYZ] = yyz; v[PK_XYZ] = xyz; v[PK_XZZ] = -(xxx + xyy); v[PK_YZZ] = -(xxy + yyy);
  v[PK_XXXX] = xxxx; v[PK_XYYY] = xyyy; v[PK_XXXY] = xxxy; v[PK_YYYY] = yyyy;
  v[PK_XXXZ] = xxxz; v[PK_YYYZ] = yyyz; v[PK_XXYY] = xxyy; v[PK_XXYZ] = xxyz;
  v[PK_XYYZ] = xyyz; v[PK_XY3S] = xyyy + xxxy; v[PK_SOFT] = m[1]; v[PK_PAD] = 0;
}

/* one tree level: nodes [lo, lo+n).  Children live on deeper levels and are
 * already in `work`.  Outputs (each optional): `out` CudaMultipoleMoments records in `real`
 * (what the upload path of the force kernels consumes), `out64` the same 27 values in double
 * (the walk, the Ewald set-up), `packed` the PackedCell rows of the force kernels directly.
 *
 * The records of a block's 64 nodes are consecutive in every output array, so they are laid out
 * in shared memory and written by the whole block with consecutive stores: a thread writing its
 * own 27-value record is 27 stores of 32 different sectors each -- the first version spent most of
 * its time in those (10 GB of L2 write transactions for 1.3 GB of records at 256^3). */
constexpr int kMomThreads = 64;
template <int MINB>
__global__ void __launch_bounds__(kMomThreads, MINB)
build_moments_level_kernel(const double *__restrict__ pos, const double *__restrict__ mass,
                           const double *__restrict__ soft, const int *__restrict__ child0,
                           const int *__restrict__ child1, const int *__restrict__ firstPart,
                           const int *__restrict__ lastPart, const double *__restrict__ geolo,
                           const double *__restrict__ geohi, const double *__restrict__ boxlo,
                           const double *__restrict__ boxhi, int lo, int n, int numNodes,
                           MomentNode *__restrict__ work, real *__restrict__ out,
                           double *__restrict__ out64, PackedCell *__restrict__ packed,
                           const unsigned char *__restrict__ flag, int mode) {
  /* flag (or NULL = every node): the nodes of this level that belong to the rank's locally essential tree
   * (force_step.cuh); the others are left alone.  mode 0: build and export; 1: build only (the level whose
   * records are exchanged between the ranks before they are exported); 2: export the records already in `work` */
  __shared__ __align__(16) double stage[kMomThreads * 32]; /* 27 doubles per record; 32 reals per packed row */
  const int base = blockIdx.x * kMomThreads;
  const int t = base + threadIdx.x;
  const int valid = min(kMomThreads, n - base);
  const bool mine = t < n && (!flag || flag[lo + t]);
  if (!__syncthreads_or(mine)) return;
  /* the per-node work records are kept component-major (work[k * numNodes + node]): neighbouring
   * threads own neighbouring nodes and their children are neighbours one level down, so every
   * load and store of a component is a run of consecutive doubles instead of a 224-byte stride */
  double *w = reinterpret_cast<double *>(work);
  constexpr int kWords = (int)(sizeof(MomentNode) / sizeof(double));
  MomentNode m;
  if (mine && mode == 2) {
    double *pm = reinterpret_cast<double *>(&m);
#pragma unroll
    for (int k = 0; k < kWords; ++k) pm[k] = w[(size_t)k * numNodes + lo + t];
    node_export(m, stage + threadIdx.x * 27);
  } else if (mine) {
    const int i = lo + t;
    const int c0 = child0[i], c1 = child1[i];
    if (c0 < 0 && c1 < 0) {
      node_make_bucket(m, pos, mass, soft, firstPart[i], lastPart[i], geolo + 3 * i, geohi + 3 * i);
    } else {
      node_clear(m);
      MomentNode o;
      double *po = reinterpret_cast<double *>(&o);
      if (c0 >= 0) {
#pragma unroll
        for (int k = 0; k < kWords; ++k) po[k] = w[(size_t)k * numNodes + c0];
        node_add_node(m, o);
      }
      if (c1 >= 0) {
#pragma unroll
        for (int k = 0; k < kWords; ++k) po[k] = w[(size_t)k * numNodes + c1];
        node_add_node(m, o);
      }
      node_radius_from_box(m, boxlo + 3 * i, boxhi + 3 * i);
    }
    {
      const double *pm = reinterpret_cast<const double *>(&m);
#pragma unroll
      for (int k = 0; k < kWords; ++k) w[(size_t)k * numNodes + i] = pm[k];
    }
    if (mode != 1) node_export(m, stage + threadIdx.x * 27);
  }
  if (mode == 1) return;
  __syncthreads();
  const size_t first = (size_t)(lo + base);
  if (out64)
    for (int k = threadIdx.x; k < valid * 27; k += kMomThreads) out64[first * 27 + k] = stage[k];
  if (out)
    for (int k = threadIdx.x; k < valid * 27; k += kMomThreads) out[first * 27 + k] = (real)stage[k];
  if (packed) {
    /* rows in `real`, through the same staging area (the double records are not needed any more) */
    real rec[27];
    if (t < n) {
#pragma unroll
      for (int k = 0; k < 27; ++k) rec[k] = (real)stage[threadIdx.x * 27 + k];
    }
    __syncthreads();
    real *rows = reinterpret_cast<real *>(stage);
    static_assert(kCellReals * sizeof(real) <= 32 * sizeof(double), "a packed row fits in a record's staging slot");
    if (t < n) pack_cell_slots(rec, rows + threadIdx.x * kCellReals);
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(packed + first);
    const uint4 *src = reinterpret_cast<const uint4 *>(rows);
    for (int k = threadIdx.x; k < valid * kCellPieces; k += kMomThreads) dst[k] = src[k];
  }
}
#endif

}  // namespace cb200
#endif
