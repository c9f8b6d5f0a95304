/* ewald_setup.cuh -- the per-step Ewald set-up (TreePiece::EwaldInit, Ewald.cpp:285-375), written
 * once for the host (csrc/treewalk.cpp: cb200h_ewald_tables) and for the device (hostcuda.cu:
 * ewald_setup_kernel, which lets a force step build its h-table without bringing the root moments
 * back to the host).
 *
 * With T2, T3, T4 the COMPLETE (not trace-free-reduced) moment tensors of the root cell and h an
 * integer wave vector, the reference's QEVAL (Ewald.cpp:9-46) reduces for the h-loop factors to
 *   hCfac = -(g0 M + g2 T2:hh/2 + g4 T4::hhhh/24),   hSfac = -(g3 T3:.hhh/6),
 *   g_k = g0 (2 pi/L)^k with signs (+,+,-,-,+,+),    g0 = exp(-pi^2 |h|^2 / (alpha^2 L^2)) / (pi |h|^2 L).
 * A component is addressed by its index counts: key = nx*25 + ny*5 + nz. */
#ifndef CB200_EWALD_SETUP_CUH
#define CB200_EWALD_SETUP_CUH

#include <math.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace cb200 {

constexpr int ewald_key(int nx, int ny, int nz) { return nx * 25 + ny * 5 + nz; }

/* root: the 27 doubles of the root cell {radius, soft, mass, cm[3], 21 reduced components scaled by
 * radius^order (FMOMR, moments.c:238-267)}.  comp[125]: every complete component by key; momc[32]:
 * MomcData order (EwaldCUDA.h:29-37). */
__host__ __device__ inline void ewald_complete_moments(const double *root, double *comp, double *momc) {
  /* stored order of CudaMultipoleMoments (cuda_typedef.h:125-127) */
  const int k2[5] = {ewald_key(2, 0, 0), ewald_key(1, 1, 0), ewald_key(1, 0, 1), ewald_key(0, 2, 0), ewald_key(0, 1, 1)};
  const int k3[7] = {ewald_key(3, 0, 0), ewald_key(1, 2, 0), ewald_key(2, 1, 0), ewald_key(0, 3, 0),
                     ewald_key(2, 0, 1), ewald_key(0, 2, 1), ewald_key(1, 1, 1)};
  const int k4[9] = {ewald_key(4, 0, 0), ewald_key(1, 3, 0), ewald_key(3, 1, 0), ewald_key(0, 4, 0), ewald_key(3, 0, 1),
                     ewald_key(0, 3, 1), ewald_key(2, 2, 0), ewald_key(2, 1, 1), ewald_key(1, 2, 1)};
  for (int i = 0; i < 125; ++i) comp[i] = 0.0;
  const double r = root[0];
  for (int i = 0; i < 5; ++i) comp[k2[i]] = root[6 + i] * r * r;
  for (int i = 0; i < 7; ++i) comp[k3[i]] = root[11 + i] * r * r * r;
  for (int i = 0; i < 9; ++i) comp[k4[i]] = root[18 + i] * r * r * r * r;
  /* trace-free completion: T[..zz] = -(T[..xx] + T[..yy]), fewest z first */
  for (int order = 2; order <= 4; ++order)
    for (int nz = 2; nz <= order; ++nz)
      for (int nx = 0; nx <= order - nz; ++nx) {
        const int ny = order - nz - nx;
        comp[ewald_key(nx, ny, nz)] = -(comp[ewald_key(nx + 2, ny, nz - 2)] + comp[ewald_key(nx, ny + 2, nz - 2)]);
      }
  /* m; xx,yy,xy,xz,yz; xxx,xyy,xxy,yyy,xxz,yyz,xyz; xxxx,xyyy,xxxy,yyyy,xxxz,yyyz,xxyy,xxyz,xyyz; zz; xzz,yzz,zzz;
   * xxzz,xyzz,xzzz,yyzz,yzzz,zzzz */
  const int km[31] = {ewald_key(2, 0, 0), ewald_key(0, 2, 0), ewald_key(1, 1, 0), ewald_key(1, 0, 1), ewald_key(0, 1, 1),
                      ewald_key(3, 0, 0), ewald_key(1, 2, 0), ewald_key(2, 1, 0), ewald_key(0, 3, 0), ewald_key(2, 0, 1),
                      ewald_key(0, 2, 1), ewald_key(1, 1, 1),
                      ewald_key(4, 0, 0), ewald_key(1, 3, 0), ewald_key(3, 1, 0), ewald_key(0, 4, 0), ewald_key(3, 0, 1),
                      ewald_key(0, 3, 1), ewald_key(2, 2, 0), ewald_key(2, 1, 1), ewald_key(1, 2, 1),
                      ewald_key(0, 0, 2), ewald_key(1, 0, 2), ewald_key(0, 1, 2), ewald_key(0, 0, 3),
                      ewald_key(2, 0, 2), ewald_key(1, 1, 2), ewald_key(1, 0, 3), ewald_key(0, 2, 2), ewald_key(0, 1, 3),
                      ewald_key(0, 0, 4)};
  momc[0] = root[2];
  for (int i = 0; i < 31; ++i) momc[1 + i] = comp[km[i]];
}

/* one row {hx, hy, hz, hCfac, hSfac} of the h-loop table for the integer wave vector (hx, hy, hz) */
__host__ __device__ inline void ewald_h_row(const double *comp, double M, double L, int hx, int hy, int hz, double *row) {
  const double kPi = 3.14159265358979323846;
  const double alpha = 2.0 / L, k4 = kPi * kPi / (alpha * alpha * L * L), c = 2.0 * kPi / L;
  const int h2 = hx * hx + hy * hy + hz * hz;
  const double h[3] = {(double)hx, (double)hy, (double)hz};
  double q2 = 0, q3 = 0, q4 = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      int cnt2[3] = {0, 0, 0};
      cnt2[a]++; cnt2[b]++;
      q2 += comp[ewald_key(cnt2[0], cnt2[1], cnt2[2])] * h[a] * h[b];
      for (int d = 0; d < 3; ++d) {
        int cnt3[3] = {cnt2[0], cnt2[1], cnt2[2]};
        cnt3[d]++;
        q3 += comp[ewald_key(cnt3[0], cnt3[1], cnt3[2])] * h[a] * h[b] * h[d];
        for (int e = 0; e < 3; ++e) {
          int cnt4[3] = {cnt3[0], cnt3[1], cnt3[2]};
          cnt4[e]++;
          q4 += comp[ewald_key(cnt4[0], cnt4[1], cnt4[2])] * h[a] * h[b] * h[d] * h[e];
        }
      }
    }
  const double g0 = exp(-k4 * h2) / (kPi * h2 * L);
  const double g2 = -c * c * g0, g3 = -c * c * c * g0, g4 = c * c * c * c * g0;
  row[0] = c * hx; row[1] = c * hy; row[2] = c * hz;
  row[3] = -(g0 * M + g2 * q2 / 2.0 + g4 * q4 / 24.0);
  row[4] = -(g3 * q3 / 6.0);
}

/* the h-vectors of the table in the reference's loop order (hx outermost, Ewald.cpp:304-375);
 * returns their number, writes up to cap triples */
__host__ __device__ inline int ewald_h_vectors(double dEwhCut, int *hxyz, int cap) {
  const int hreps = (int)ceil(dEwhCut);
  int n = 0;
  for (int hx = -hreps; hx <= hreps; ++hx)
    for (int hy = -hreps; hy <= hreps; ++hy)
      for (int hz = -hreps; hz <= hreps; ++hz) {
        const int h2 = hx * hx + hy * hy + hz * hz;
        if (h2 == 0 || h2 > dEwhCut * dEwhCut) continue;
        if (n < cap) { hxyz[3 * n] = hx; hxyz[3 * n + 1] = hy; hxyz[3 * n + 2] = hz; }
        ++n;
      }
  return n;
}

}  // namespace cb200
#endif
