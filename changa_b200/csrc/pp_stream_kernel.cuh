/* pp_stream_kernel.cuh -- particle-particle lists, second design (float build).
 *
 * Replaces particleGravityComputation (HostCUDA.cu:1565-1751) like part_list_x2_kernel does, with
 * the same arithmetic (SPLINE of gravity.h:147-182; r = shift + source - target, HostCUDA.cu:1655-1663)
 * and the same decomposition (one warp owns one bucket, lane = source, two targets per packed
 * f32x2 instruction), but with everything AROUND the pair evaluations rebuilt after the ncu
 * capture of round 1 (ALU pipe 49.5 % vs FMA pipe 27.8 %, 1.44 long-scoreboard stalls per issue,
 * spills at 128 registers):
 *
 *   - a lane owns TWO list entries per iteration (a 64-entry chunk): the per-chunk work (list
 *     load, gather, replica decode, waits, loop) is paid once per 64 sources, and one
 *     shared-memory read of a target pair serves two bodies;
 *   - the 32-byte source rows are gathered with cp.async (16 + 4 bytes per entry: position+mass
 *     and softening) into a per-warp two-stage ring one chunk ahead; a lane copies and later
 *     reads only its OWN rows, so the chunk needs no __syncwarp and no register is held while the
 *     rows are in flight;
 *   - the two bodies that share a target pair (sources A and B of the lane) sit in one basic block
 *     and interleave; pairs are guarded by a warp-uniform test (one copy of the code: see the note at
 *     the chunk loop about the instruction cache);
 *   - the Newtonian body has no select on coincidence: a half inside the softening sphere OR at
 *     zero distance reads rsqrt(+inf) = 0 (one FSETP + one FSEL per half), and the rare lanes
 *     that met one redo exactly those halves with the scalar spline afterwards;
 *   - list entries without a source (the tail of the last chunk) are a massless source 1e18 away:
 *     contributes exact zeros, no divergent region; a last chunk of <= 32 entries skips the
 *     second source warp-uniformly;
 *   - the replica shift is decoded only when some entry of the chunk is not in the home box.
 *
 * FMA-pipe budget of a body (two pairs): 17 packed instructions = 34 pipe cycles per SM
 * sub-partition; ALU pipe: 2 FSETP + 2 FSEL + 2 FMNMX + 1 PLOP3; 2 MUFU.RSQ; 1.5 LDS.
 */
#ifndef CB200_PP_STREAM_KERNEL_CUH
#define CB200_PP_STREAM_KERNEL_CUH

#include <type_traits>
#include "gravity_kernels.cuh"

#ifndef CUDA_USE_DOUBLE
namespace cb200 {

constexpr int kPpChunk = 64;
constexpr int kHomeBox = 0xDB; /* ((3) | (3 << 3) | (3 << 6)): replica (0,0,0) in bits 22..30 */

template <int PB>
struct PartWarpSmem {
  static constexpr int stageBytes = kPpChunk * 16 + kPpChunk * 4; /* {x,y,z,m}[64] | soft[64] */
  static constexpr int ring = 0;                                  /* two stages */
  static constexpr int preList = ring + 2 * stageBytes;           /* first 128 list entries of the NEXT bucket */
  static constexpr int preTargets = preList + 2 * kPpChunk * 8;   /* its target rows (PackedPart) */
  static constexpr int targets = preTargets + PB * 32;            /* PB/2 TargetSoftPair */
  static constexpr int red = targets + (PB / 2) * 48;             /* reduction scratch: 5*PB rows of 33 floats */
  static constexpr int bytes = (red + 5 * PB * kRedPitch + 15) & ~15;
};
template <int PB>
constexpr size_t part_list_stream_smem_bytes() {
  return (size_t)kListWarps * PartWarpSmem<PB>::bytes;
}

struct SrcReg { float x, y, z, m, soft; };

__device__ __forceinline__ void cp_async4_s(unsigned saddr, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(gmem) : "memory");
}

/* "near" test shared by the fast body and the spline fix-up: r^2 < (soft_s + soft_t)^2 + tiny.  The
 * tiny term (folded into the packed FMA that squares twoh) sends r = 0 to the fix-up, which skips it
 * (HostCUDA.cu:1665), without a second compare per half. */
constexpr float kPpTiny = 1e-37f;

/* Newtonian part of (one source) x (two targets).  A half inside soft_s + soft_t, or at zero distance
 * (the self pair every local bucket list holds; skipped, HostCUDA.cu:1665), contributes nothing here;
 * `slow` becomes 1 when some half still needs the spline. */
__device__ __forceinline__ void pp_body(const SrcReg &s, const TargetSoftPair &p, f32x2 &ax, f32x2 &ay,
                                        f32x2 &az, f32x2 &pot, float &idt0, float &idt1, unsigned &slow) {
  const f32x2 rx = sub2(bc2(s.x), p.x), ry = sub2(bc2(s.y), p.y), rz = sub2(bc2(s.z), p.z);
  const f32x2 rsq = fma2(rz, rz, fma2(ry, ry, mul2(rx, rx)));
  const f32x2 twoh = add2(p.soft, bc2(s.soft));
  float q0, q1, h0, h1;
  unpk2(rsq, q0, q1);
  unpk2(fma2(twoh, twoh, bc2(kPpTiny)), h0, h1);
  /* e = (q >= h) ? q : +inf, and the fix-up flag (q < h && q != 0) for either half: 4 FSETP, 2 FSEL,
   * one predicate OR, one predicated move -- written in PTX so the flag never becomes integer logic */
  float e0, e1;
  asm("{\n\t.reg .pred f0, f1, s0, s1;\n\t"
      "setp.ge.f32 f0, %3, %5;\n\t"
      "setp.ge.f32 f1, %4, %6;\n\t"
      "selp.f32 %0, %3, 0f7F800000, f0;\n\t"
      "selp.f32 %1, %4, 0f7F800000, f1;\n\t"
      "setp.neu.and.f32 s0, %3, 0f00000000, !f0;\n\t"
      "setp.neu.and.f32 s1, %4, 0f00000000, !f1;\n\t"
      "or.pred s0, s0, s1;\n\t"
      "@s0 mov.b32 %2, 1;\n\t}"
      : "=f"(e0), "=f"(e1), "+r"(slow)
      : "f"(q0), "f"(q1), "f"(h0), "f"(h1));
  const float d0 = rsqrt_dev(e0), d1 = rsqrt_dev(e1);
  const f32x2 d = pk2(d0, d1);
  const f32x2 d2 = mul2(d, d);
  const f32x2 dm = mul2s(s.m, d);
  const f32x2 b = mul2(d2, d);
  const f32x2 bm = mul2(d2, dm);
  ax = fma2(rx, bm, ax);
  ay = fma2(ry, bm, ay);
  az = fma2(rz, bm, az);
  pot = sub2(pot, dm);
  float i0, i1;
  unpk2(fma2(p.m, b, bm), i0, i1); /* (m_t + m_s) b */
  idt0 = fmaxf(idt0, i0);
  idt1 = fmaxf(idt1, i1);
}

/* The halves pp_body left out, with the scalar spline (gravity.h:147-182; same near test, bit for
 * bit).  Rare, so it is kept out of the unrolled bodies: one routine, a plain loop over the bucket's
 * target pairs, adding into THIS lane's column of the bucket's reduction scratch (rows = particle*5 +
 * component, zeroed by the caller on first use), which the final reduction sums anyway. */
__device__ __noinline__ void pp_near_lane(float sx, float sy, float sz, float sm, float ssoft, int npairs,
                                          unsigned tgtAddr, unsigned redCol) {
  for (int j = 0; j < npairs; ++j) {
    const TargetSoftPair p = lds_target_soft_pair(tgtAddr + j * 48u);
    float x[2], y[2], z[2], m[2], h[2];
    unpk2(p.x, x[0], x[1]); unpk2(p.y, y[0], y[1]); unpk2(p.z, z[0], z[1]); unpk2(p.m, m[0], m[1]);
    unpk2(p.soft, h[0], h[1]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float rx = sx - x[e], ry = sy - y[e], rz = sz - z[e];
      const float q = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
      const float t = h[e] + ssoft;
      if (!(q >= fmaf(t, t, kPpTiny)) && q != 0.0f) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
        const real4 tp = {x[e], y[e], z[e], m[e]};
        pp_pair(sx, sy, sz, sm, ssoft, tp, h[e], a0, a1, a2, a3, a4);
        const unsigned r0 = redCol + (unsigned)((2 * j + e) * 5) * kRedPitch;
        float *c = reinterpret_cast<float *>(__cvta_shared_to_generic(r0));
        constexpr int pitch = kRedPitch / 4;
        c[0] += a0; c[pitch] += a1; c[2 * pitch] += a2; c[3 * pitch] += a3;
        c[4 * pitch] = fmaxf(c[4 * pitch], a4);
      }
    }
  }
}

template <int PB, int MINB>
__global__ void __launch_bounds__(kListWarps * 32, MINB)
part_list_stream_kernel(const PackedPart *__restrict__ parts, VariablePartData *__restrict__ vars,
                        const PackedPart *__restrict__ sources, const ILCell *__restrict__ list,
                        const int *__restrict__ markers, const int *__restrict__ starts,
                        const int *__restrict__ sizes, int nBuckets, float fperiod,
                        unsigned int *__restrict__ nextBucket) {
  static_assert(PB % 2 == 0 && 5 * PB <= 64, "two reduction rows per lane at most");
  constexpr int NP = PB / 2;
  typedef PartWarpSmem<PB> S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wsm = smem_raw + (size_t)warp * S::bytes;
  const unsigned wbase = smem_u32(wsm);
  const unsigned posAddr = pin_u32(wbase + S::ring + lane * 16);            /* my first row; second at +512 */
  const unsigned softAddr = pin_u32(wbase + S::ring + kPpChunk * 16 + lane * 4);
  const unsigned tgtAddr = pin_u32(wbase + S::targets);
  const unsigned redRow = wbase + S::red + lane * kRedPitch;
  const unsigned redCol = wbase + S::red + lane * 4;
  const ILCell none = {-1, kHomeBox << 22};
  const SrcReg nowhere = {1e18f, 1e18f, 1e18f, 0.0f, 0.0f};

  auto grab = [&]() { return grab_bucket(nextBucket, nBuckets, lane); };
  /* rows of my two entries of a chunk -> ring stage `st` */
  auto gather = [&](unsigned st, int ia, int ib) {
    const unsigned o = st * S::stageBytes;
    if (ia >= 0) {
      const PackedPart *q = sources + ia;
      cp_async16_s(posAddr + o, q);
      cp_async4_s(softAddr + o, &q->soft);
    }
    if (ib >= 0) {
      const PackedPart *q = sources + ib;
      cp_async16_s(posAddr + o + 512, q);
      cp_async4_s(softAddr + o + 128, &q->soft);
    }
  };
  auto prefetch_bucket = [&](const BucketMeta &b) {
    const ILCell *nl = list + b.begin;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (32 * i + lane < b.len) cp_async8_s(wbase + S::preList + (32 * i + lane) * 8, nl + 32 * i + lane);
    if (lane < min(PB, b.count)) {
      const PackedPart *q = parts + b.first + lane;
      cp_async16_s(wbase + S::preTargets + lane * 32, q);
      cp_async16_s(wbase + S::preTargets + lane * 32 + 16, reinterpret_cast<const char *>(q) + 16);
    }
  };
  auto load_src = [&](unsigned st, unsigned second, const ILCell &e) {
    SrcReg s = nowhere;
    if (e.index >= 0) {
      const uint4 v = lds128(posAddr + st * S::stageBytes + second * 512);
      s.x = __uint_as_float(v.x); s.y = __uint_as_float(v.y); s.z = __uint_as_float(v.z); s.m = __uint_as_float(v.w);
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s.soft) : "r"(softAddr + st * S::stageBytes + second * 128));
    }
    return s;
  };
  auto shift_src = [&](SrcReg &s, int off) {
    s.x = fmaf(float(replica_x(off)), fperiod, s.x);
    s.y = fmaf(float(replica_y(off)), fperiod, s.y);
    s.z = fmaf(float(replica_z(off)), fperiod, s.z);
  };

  int k = grab();
  BucketMeta m = {0, 0, 0, 0};
  if (k < nBuckets) {
    m = load_bucket_meta(markers, starts, sizes, k);
    prefetch_bucket(m);
  }
  cp_async_commit();

  while (k < nBuckets) {
    const int kn = grab();
    BucketMeta mn = {0, 0, 0, 0};
    if (kn < nBuckets) mn = load_bucket_meta(markers, starts, sizes, kn);
    bool prefetched = false;

    const ILCell *__restrict__ mylist = list + m.begin;
    const int len = m.len, nchunks = (len + kPpChunk - 1) / kPpChunk;

    for (int p0 = 0; p0 < m.count && len > 0; p0 += PB) { /* one pass unless the bucket outgrows PB */
      const int np = min(PB, m.count - p0);
      const int npairs = (np + 1) >> 1;
      const bool lastPass = p0 + PB >= m.count;
      ILCell curA = none, curB = none, nxtA = none, nxtB = none;
      cp_async_wait<0>();
      __syncwarp();
      if (p0 == 0) { /* from the staging area (entries a lane reads are the ones it copied) */
        const ILCell *pl = reinterpret_cast<const ILCell *>(wsm + S::preList);
        if (lane < len) curA = pl[lane];
        if (32 + lane < len) curB = pl[32 + lane];
        if (64 + lane < len) nxtA = pl[64 + lane];
        if (96 + lane < len) nxtB = pl[96 + lane];
      } else {
        if (lane < len) curA = mylist[lane];
        if (32 + lane < len) curB = mylist[32 + lane];
        if (64 + lane < len) nxtA = mylist[64 + lane];
        if (96 + lane < len) nxtB = mylist[96 + lane];
      }
      if (lane < 2 * npairs) { /* an odd bucket's last slot repeats its last particle; that half is never stored */
        const int src = min(lane, np - 1);
        float4 v;
        float vs;
        if (p0 == 0) {
          v = *reinterpret_cast<const float4 *>(wsm + S::preTargets + src * 32);
          vs = *reinterpret_cast<const float *>(wsm + S::preTargets + src * 32 + 16);
        } else {
          const PackedPart *q = parts + m.first + p0 + src;
          v = *reinterpret_cast<const float4 *>(q);
          vs = q->soft;
        }
        float *dst = reinterpret_cast<float *>(wsm + S::targets) + (lane >> 1) * 12 + (lane & 1);
        dst[0] = v.x; dst[2] = v.y; dst[4] = v.z; dst[6] = v.w; dst[8] = vs;
      }
      gather(0, curA.index, curB.index);
      cp_async_commit();
      __syncwarp();

      f32x2 ax[NP], ay[NP], az[NP], pot[NP];
      float idt[PB];
#pragma unroll
      for (int j = 0; j < NP; ++j) { ax[j] = ay[j] = az[j] = pot[j] = 0ull; idt[2 * j] = idt[2 * j + 1] = 0.0f; }
      bool dirty = false; /* the scratch holds spline contributions of this pass */

      /* The chunk loop.  ONE copy of the bodies, guarded per target pair by a warp-uniform test: a first
       * version instantiated the loop per pair count (no guards, 12 copies, 4768 SASS instructions = 76 KB)
       * and ran at 13.6 no-instruction stalls per issue -- the 16 resident warps of an SM sat in different
       * copies and thrashed the instruction cache (profiles/r02b_ncu_part_list_stream_icache.json). */
      for (int c = 0; c < nchunks; ++c) {
        const unsigned st = c & 1;
        if (c + 1 < nchunks) gather(st ^ 1, nxtA.index, nxtB.index);
        if (lastPass && c == nchunks - 1) { /* the next bucket's first loads ride under this chunk */
          prefetch_bucket(mn);
          prefetched = true;
        }
        cp_async_commit();
        ILCell nnA = none, nnB = none;
        const int e2 = (c + 2) * kPpChunk + lane;
        if (e2 < len) nnA = mylist[e2];
        if (e2 + 32 < len) nnB = mylist[e2 + 32];
        cp_async_wait<1>();

        SrcReg sa = load_src(st, 0, curA);
        const bool hasB = c * kPpChunk + 32 < len; /* warp-uniform */
        SrcReg sb = nowhere;
        if (hasB) sb = load_src(st, 1, curB);
        const int away = ((curA.offsetID >> 22) ^ kHomeBox) | ((curB.offsetID >> 22) ^ kHomeBox);
        if (__any_sync(kFull, (away & 0x1ff) != 0)) {
          shift_src(sa, curA.offsetID);
          shift_src(sb, curB.offsetID);
        }
        unsigned slowA = 0, slowB = 0;
        if (hasB) {
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            if (j < npairs) { /* two independent bodies on one read of the target pair */
              const TargetSoftPair p = lds_target_soft_pair(tgtAddr + j * 48u);
              pp_body(sa, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1], slowA);
              pp_body(sb, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1], slowB);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            if (j < npairs) {
              const TargetSoftPair p = lds_target_soft_pair(tgtAddr + j * 48u);
              pp_body(sa, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1], slowA);
            }
          }
        }
        if (__any_sync(kFull, (slowA | slowB) != 0)) { /* rare: a pair inside the softening length */
          if (!dirty) {
            for (int r = 0; r < 10 * npairs; ++r)
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(redCol + r * kRedPitch), "f"(0.0f) : "memory");
            dirty = true;
          }
          if (slowA) pp_near_lane(sa.x, sa.y, sa.z, sa.m, sa.soft, npairs, tgtAddr, redCol);
          if (slowB) pp_near_lane(sb.x, sb.y, sb.z, sb.m, sb.soft, npairs, tgtAddr, redCol);
        }
        curA = nxtA; curB = nxtB;
        nxtA = nnA; nxtB = nnB;
      }

      /* park partial sums: row (particle*5 + component), column lane.  When the spline fix-up wrote into
       * the scratch (rare, warp-uniform) the sums are combined with what it left there. */
      auto lds = [](unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; };
      auto sts = [](unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); };
      if (!dirty) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          if (j < npairs) {
            float a0, a1, b0, b1, c0, c1, e0, e1;
            unpk2(ax[j], a0, a1); unpk2(ay[j], b0, b1); unpk2(az[j], c0, c1); unpk2(pot[j], e0, e1);
            const unsigned r0 = redCol + (2 * j) * 5 * kRedPitch;
            sts(r0, a0); sts(r0 + kRedPitch, b0); sts(r0 + 2 * kRedPitch, c0); sts(r0 + 3 * kRedPitch, e0);
            sts(r0 + 4 * kRedPitch, idt[2 * j]);
            sts(r0 + 5 * kRedPitch, a1); sts(r0 + 6 * kRedPitch, b1); sts(r0 + 7 * kRedPitch, c1); sts(r0 + 8 * kRedPitch, e1);
            sts(r0 + 9 * kRedPitch, idt[2 * j + 1]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          if (j < npairs) {
            float v[10];
            unpk2(ax[j], v[0], v[5]); unpk2(ay[j], v[1], v[6]); unpk2(az[j], v[2], v[7]); unpk2(pot[j], v[3], v[8]);
            v[4] = idt[2 * j]; v[9] = idt[2 * j + 1];
            const unsigned r0 = redCol + (2 * j) * 5 * kRedPitch;
#pragma unroll
            for (int q = 0; q < 10; ++q) {
              const float prev = lds(r0 + q * kRedPitch);
              sts(r0 + q * kRedPitch, (q % 5) == 4 ? fmaxf(prev, v[q]) : prev + v[q]);
            }
          }
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
          const unsigned row = redRow + h * 32 * kRedPitch; /* row v: bank (v + i) % 32 for element i */
          const bool isMax = (v % 5) == 4;
          float acc = 0.0f;
          if (isMax) { /* dtGrav rows */
#pragma unroll
            for (int i = 0; i < 32; ++i) acc = fmaxf(acc, lds(row + i * 4));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += lds(row + i * 4);
          }
          /* accumulate, never overwrite (HostCUDA.cu:1749-1751); dtGrav is a running max */
          out[v] = isMax ? fmaxf(old[h], acc) : old[h] + acc;
        }
      }
      __syncwarp();
    }
    if (!prefetched) prefetch_bucket(mn); /* empty list: nothing rode under a chunk */
    cp_async_commit();
    k = kn; m = mn;
  }
  cp_async_wait<0>();
}

}  // namespace cb200
#endif /* !CUDA_USE_DOUBLE */
#endif
