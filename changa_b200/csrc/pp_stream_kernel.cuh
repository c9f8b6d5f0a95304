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
  static constexpr int red = targets + (PB / 2) * 48;             /* reduction scratch: per target pair 4 packed sum rows, then
                                                                     one packed dtGrav row per pair (spline fix-up only) */
  static constexpr int bytes = (red + 5 * (PB / 2) * kRed2Pitch + 15) & ~15;
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
 * target pairs, adding into THIS lane's 8-byte column of the bucket's reduction scratch (rows 4 * pair +
 * component hold {half 0, half 1}; rows 4 * NP + pair the two dtGrav maxima; zeroed by the caller on
 * first use), which the lane folds into its registers before the reduction. */
__device__ __noinline__ void pp_near_lane(float sx, float sy, float sz, float sm, float ssoft, int npairs,
                                          unsigned tgtAddr, unsigned redCol, int maxRow0) {
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
        const unsigned r0 = redCol + (unsigned)(4 * j) * kRed2Pitch + e * 4;
        float *c = reinterpret_cast<float *>(__cvta_shared_to_generic(r0));
        constexpr int pitch = kRed2Pitch / 4;
        c[0] += a0; c[pitch] += a1; c[2 * pitch] += a2; c[3 * pitch] += a3;
        float *mx = reinterpret_cast<float *>(__cvta_shared_to_generic(redCol + (unsigned)(maxRow0 + j) * kRed2Pitch + e * 4));
        *mx = fmaxf(*mx, a4);
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
  static_assert(PB % 2 == 0 && 2 * PB <= 32, "one packed sum row per lane (bucket_reduce_store)");
  constexpr int NP = PB / 2;
  typedef PartWarpSmem<PB> S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wsm = smem_raw + (size_t)warp * S::bytes;
  const unsigned wbase = smem_u32(wsm);
  const unsigned posAddr = pin_u32(wbase + S::ring + lane * 16);            /* my first row; second at +512 */
  const unsigned softAddr = pin_u32(wbase + S::ring + kPpChunk * 16 + lane * 4);
  const unsigned tgtAddr = pin_u32(wbase + S::targets);
  const unsigned redBase = wbase + S::red;
  const unsigned redCol = redBase + lane * 8;
  const ILCell none = {-1, kHomeBox << 22};
  const SrcReg nowhere = {1e18f, 1e18f, 1e18f, 0.0f, 0.0f};

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
  auto prefetch_bucket = [&](const RawMeta &b) { /* per-lane values: nothing here needs them uniform */
    const ILCell *nl = list + b.begin;
    const int blen = b.end - b.begin;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (32 * i + lane < blen) cp_async8_s(wbase + S::preList + (32 * i + lane) * 8, nl + 32 * i + lane);
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

  /* the bucket pipeline (gravity_kernels.cuh, grab_bucket_raw): k in work, k1 held, k2's atomic in flight */
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
    unsigned k2raw = (unsigned)k1;
    if (deep && k1 < nBuckets) k2raw = grab_bucket_raw(nextBucket, nBuckets, lane);
    RawMeta mn = {0, 0, 0, 0};
    if (k1 < nBuckets) mn = load_bucket_meta_raw(markers, starts, sizes, k1);
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
          /* the next target pair is requested before this one is used, outside the guard: same time as
           * loading inside the guard (5.41 ms at 256^3 either way), but ptxas allocates this form
           * without spills at the 128-register cap */
          TargetSoftPair p = lds_target_soft_pair(tgtAddr);
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            TargetSoftPair pn = p;
            if (j + 1 < NP) pn = lds_target_soft_pair(tgtAddr + (j + 1) * 48u);
            if (j < npairs) { /* two independent bodies on one read of the target pair */
              pp_body(sa, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1], slowA);
              pp_body(sb, p, ax[j], ay[j], az[j], pot[j], idt[2 * j], idt[2 * j + 1], slowB);
            }
            p = pn;
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
            for (int r = 0; r < 4 * npairs; ++r) sts64(redCol + r * kRed2Pitch, 0ull);
            for (int r = 0; r < npairs; ++r) sts64(redCol + (4 * NP + r) * kRed2Pitch, 0ull);
            dirty = true;
          }
          if (slowA) pp_near_lane(sa.x, sa.y, sa.z, sa.m, sa.soft, npairs, tgtAddr, redCol, 4 * NP);
          if (slowB) pp_near_lane(sb.x, sb.y, sb.z, sb.m, sb.soft, npairs, tgtAddr, redCol, 4 * NP);
        }
        curA = nxtA; curB = nxtB;
        nxtA = nnA; nxtB = nnB;
      }

      if (dirty) { /* rare, warp-uniform: fold what the spline fix-up left in my column into my sums */
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          if (j < npairs) {
            const unsigned r0 = redCol + (4 * j) * kRed2Pitch;
            ax[j] = add2(ax[j], lds64(r0)); ay[j] = add2(ay[j], lds64(r0 + kRed2Pitch));
            az[j] = add2(az[j], lds64(r0 + 2 * kRed2Pitch)); pot[j] = add2(pot[j], lds64(r0 + 3 * kRed2Pitch));
            float m0, m1;
            unpk2(lds64(redCol + (4 * NP + j) * kRed2Pitch), m0, m1);
            idt[2 * j] = fmaxf(idt[2 * j], m0); idt[2 * j + 1] = fmaxf(idt[2 * j + 1], m1);
          }
        }
        __syncwarp();
      }
      bucket_reduce_store<NP>(ax, ay, az, pot, idt, np, npairs, redBase, lane,
                              reinterpret_cast<float *>(vars + m.first + p0));
    }
    if (!prefetched) { /* empty list: nothing rode under a chunk, and nothing waited for this bucket's own prefetch */
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

}  // namespace cb200
#endif /* !CUDA_USE_DOUBLE */
#endif
