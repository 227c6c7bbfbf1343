// kernels_arb.cu -- fused generator res-block convolution (SURVEY A.9 AdaINResBlock1, K6/K7/K8):
//
//   y[t, co] = bias[co] + sum_tap sum_c  snake( x[t + tap*dil - pad, c] * scale[b,c] + shift[b,c] ) * W[co, tap, c]
//
// as ONE persistent tcgen05 kernel per conv.  Compared with the generic conv_tc path (colstats ->
// adain_coef -> apply_bf16 -> conv_tc) it removes three HBM passes and most of the L2->SM operand
// traffic:
//   * the AdaIN scale/shift + Snake operand transform runs inside the kernel: 8 producer warps read the
//     raw activations (fp32 residual stream or bf16 intermediate) once per tile -- a halo tile of
//     128*MSUB + 2*pad rows per 64-channel chunk -- and write the bf16 operand straight into the
//     128B-swizzled smem layout the tensor core reads;
//   * every tap of the conv reads a ROW-SHIFTED VIEW of that one halo tile (UMMA descriptor start
//     address + tap*dil*128 B), so A is fetched once per tile instead of once per tap, and each weight
//     tile (TMA, 4/3-stage ring) is shared by MSUB 128-row sub-tiles;
//   * the epilogue (TMEM -> registers -> smem transpose -> global) adds bias / residual, writes fp32
//     and/or bf16, and emits the per-128-row column sums (sum x, sum x^2) the next AdaIN needs, so no
//     separate statistics pass reads the tensor again.
// TMEM holds two accumulator sets (2 x MSUB x BN = 512 columns): the epilogue of tile i overlaps the
// MMAs of tile i+1; the operand producers run one channel chunk ahead of the MMA warp.
//
// Warp roles (448 threads, 1 CTA/SM, persistent over tiles blockIdx.x + i*gridDim.x):
//   warp 0      weight-tile TMA producer          warp 1      TMEM alloc + tcgen05.mma issuer
//   warps 2-5   epilogue                          warps 6-13  operand producers (transform)
#include "kernels.h"
#include "tc_ptx.cuh"

namespace kkx {

namespace {

constexpr int kArbThreads = 448;
constexpr int kArbMaxB = 1024;

template <int BN, int MSUB>
struct ArbCfg {
  static constexpr int KCH = BN / 64;                       // 64-channel chunks
  static constexpr int RA = MSUB * 128 + 56;                // rows per A slot (halo <= 2*25, 8-row granule)
  static constexpr uint32_t A_SLOT = RA * 128;              // bytes (multiple of 1024)
  static constexpr int NA = (BN == 128) ? 2 : 3;            // A slots
  static constexpr uint32_t B_STAGE = BN * 128;             // bytes
  static constexpr int NB = (BN == 128) ? 4 : 3;            // B stages
  static constexpr int PITCH = 36;                          // floats per staged epilogue row
  static constexpr uint32_t STG = 2 * 128 * PITCH * 4;      // two transpose buffers
  static constexpr uint32_t STAT = 4 * BN * 2 * 4;          // per-warp column sums
  static constexpr int NBAR = 2 * NA + 2 * NB + 4;
  static constexpr uint32_t SMEM = NA * A_SLOT + NB * B_STAGE + STG + STAT + NBAR * 8 + 16 + (kArbMaxB + 1) * 4 + 1024;
  static_assert(A_SLOT % 1024 == 0, "A slot must keep the 1024-byte swizzle alignment");
};

__device__ __forceinline__ uint64_t umma_desc_sw128_bo(uint32_t saddr, int mode) {
  uint64_t d = umma_desc_sw128(saddr);
  if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7u) << 49;   // matrix base offset = row phase inside the 1024 B swizzle pattern
  return d;
}

__device__ __forceinline__ float snake_f(float v, float al, float ial) {
  const float s = __sinf(al * v);
  return fmaf(s * s, ial, v);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ void tile_decode(const int* ts, int B, int tile, int& b, int& mt) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ts[mid] <= tile) lo = mid; else hi = mid;
  }
  b = lo; mt = tile - ts[lo];
}

template <int BN, int MSUB, bool IN_BF16>
__global__ void __launch_bounds__(kArbThreads, 1) arb_conv_kernel(const __grid_constant__ CUtensorMap tmB, ArbConvArgs a) {
  using Cfg = ArbCfg<BN, MSUB>;
  constexpr int KCH = Cfg::KCH, NA = Cfg::NA, NB = Cfg::NB, PITCH = Cfg::PITCH;
  constexpr int MT = MSUB * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + NA * Cfg::A_SLOT;
  const uint32_t stg_base = b_base + NB * Cfg::B_STAGE;
  const uint32_t stat_base = stg_base + Cfg::STG;
  const uint32_t bar_base = stat_base + Cfg::STAT;
  const uint32_t tmem_slot = bar_base + Cfg::NBAR * 8;
  const uint32_t ts_base = tmem_slot + 16;
  auto fullA = [&](int s) { return bar_base + s * 8; };
  auto emptyA = [&](int s) { return bar_base + (NA + s) * 8; };
  auto fullB = [&](int s) { return bar_base + (2 * NA + s) * 8; };
  auto emptyB = [&](int s) { return bar_base + (2 * NA + NB + s) * 8; };
  auto tfull = [&](int j) { return bar_base + (2 * NA + 2 * NB + j) * 8; };
  auto tempty = [&](int j) { return bar_base + (2 * NA + 2 * NB + 2 + j) * 8; };
  int* const s_ts = reinterpret_cast<int*>(gbase + (ts_base - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int B = a.B;
  for (int i = threadIdx.x; i <= B; i += kArbThreads) s_ts[i] = a.tile_start[i];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < NA; s++) { mbar_init(fullA(s), 8); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < NB; s++) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int j = 0; j < 2; j++) { mbar_init(tfull(j), 1); mbar_init(tempty(j), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int ntiles = a.total_tiles;
  const int ks = a.ks, dil = a.dil, pad = a.pad;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight tiles (TMA)
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < KCH; c++)
          for (int tap = 0; tap < ks; tap++, g++) {
            const int s = g % NB;
            mbar_wait(emptyB(s), (((uint32_t)(g / NB)) & 1u) ^ 1u);
            mbar_expect_tx(fullB(s), Cfg::B_STAGE);
            tma_load_2d(b_base + s * Cfg::B_STAGE, &tmB, tap * BN + c * 64, 0, fullB(s));
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issue
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int gA = 0, gB = 0, ti = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
        int b, mt;
        tile_decode(s_ts, B, tile, b, mt);
        const int rem = a.len[b] - mt * MT;                 // rows of this item left from the tile start
        const int buf = ti & 1;
        mbar_wait(tempty(buf), (((uint32_t)(ti >> 1)) & 1u) ^ 1u);
        tc_fence_after();
        for (int c = 0; c < KCH; c++, gA++) {
          const int sa = gA % NA;
          mbar_wait(fullA(sa), ((uint32_t)(gA / NA)) & 1u);
          tc_fence_after();
          const uint32_t slot = a_base + sa * Cfg::A_SLOT;
          for (int tap = 0; tap < ks; tap++, gB++) {
            const int sb = gB % NB;
            mbar_wait(fullB(sb), ((uint32_t)(gB / NB)) & 1u);
            tc_fence_after();
            const uint64_t bd = umma_desc_sw128(b_base + sb * Cfg::B_STAGE);
#pragma unroll
            for (int sub = 0; sub < MSUB; sub++) {
              if (sub * 128 >= rem) continue;
              const uint64_t ad = umma_desc_sw128_bo(slot + (uint32_t)(sub * 128 + tap * dil) * 128u, a.desc_mode);
              const uint32_t td = tmem_base + (uint32_t)((buf * MSUB + sub) * BN);
#pragma unroll
              for (int k = 0; k < 4; k++)
                umma_bf16(td, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (c | tap | k) ? 1u : 0u);
            }
            umma_commit(emptyB(sb));
          }
          umma_commit(emptyA(sa));
        }
        umma_commit(tfull(buf));
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int et = q * 32 + lane;           // accumulator row held by this thread
    const int t = threadIdx.x - 64;         // 0..127
    const int c4 = (t & 7) << 2;
    float* const stage_f = reinterpret_cast<float*>(gbase + (stg_base - base));
    float* const stat_f = reinterpret_cast<float*>(gbase + (stat_base - base));   // [4][BN][2]
    int chunk_ctr = 0, ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ti++) {
      int b, mt;
      tile_decode(s_ts, B, tile, b, mt);
      const int L = a.len[b], off = a.off[b];
      const int m0 = mt * MT;
      const int nsub = min(MSUB, (L - m0 + 127) >> 7);
      const int buf = ti & 1;
      auto fetch = [&](int sub, int c, float4* rv) {   // residual operand of (sub, 32-col chunk c)
        const int n = c + c4;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int mm = m0 + sub * 128 + (t >> 3) + 16 * i;
          rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.res && mm < L) rv[i] = *reinterpret_cast<const float4*>(a.res + (size_t)(off + mm) * BN + n);
        }
      };
      float4 rv[8];
      fetch(0, 0, rv);
      mbar_wait(tfull(buf), ((uint32_t)(ti >> 1)) & 1u);
      tc_fence_after();
      for (int sub = 0; sub < nsub; sub++) {
        const int sm0 = m0 + sub * 128;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * MSUB + sub) * BN + c), v);
          if (sub == nsub - 1 && c + 32 >= BN) {   // last TMEM read of this tile: hand the accumulators back
            tc_fence_before();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty(buf)) : "memory");
          }
          float* buf_f = stage_f + (chunk_ctr & 1) * (128 * PITCH);
          chunk_ctr++;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(buf_f + et * PITCH + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int n = c + c4;
          const float4 bb = *reinterpret_cast<const float4*>(a.bias + n);
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int row = (t >> 3) + 16 * i;
            const int mm = sm0 + row;
            if (mm >= L) continue;
            float4 o = *reinterpret_cast<const float4*>(buf_f + row * PITCH + c4);
            o.x += bb.x + rv[i].x; o.y += bb.y + rv[i].y; o.z += bb.z + rv[i].z; o.w += bb.w + rv[i].w;
            s0 += o.x; s1 += o.y; s2 += o.z; s3 += o.w;
            q0 = fmaf(o.x, o.x, q0); q1 = fmaf(o.y, o.y, q1); q2 = fmaf(o.z, o.z, q2); q3 = fmaf(o.w, o.w, q3);
            const size_t gi = (size_t)(off + mm) * BN + n;
            if (a.out_bf16) {
              uint2 pk;
              pk.x = pack_bf16(o.x, o.y); pk.y = pack_bf16(o.z, o.w);
              *reinterpret_cast<uint2*>(a.out_bf16 + gi) = pk;
            }
            if (a.out_f32) {
              o.x *= a.oscale; o.y *= a.oscale; o.z *= a.oscale; o.w *= a.oscale;
              if (a.accumulate) {
                const float4 pv = *reinterpret_cast<const float4*>(a.out_f32 + gi);
                o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
              }
              *reinterpret_cast<float4*>(a.out_f32 + gi) = o;
            }
          }
          // next chunk's residual operand in flight during the next TMEM -> smem hop
          if (c + 32 < BN) fetch(sub, c + 32, rv);
          else if (sub + 1 < nsub) fetch(sub + 1, 0, rv);
          if (a.part) {
            // rows of one column quad live in lanes l, l+8, l+16, l+24
            s0 += __shfl_xor_sync(0xffffffffu, s0, 8); s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 8); s3 += __shfl_xor_sync(0xffffffffu, s3, 8);
            q0 += __shfl_xor_sync(0xffffffffu, q0, 8); q1 += __shfl_xor_sync(0xffffffffu, q1, 8);
            q2 += __shfl_xor_sync(0xffffffffu, q2, 8); q3 += __shfl_xor_sync(0xffffffffu, q3, 8);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 16); s3 += __shfl_xor_sync(0xffffffffu, s3, 16);
            q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
            q2 += __shfl_xor_sync(0xffffffffu, q2, 16); q3 += __shfl_xor_sync(0xffffffffu, q3, 16);
            if (lane < 8) {
              float* sp = stat_f + ((size_t)(warp - 2) * BN + n) * 2;
              *reinterpret_cast<float4*>(sp) = make_float4(s0, q0, s1, q1);
              *reinterpret_cast<float4*>(sp + 4) = make_float4(s2, q2, s3, q3);
            }
          }
        }
        if (a.part) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float* pp = a.part + ((size_t)b * a.nchunk + (size_t)(sm0 >> 7)) * 2 * BN;
          for (int n = t; n < BN; n += 128) {
            float S = 0.f, Q = 0.f;
#pragma unroll
            for (int w = 0; w < 4; w++) { S += stat_f[((size_t)w * BN + n) * 2]; Q += stat_f[((size_t)w * BN + n) * 2 + 1]; }
            pp[n] = S; pp[BN + n] = Q;
          }
          // the next sub-tile's first statistic write happens after its own bar.sync
        }
      }
      if (nsub <= 0) {   // cannot happen (tiles only cover rows < len); keep the pipeline protocol intact anyway
        tc_fence_before();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty(buf)) : "memory");
      }
    }
  } else {
    // ------------------------------------------------------------------ operand producers
    const int pt = threadIdx.x - 192;       // 0..255
    const int cg = pt & 7;                  // 8-channel group inside the 64-channel chunk
    const int rl = pt >> 3;                 // row lane 0..31
    constexpr int GP = (MSUB == 2) ? 5 : 3; // passes (of 32 rows) per load group
    constexpr int NG = 2;
    const int ra_used = MT + 2 * pad;       // <= Cfg::RA
    int gA = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int b, mt;
      tile_decode(s_ts, B, tile, b, mt);
      const int L = a.len[b], off = a.off[b];
      const int r0 = mt * MT - pad;         // item-relative row of A-slot row 0
      for (int c = 0; c < KCH; c++, gA++) {
        const int sa = gA % NA;
        const int ch0 = c * 64 + cg * 8;
        float sc[8], sh[8], al[8], ial[8];
        {
          const float4 t0 = *reinterpret_cast<const float4*>(a.scale + (size_t)b * BN + ch0);
          const float4 t1 = *reinterpret_cast<const float4*>(a.scale + (size_t)b * BN + ch0 + 4);
          sc[0] = t0.x; sc[1] = t0.y; sc[2] = t0.z; sc[3] = t0.w; sc[4] = t1.x; sc[5] = t1.y; sc[6] = t1.z; sc[7] = t1.w;
          const float4 u0 = *reinterpret_cast<const float4*>(a.shift + (size_t)b * BN + ch0);
          const float4 u1 = *reinterpret_cast<const float4*>(a.shift + (size_t)b * BN + ch0 + 4);
          sh[0] = u0.x; sh[1] = u0.y; sh[2] = u0.z; sh[3] = u0.w; sh[4] = u1.x; sh[5] = u1.y; sh[6] = u1.z; sh[7] = u1.w;
          const float4 v0 = *reinterpret_cast<const float4*>(a.alpha + ch0);
          const float4 v1 = *reinterpret_cast<const float4*>(a.alpha + ch0 + 4);
          al[0] = v0.x; al[1] = v0.y; al[2] = v0.z; al[3] = v0.w; al[4] = v1.x; al[5] = v1.y; al[6] = v1.z; al[7] = v1.w;
#pragma unroll
          for (int e = 0; e < 8; e++) ial[e] = __fdividef(1.f, al[e]);
        }
        uint8_t* const slot = gbase + (a_base - base) + (size_t)sa * Cfg::A_SLOT;
#pragma unroll
        for (int g = 0; g < NG; g++) {
          float xv[GP][8];
          // issue this group's loads first (they overlap the slot wait and the previous group's math)
#pragma unroll
          for (int p = 0; p < GP; p++) {
            const int rloc = (g * GP + p) * 32 + rl;
            const int gr = r0 + rloc;
            const bool ok = rloc < ra_used && gr >= 0 && gr < L;
            if (IN_BF16) {
              uint4 raw = make_uint4(0u, 0u, 0u, 0u);
              if (ok) raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)(off + gr) * BN + ch0);
              xv[p][0] = __uint_as_float(raw.x << 16); xv[p][1] = __uint_as_float(raw.x & 0xFFFF0000u);
              xv[p][2] = __uint_as_float(raw.y << 16); xv[p][3] = __uint_as_float(raw.y & 0xFFFF0000u);
              xv[p][4] = __uint_as_float(raw.z << 16); xv[p][5] = __uint_as_float(raw.z & 0xFFFF0000u);
              xv[p][6] = __uint_as_float(raw.w << 16); xv[p][7] = __uint_as_float(raw.w & 0xFFFF0000u);
            } else {
              float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
              if (ok) {
                const float* xp = reinterpret_cast<const float*>(a.x) + (size_t)(off + gr) * BN + ch0;
                f0 = *reinterpret_cast<const float4*>(xp);
                f1 = *reinterpret_cast<const float4*>(xp + 4);
              }
              xv[p][0] = f0.x; xv[p][1] = f0.y; xv[p][2] = f0.z; xv[p][3] = f0.w;
              xv[p][4] = f1.x; xv[p][5] = f1.y; xv[p][6] = f1.z; xv[p][7] = f1.w;
            }
          }
          if (g == 0) mbar_wait(emptyA(sa), (((uint32_t)(gA / NA)) & 1u) ^ 1u);
#pragma unroll
          for (int p = 0; p < GP; p++) {
            const int rloc = (g * GP + p) * 32 + rl;
            if (rloc >= ra_used) continue;
            const int gr = r0 + rloc;
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (gr >= 0 && gr < L) {
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; e++) y[e] = snake_f(fmaf(xv[p][e], sc[e], sh[e]), al[e], ial[e]);
              pk.x = pack_bf16(y[0], y[1]); pk.y = pack_bf16(y[2], y[3]);
              pk.z = pack_bf16(y[4], y[5]); pk.w = pack_bf16(y[6], y[7]);
            }
            *reinterpret_cast<uint4*>(slot + (size_t)rloc * 128 + ((cg ^ (rloc & 7)) << 4)) = pk;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fullA(sa)) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int BN, int MSUB, bool IN_BF16>
void launch_arb_t(const ArbConvArgs& a, cudaStream_t st) {
  using Cfg = ArbCfg<BN, MSUB>;
  static bool attr_set[64] = {false};
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev]) {
    KKX_CUDA(cudaFuncSetAttribute(arb_conv_kernel<BN, MSUB, IN_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    KKX_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
    attr_set[dev] = true;
  }
  const int nsm = dev < 64 && sms[dev] > 0 ? sms[dev] : 148;
  const int grid = a.total_tiles < nsm ? a.total_tiles : nsm;
  arb_conv_kernel<BN, MSUB, IN_BF16><<<grid, kArbThreads, Cfg::SMEM, st>>>(*reinterpret_cast<const CUtensorMap*>(a.tmB), a);
}

}  // namespace

int arb_tile_rows(int C) { return C == 128 ? 256 : 128; }

bool arb_conv_supported(int C, int ks, int dil, int B) {
  return (C == 128 || C == 256) && ks >= 1 && dil >= 1 && dil * (ks - 1) <= 50 && (ks & 1) && B <= kArbMaxB;
}

void launch_arb_conv(const ArbConvArgs& a, cudaStream_t st) {
  if (g_dry_run) return;
  if (a.total_tiles <= 0 || a.B <= 0) return;
  if (!arb_conv_supported(a.C, a.ks, a.dil, a.B)) throw ArgError("launch_arb_conv: unsupported shape");
  if (g_launch_stats) g_launch_stats->conv_flops += 2.0 * (double)a.sum_m * a.C * a.C * a.ks;
  if (a.C == 128) {
    if (a.in_bf16) launch_arb_t<128, 2, true>(a, st); else launch_arb_t<128, 2, false>(a, st);
  } else {
    if (a.in_bf16) launch_arb_t<256, 1, true>(a, st); else launch_arb_t<256, 1, false>(a, st);
  }
  if (g_launch_stats && g_launch_stats->profile && g_launch_stats->detail) {
    char nm[96]; snprintf(nm, sizeof nm, "arb_conv[c%d k%d d%d %s m%lld]", a.C, a.ks, a.dil, a.in_bf16 ? "bf16" : "f32", a.sum_m);
    post_launch(nm, st);
  } else post_launch("arb_conv", st);
}

}  // namespace kkx
