// Tensor-core path (tcgen05 + TMEM + TMA) for the fused dilated residual layer, sm_100a only.
//
// Precision: the reference's gradients are only reproducible with fp32-equivalent arithmetic (a
// plain TF32 forward flips max-over-stages winners and moves late-stage gradients by several
// percent -- DESIGN.md "precision").  Every GEMM here is therefore error-compensated 3xTF32:
//     x * W  ~=  x_hi*W_hi + x_hi*W_lo + x_lo*W_hi
// x_hi = trunc_tf32(x) is what the tensor core sees when handed the raw fp32 word (low 13 mantissa
// bits ignored); x_lo = rna_tf32(x - x_hi) is produced in registers and parked in TMEM as the
// A operand of the third product; W_hi = rna_tf32(W), W_lo = rna_tf32(W - W_hi) are prepared once
// per optimizer step by tc_pack_layer_kernel.  Accumulation is fp32 in TMEM.
//
// Data flow per 128-frame tile (one CTA per SM, persistent):
//   TMA  : 3 tap tiles x (t0-d, t0, t0+d) -> smem, SWIZZLE_128B, out-of-range rows zero-filled
//          (that zero fill IS the conv padding, networks.py:339)
//   MMA  : H[128x64] (TMEM) = sum over taps of the three products (72 x tcgen05.mma m128 n64 k8)
//   EPI1 : H -> regs, +bd, relu -> h (global, for backward) ; h_hi/h_lo back into TMEM
//   MMA  : O[128x64] (TMEM) = 3xTF32(h, W1)                         (24 x tcgen05.mma)
//   EPI2 : O -> regs, +b1, dropout, + x (kept in regs from the centre tap), * mask -> y (global)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
// (thread <-> frame row <-> TMEM lane).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "layout.h"
#include "tc_ptx.cuh"

namespace mstcn {
namespace tc {

constexpr int TM = 128;                       // frames per tile (UMMA M)
constexpr int kThreads = 192;
constexpr int kSubA = TM * 32 * 4;            // 16 KB: 128 rows x 32 fp32 (one 128B-swizzle column of A)
constexpr int kSubB = 64 * 32 * 4;            // 8 KB : 64 rows  x 32 fp32 (same for B)
constexpr int kSlot = 2 * kSubA;              // one tap tile: 64 channels = two sub-tiles
// per-layer weight image (floats): [Wd_hi 6 sub | Wd_lo 6 sub | W1_hi 2 sub | W1_lo 2 sub]
constexpr int kWimgFloats = (6 + 6 + 2 + 2) * (kSubB / 4);          // 32768 floats = 128 KB
constexpr int kOffWdHi = 0, kOffWdLo = 6 * kSubB, kOffW1Hi = 12 * kSubB, kOffW1Lo = 14 * kSubB;
constexpr int kOffSlots = 16 * kSubB;                                // 131072
constexpr int kOffBias = kOffSlots + 3 * kSlot;                      // 229376
constexpr int kOffBars = kOffBias + 2 * 64 * 4;                      // 229888
constexpr int kNumBars = 10;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kTcFwdSmem = kOffTmemPtr + 16 + 1024;                  // + slack to 1024-align the base
// TMEM columns
constexpr uint32_t kColAlo = 0, kColH = 192, kColHlo = 256, kColO = 320, kTmemCols = 512;

// element (n = output row, kk = K index) of a K-major SWIZZLE_128B operand image made of [rows x 32] sub-tiles
__host__ __device__ inline int wimg_index(int n, int kk, int rows) {
  const int sub = kk >> 5, k32 = kk & 31;
  return sub * rows * 32 + (n >> 3) * 256 + (n & 7) * 32 + ((((k32 >> 2) ^ (n & 7)) & 7) << 2) + (k32 & 3);
}

// One block per (stage, layer): native (out,in,tap) weights -> hi/lo TF32 images in UMMA layout.
// K index of the dilated conv = tap*64 + in_channel.
__global__ void __launch_bounds__(256) tc_pack_layer_kernel(Layout lay, const float* __restrict__ params,
                                                            float* __restrict__ wimg) {
  const int s = blockIdx.x / lay.L, l = blockIdx.x % lay.L;
  const float* wd = params + lay.wd(s, l);
  const float* w1 = params + lay.w1(s, l);
  float* img = wimg + (size_t)blockIdx.x * kWimgFloats;
  for (int i = threadIdx.x; i < 12288; i += blockDim.x) {
    const int o = i / 192, r = i % 192, c = r / 3, k = r % 3;      // native index (o, c, k)
    const float w = wd[i];
    const uint32_t hi = tf32_rna(w);
    const uint32_t lo = tf32_rna(w - __uint_as_float(hi));
    const int idx = wimg_index(o, k * 64 + c, 64);
    img[idx] = __uint_as_float(hi);
    img[kOffWdLo / 4 + idx] = __uint_as_float(lo);
  }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
    const int o = i >> 6, c = i & 63;
    const float w = w1[i];
    const uint32_t hi = tf32_rna(w);
    const uint32_t lo = tf32_rna(w - __uint_as_float(hi));
    const int idx = wimg_index(o, c, 64);
    img[kOffW1Hi / 4 + idx] = __uint_as_float(hi);
    img[kOffW1Lo / 4 + idx] = __uint_as_float(lo);
  }
}

struct TcLayerFwdArgs {
  const int* lens; const float* wimg; const float* bd; const float* b1;
  float* y; float* h;
  int B, T, d, tiles_per_video, num_tiles;
  int train; uint32_t layer_id; uint64_t seed, offset;
};

// byte offset of (row, 16-byte chunk q of 8) inside one [128 x 32 fp32] SWIZZLE_128B sub-tile
__device__ __forceinline__ uint32_t sw128_off(int row, int q) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((q ^ row) & 7) << 4));
}

__global__ void __launch_bounds__(kThreads, 1)
tc_layer_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, TcLayerFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sBias = reinterpret_cast<float*>(smem + kOffBias);            // bd[64] | b1[64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* bar_full = bars;            // [3] TMA bytes of tap k landed
  uint64_t* bar_lo = bars + 3;          // [3] x_lo of tap k parked in TMEM (128 arrivals)
  uint64_t* bar_g1 = bars + 6;          // H accumulator complete
  uint64_t* bar_h = bars + 7;           // h_hi / h_lo parked in TMEM (128 arrivals)
  uint64_t* bar_g2 = bars + 8;          // O accumulator complete
  uint64_t* bar_free = bars + 9;        // tap slots may be overwritten
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup ----
  for (int i = tid; i < kWimgFloats / 4; i += kThreads)
    reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(a.wimg) + i);
  if (tid < 64) sBias[tid] = __ldg(a.bd + tid);
  else if (tid < 128) sBias[tid] = __ldg(a.b1 + tid - 64);
  fence_proxy_async_smem();                     // generic-proxy weight writes -> visible to the MMA (async proxy)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int k = 0; k < 3; ++k) { mbar_init(bar_full + k, 1); mbar_init(bar_lo + k, 128); }
    mbar_init(bar_g1, 1); mbar_init(bar_h, 128); mbar_init(bar_g2, 1); mbar_init(bar_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  constexpr uint32_t idesc = umma_idesc_tf32(TM, 64);
  const uint32_t sbase = smem_u32(smem);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b)) continue;
        mbar_wait(bar_free, (it & 1) ^ 1);
        const int order[3] = {1, 0, 2};
#pragma unroll
        for (int oi = 0; oi < 3; ++oi) {
          const int k = order[oi];
          const int tf = t0 + (k - 1) * a.d;
          const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
          if (present) {
            mbar_arrive_expect_tx(bar_full + k, kSlot);
            uint8_t* dst = smem + kOffSlots + k * kSlot;
            tma_load_3d(dst, &tm_x, bar_full + k, 0, tf, b);
            tma_load_3d(dst + kSubA, &tm_x, bar_full + k, 32, tf, b);
          } else {
            mbar_arrive(bar_full + k);          // keep the phase in step; the tap contributes exactly 0
          }
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b)) continue;
        const uint32_t p = it & 1;
        uint32_t acc = 0;
        const int order[3] = {1, 0, 2};
#pragma unroll
        for (int oi = 0; oi < 3; ++oi) {
          const int k = order[oi];
          const int tf = t0 + (k - 1) * a.d;
          const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
          mbar_wait(bar_full + k, p);
          tc_fence_after_sync();
          if (present) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = umma_desc_sw128(sbase + kOffSlots + k * kSlot + s * kSubA + ks * 32);
                const uint32_t woff = (k * 2 + s) * kSubB + ks * 32;
                umma_tf32_ss(tmem + kColH, ad, umma_desc_sw128(sbase + kOffWdHi + woff), idesc, acc);
                acc = 1;
                umma_tf32_ss(tmem + kColH, ad, umma_desc_sw128(sbase + kOffWdLo + woff), idesc, 1);
              }
          }
          mbar_wait(bar_lo + k, p);
          tc_fence_after_sync();
          if (present) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t woff = (k * 2 + s) * kSubB + ks * 32;
                umma_tf32_ts(tmem + kColH, tmem + kColAlo + k * 64 + s * 32 + ks * 8,
                             umma_desc_sw128(sbase + kOffWdHi + woff), idesc, 1);
              }
          }
        }
        umma_commit(bar_g1);
        mbar_wait(bar_h, p);
        tc_fence_after_sync();
        acc = 0;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t woff = s * kSubB + ks * 32;
            const uint32_t ah = tmem + kColH + s * 32 + ks * 8, al = tmem + kColHlo + s * 32 + ks * 8;
            umma_tf32_ts(tmem + kColO, ah, umma_desc_sw128(sbase + kOffW1Hi + woff), idesc, acc);
            acc = 1;
            umma_tf32_ts(tmem + kColO, ah, umma_desc_sw128(sbase + kOffW1Lo + woff), idesc, 1);
            umma_tf32_ts(tmem + kColO, al, umma_desc_sw128(sbase + kOffW1Hi + woff), idesc, 1);
          }
        umma_commit(bar_g2);
        ++it;
      }
    }
  } else {
    // =============================== epilogue warps ==============================
    const int wq = warp & 3;                    // TMEM lane quadrant this warp may touch
    const int row = wq * 32 + lane;             // frame row inside the tile
    const int etid = tid - 64;                  // 0..127
    const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
      const size_t vbase = (size_t)b * a.T * C;
      if (t0 >= len) {                          // padding tile: y = 0 (mask); h is never read there
        for (int i = etid; i < TM * 16; i += 128) {
          const int t = t0 + (i >> 4);
          if (t < a.T) reinterpret_cast<float4*>(a.y + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        continue;
      }
      const uint32_t p = it & 1;
      const int t = t0 + row;
      float xc[64];
      const int order[3] = {1, 0, 2};
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * a.d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
        mbar_wait(bar_full + k, p);
        if (present) {
          const uint8_t* slot = smem + kOffSlots + k * kSlot;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            uint32_t lo[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(slot + s * kSubA + sw128_off(row, q));
              lo[4 * q + 0] = tf32_lo_of(v.x); lo[4 * q + 1] = tf32_lo_of(v.y);
              lo[4 * q + 2] = tf32_lo_of(v.z); lo[4 * q + 3] = tf32_lo_of(v.w);
              if (k == 1) { xc[s * 32 + 4 * q] = v.x; xc[s * 32 + 4 * q + 1] = v.y; xc[s * 32 + 4 * q + 2] = v.z; xc[s * 32 + 4 * q + 3] = v.w; }
            }
            tmem_st32(trow + kColAlo + k * 64 + s * 32, lo);
          }
          tmem_wait_st();
        }
        tc_fence_before_sync();
        mbar_arrive(bar_lo + k);
      }
      // ---- EPI1: H -> relu -> h (global) and back into TMEM as the A operand of the 1x1 ----
      mbar_wait(bar_g1, p);
      tc_fence_after_sync();
      if (etid == 0) mbar_arrive(bar_free);     // every MMA and every epilogue read of the tap slots is done
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        uint32_t v[32], lo[32];
        tmem_ld32(trow + kColH + s * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float hv = fmaxf(__uint_as_float(v[i]) + sBias[s * 32 + i], 0.f);
          v[i] = __float_as_uint(hv);
          lo[i] = tf32_lo_of(hv);
        }
        tmem_st32(trow + kColH + s * 32, v);
        tmem_st32(trow + kColHlo + s * 32, lo);
        if (a.h != nullptr && t < a.T) {
          float4* dst = reinterpret_cast<float4*>(a.h + vbase + (size_t)t * C + s * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
        }
      }
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(bar_h);
      // ---- EPI2: O -> +b1, dropout, residual, mask -> y ----
      uint2 bits = make_uint2(0xffffffffu, 0xffffffffu);
      if (a.train) bits = dropout_bits(a.seed, a.offset, a.layer_id, (uint32_t)(b * a.T + t));
      const float m = (t < len) ? 1.f : 0.f;
      mbar_wait(bar_g2, p);
      tc_fence_after_sync();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        uint32_t v[32];
        tmem_ld32(trow + kColO + s * 32, v);
        tmem_wait_ld();
        const uint32_t w = s == 0 ? bits.x : bits.y;
        float o[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float ov = __uint_as_float(v[i]) + sBias[64 + s * 32 + i];
          if (a.train) ov *= ((w >> i) & 1u) ? 2.f : 0.f;
          o[i] = (xc[s * 32 + i] + ov) * m;
        }
        if (t < a.T) {
          float4* dst = reinterpret_cast<float4*>(a.y + vbase + (size_t)t * C + s * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
      }
      tc_fence_before_sync();
      ++it;
    }
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tc
}  // namespace mstcn
