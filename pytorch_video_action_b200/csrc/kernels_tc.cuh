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
constexpr int kSubA = TM * 32 * 4;            // 16 KB: 128 rows x 32 fp32 (one 128B-swizzle column of A)
constexpr int kSubB = 64 * 32 * 4;            // 8 KB : 64 rows  x 32 fp32 (same for B)
constexpr int kSlot = 2 * kSubA;              // one tap tile: 64 channels = two sub-tiles
// per-layer weight image (floats): [Wd_hi 6 sub | Wd_lo 6 sub | W1_hi 2 sub | W1_lo 2 sub]
constexpr int kWimgFloats = (6 + 6 + 2 + 2) * (kSubB / 4);          // 32768 floats = 128 KB
constexpr int kOffWdHi = 0, kOffWdLo = 6 * kSubB, kOffW1Hi = 12 * kSubB, kOffW1Lo = 14 * kSubB;
constexpr int kOffSlots = 16 * kSubB;                                // 131072
constexpr int kOffBias = kOffSlots + 3 * kSlot;                      // 229376
constexpr int kOffBars = kOffBias + 2 * 64 * 4;                      // 229888
constexpr int kNumBars = 14;
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
  float* img = wimg + (size_t)blockIdx.x * 2 * kWimgFloats;
  const int tid0 = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;
  for (int i = tid0; i < 12288; i += nth) {
    const int o = i / 192, r = i % 192, c = r / 3, k = r % 3;      // native index (o, c, k)
    const float w = wd[i];
    const uint32_t hi = tf32_rna(w);
    const uint32_t lo = tf32_rna(w - __uint_as_float(hi));
    const int idx = wimg_index(o, k * 64 + c, 64);
    img[idx] = __uint_as_float(hi);
    img[kOffWdLo / 4 + idx] = __uint_as_float(lo);
  }
  for (int i = tid0; i < 4096; i += nth) {
    const int o = i >> 6, c = i & 63;
    const float w = w1[i];
    const uint32_t hi = tf32_rna(w);
    const uint32_t lo = tf32_rna(w - __uint_as_float(hi));
    const int idx = wimg_index(o, c, 64);
    img[kOffW1Hi / 4 + idx] = __uint_as_float(hi);
    img[kOffW1Lo / 4 + idx] = __uint_as_float(lo);
  }
  // backward images (same shape): B[n = in_channel][K = tap*64 + out_channel] = Wd[o][c][k] for the
  // input-gradient GEMM, and B[n = in][K = out] = W1[o][c] for gh = W1^T go
  float* imgb = img + kWimgFloats;
  for (int i = tid0; i < 12288; i += nth) {
    const int o = i / 192, r = i % 192, c = r / 3, k = r % 3;
    const float w = wd[i];
    const uint32_t hi = tf32_rna(w);
    const uint32_t lo = tf32_rna(w - __uint_as_float(hi));
    const int idx = wimg_index(c, k * 64 + o, 64);
    imgb[idx] = __uint_as_float(hi);
    imgb[kOffWdLo / 4 + idx] = __uint_as_float(lo);
  }
  for (int i = tid0; i < 4096; i += nth) {
    const int o = i >> 6, c = i & 63;
    const float w = w1[i];
    const uint32_t hi = tf32_rna(w);
    const uint32_t lo = tf32_rna(w - __uint_as_float(hi));
    const int idx = wimg_index(c, o, 64);
    imgb[kOffW1Hi / 4 + idx] = __uint_as_float(hi);
    imgb[kOffW1Lo / 4 + idx] = __uint_as_float(lo);
  }
}

struct TcLayerFwdArgs {
  const int* lens; const float* wimg; const float* bd; const float* b1;
  float* y; float* h;
  int B, T, d, tiles_per_video, num_tiles;      // d < 0 in backward-gx mode (taps at t -/+ d swap roles)
  int skip_extra;                               // tiles starting at or after len + skip_extra are all zero
  int train; uint32_t layer_id; uint64_t seed, offset;
  long long* dbg;     // optional: SM-clock timestamps of CTA 0's first tile (mstcn_debug_tc_timing)
};

#define TC_STAMP(slot) do { if (a.dbg != nullptr && blockIdx.x == 0) a.dbg[slot] = clock64(); } while (0)

// byte offset of (row, 16-byte chunk q of 8) inside one [128 x 32 fp32] SWIZZLE_128B sub-tile
__device__ __forceinline__ uint32_t sw128_off(int row, int q) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((q ^ row) & 7) << 4));
}
// same for the SWIZZLE_128B_ATOM_32B mode the weight-gradient tiles use (32-byte chunks swizzled by row & 3)
__device__ __forceinline__ uint32_t sw32_off(int row, int q) {
  return (uint32_t)(row * 128 + ((((q >> 1) ^ row) & 3) << 5) + ((q & 1) << 4));
}
// output staging tile [128 rows][16 chunks of 16 B], chunk XOR-swizzled by the row
__device__ __forceinline__ uint32_t stage_off(int row, int q) { return (uint32_t)(row * 256 + (((q ^ row) & 15) << 4)); }

// x_lo = x - trunc_tf32(x): exact; handed to the tensor core as is (it keeps the top 11 bits)
__device__ __forceinline__ uint32_t lo_bits(float x) {
  return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u));
}

// One epilogue warp pair (same TMEM lane quadrant q, column halves s = 0 / 1) has staged its 32 rows in
// `stage`; after the pair barrier each warp writes 16 full 256-byte rows, two rows per instruction.
__device__ __forceinline__ void copy_out_rows(const uint8_t* stage, float* __restrict__ gvid, int t0, int T,
                                              int q, int s, int lane) {
  named_bar_sync(1 + q, 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 32 * q + 16 * s + 2 * i + (lane >> 4), t = t0 + r;
    const float4 v = *reinterpret_cast<const float4*>(stage + stage_off(r, lane & 15));
    if (t < T) reinterpret_cast<float4*>(gvid + (size_t)t * C)[lane & 15] = v;
  }
}

constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 64 + 32 * kEpiWarps;     // 320

// MODE 0: DilatedResidualLayer.forward.
// MODE 1: the input-gradient half of its backward, gx[t] = gy[t]*mask + sum_k Wd[:,:,k]^T gu[t-(k-1)d]:
//         same tap GEMM on gu with the transposed weight image (a.d = -dilation), no 1x1, and an epilogue
//         that adds the residual-branch gradient.  tm_x maps gu, tm_g maps gy, a.y receives gx.
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g, TcLayerFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sBias = reinterpret_cast<float*>(smem + kOffBias);            // bd[64] | b1[64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* bar_full = bars;            // [3] TMA bytes of tap k landed
  uint64_t* bar_lo = bars + 3;          // [3] x_lo of tap k parked in TMEM (one arrival per epilogue warp)
  uint64_t* bar_g1 = bars + 6;          // H accumulator complete
  uint64_t* bar_h = bars + 7;           // h_hi / h_lo parked in TMEM
  uint64_t* bar_g2 = bars + 8;          // O accumulator complete
  uint64_t* bar_free = bars + 9;        // [3] tap slot k may be overwritten
  uint64_t* bar_wd = bars + 12;         // dilated-conv weight images landed (once per launch)
  uint64_t* bar_w1 = bars + 13;         // 1x1 weight images landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) TC_STAMP(0);

  // ---- one-time setup (nothing here depends on the previous kernel: it overlaps that kernel's tail
  //      under programmatic dependent launch) ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int k = 0; k < 3; ++k) { mbar_init(bar_full + k, 1); mbar_init(bar_lo + k, kEpiWarps); }
    mbar_init(bar_g1, 1); mbar_init(bar_h, kEpiWarps); mbar_init(bar_g2, 1);
    mbar_init(bar_free + 0, kEpiWarps); mbar_init(bar_free + 1, 1); mbar_init(bar_free + 2, kEpiWarps);
    mbar_init(bar_wd, 1); mbar_init(bar_w1, 1);
    fence_barrier_init();
    // 128 KB operand image: 12 + 4 bulk copies of one 8 KB sub-tile each (async proxy -> no proxy fence)
    mbar_arrive_expect_tx(bar_wd, 12 * kSubB);
    for (int i = 0; i < 12; ++i) bulk_load(smem + i * kSubB, a.wimg + i * (kSubB / 4), kSubB, bar_wd);
    if (MODE == 0) {
      mbar_arrive_expect_tx(bar_w1, 4 * kSubB);
      for (int i = 12; i < 16; ++i) bulk_load(smem + i * kSubB, a.wimg + i * (kSubB / 4), kSubB, bar_w1);
    } else {
      tma_prefetch_desc(&tm_g);
    }
  }
  if (MODE == 0) {
    if (tid >= 64 && tid < 128) sBias[tid - 64] = __ldg(a.bd + tid - 64);
    else if (tid >= 128 && tid < 192) sBias[tid - 64] = __ldg(a.b1 + tid - 128);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) TC_STAMP(1);
  pdl_launch_dependents();                      // the next layer's prologue may start as SMs free up
  pdl_wait();                                   // the previous kernel's activations are complete and visible
  if (tid == 0) TC_STAMP(2);
  const uint32_t tmem = *tmem_ptr;
  constexpr uint32_t idesc = umma_idesc_tf32(TM, 64);
  const uint32_t sbase = smem_u32(smem);
  const int order[3] = {1, 0, 2};               // centre tap first: it always exists and seeds the accumulator

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b) + a.skip_extra) continue;
#pragma unroll
        for (int oi = 0; oi < 3; ++oi) {
          const int k = order[oi];
          const int tf = t0 + (k - 1) * a.d;
          const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
          const bool with_gy = MODE == 1 && oi == 2;      // gy rides on the last tap's barrier (freed last)
          mbar_wait(bar_free + k, (it & 1) ^ 1);
          if (present || with_gy) {
            mbar_arrive_expect_tx(bar_full + k, (present ? kSlot : 0) + (with_gy ? kSlot : 0));
            if (present) {
              uint8_t* dst = smem + kOffSlots + k * kSlot;
              tma_load_3d(dst, &tm_x, bar_full + k, 0, tf, b);
              tma_load_3d(dst + kSubA, &tm_x, bar_full + k, 32, tf, b);
            }
            if (with_gy) {
              tma_load_3d(smem + kOffW1Hi, &tm_g, bar_full + k, 0, t0, b);
              tma_load_3d(smem + kOffW1Hi + kSubA, &tm_g, bar_full + k, 32, t0, b);
            }
            if (it == 0 && oi == 0) TC_STAMP(3);
          } else {
            mbar_arrive(bar_full + k);          // keep the phase in step; the tap contributes exactly 0
          }
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // Converged warp, one predicated issuer (see tc_ptx.cuh).  Descriptor low words = base + constant.
    const uint32_t leader = lane == 0 ? 1u : 0u;
    // REDUX results live in uniform registers: the compiler then knows every operand below is uniform
    const uint32_t usbase = __reduce_or_sync(0xffffffffu, sbase), utmem = __reduce_or_sync(0xffffffffu, tmem);
    const uint32_t a0 = umma_desc_lo(usbase + kOffSlots);
    const uint32_t wdh = umma_desc_lo(usbase + kOffWdHi), wdl = umma_desc_lo(usbase + kOffWdLo);
    const uint32_t w1h = umma_desc_lo(usbase + kOffW1Hi), w1l = umma_desc_lo(usbase + kOffW1Lo);
    const uint32_t tH = utmem + kColH, tO = utmem + kColO, tAlo = utmem + kColAlo, tHlo = utmem + kColHlo;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      if (t0 >= __ldg(a.lens + b) + a.skip_extra) continue;
      const uint32_t p = it & 1;
      if (it == 0) { mbar_wait(bar_wd, 0); if (lane == 0) TC_STAMP(4); }
      // x_hi * (W_hi + W_lo): needs only the TMA data.  Centre tap first (always present, seeds H).
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * a.d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
        mbar_wait(bar_full + k, p);
        if (it == 0 && oi == 0 && lane == 0) TC_STAMP(5);
        tc_fence_after_sync();
        if (present) {
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t ad = a0 + ((k * kSlot + s * kSubA + ks * 32) >> 4);
              const uint32_t wo = ((k * 2 + s) * kSubB + ks * 32) >> 4;
              umma_tf32_ss(tH, ad, wdh + wo, idesc, (oi | s | ks) != 0, leader);
              umma_tf32_ss(tH, ad, wdl + wo, idesc, 1, leader);
            }
        }
      }
      // x_lo * W_hi: A operand from TMEM once the epilogue warps have parked it
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * a.d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
        mbar_wait(bar_lo + k, p);
        tc_fence_after_sync();
        if (present) {
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32_ts(tH, tAlo + k * 64 + s * 32 + ks * 8, wdh + (((k * 2 + s) * kSubB + ks * 32) >> 4), idesc, 1, leader);
        }
      }
      umma_commit(bar_g1, leader);
      if (it == 0 && lane == 0) TC_STAMP(6);
      if (MODE == 1) { ++it; continue; }
      mbar_wait(bar_h, p);
      if (it == 0 && lane == 0) TC_STAMP(7);
      if (it == 0) mbar_wait(bar_w1, 0);
      tc_fence_after_sync();
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t wo = (s * kSubB + ks * 32) >> 4;
          const uint32_t ah = tH + s * 32 + ks * 8, al = tHlo + s * 32 + ks * 8;
          umma_tf32_ts(tO, ah, w1h + wo, idesc, (s | ks) != 0, leader);
          umma_tf32_ts(tO, ah, w1l + wo, idesc, 1, leader);
          umma_tf32_ts(tO, al, w1h + wo, idesc, 1, leader);
        }
      umma_commit(bar_g2, leader);
      if (it == 0 && lane == 0) TC_STAMP(8);
      ++it;
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ==============================
    // warp pair (q, s): q = TMEM lane quadrant (= warp % 4, a hardware rule), s = column half
    const int q = warp & 3, s = (warp - 2) >> 2;
    const int row = q * 32 + lane;              // frame row inside the tile
    const int etid = tid - 64;                  // 0..255
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 32);
    const float* biasd = sBias + s * 32;
    const float* bias1 = sBias + 64 + s * 32;
    uint8_t* stage_h = smem + kOffSlots;                // tap-0 slot doubles as h staging
    uint8_t* stage_y = smem + kOffSlots + 2 * kSlot;    // tap-2 slot doubles as y staging
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
      const size_t vbase = (size_t)b * a.T * C;
      if (t0 >= len + a.skip_extra) {           // nothing but zeros reaches this tile: y = 0 (h is never read there)
        for (int i = etid; i < TM * 16; i += 32 * kEpiWarps) {
          const int t = t0 + (i >> 4);
          if (t < a.T) reinterpret_cast<float4*>(a.y + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        continue;
      }
      const uint32_t p = it & 1;
      const int t = t0 + row;
      float xc[32];
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * a.d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
        mbar_wait(bar_full + k, p);
        if (it == 0 && oi == 0 && etid == 0) TC_STAMP(9);
        if (present) {
          const uint8_t* sub = smem + kOffSlots + k * kSlot + s * kSubA;
          uint32_t lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(sub + sw128_off(row, c));
            lo[4 * c + 0] = lo_bits(v.x); lo[4 * c + 1] = lo_bits(v.y);
            lo[4 * c + 2] = lo_bits(v.z); lo[4 * c + 3] = lo_bits(v.w);
            if (MODE == 0 && k == 1) { xc[4 * c] = v.x; xc[4 * c + 1] = v.y; xc[4 * c + 2] = v.z; xc[4 * c + 3] = v.w; }
          }
          tmem_st32(trow + kColAlo + k * 64, lo);
          tmem_wait_st();
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_lo + k);
        if (it == 0 && etid == 0) TC_STAMP(10 + oi);
      }
      // ---- EPI1: H -> +bd, relu -> h ; h_hi / h_lo back into TMEM as the A operand of the 1x1 ----
      mbar_wait(bar_g1, p);
      if (it == 0 && etid == 0) TC_STAMP(13);
      tc_fence_after_sync();
      if (etid == 0) mbar_arrive(bar_free + 1);       // centre slot: every MMA and epilogue read is done
      if (MODE == 1) {
        // gx = (W^T gu) + gy * mask ; gy was TMA-loaded into the (unused) 1x1-weight region
        const float m1 = (t < len) ? 1.f : 0.f;
        const uint8_t* gsub = smem + kOffW1Hi + s * kSubA;
        uint32_t v[32];
        tmem_ld32(trow + kColH, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 g = *reinterpret_cast<const float4*>(gsub + sw128_off(row, c));
          *reinterpret_cast<float4*>(stage_h + stage_off(row, s * 8 + c)) =
              make_float4(__uint_as_float(v[4 * c]) + g.x * m1, __uint_as_float(v[4 * c + 1]) + g.y * m1,
                          __uint_as_float(v[4 * c + 2]) + g.z * m1, __uint_as_float(v[4 * c + 3]) + g.w * m1);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 2);     // gy (and tap 2) slot may be refilled
        copy_out_rows(stage_h, a.y + vbase, t0, a.T, q, s, lane);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 0);
        ++it;
        continue;
      }
      {
        uint32_t v[32], lo[32];
        tmem_ld32(trow + kColH, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float hv = fmaxf(__uint_as_float(v[i]) + biasd[i], 0.f);
          v[i] = __float_as_uint(hv);
          lo[i] = lo_bits(hv);
        }
        tmem_st32(trow + kColH, v);
        tmem_st32(trow + kColHlo, lo);
        if (a.h != nullptr) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stage_h + stage_off(row, s * 8 + c)) =
                make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                            __uint_as_float(v[4 * c + 3]));
        }
        tmem_wait_st();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h);
      if (it == 0 && etid == 0) TC_STAMP(14);
      if (a.h != nullptr) copy_out_rows(stage_h, a.h + vbase, t0, a.T, q, s, lane);
      fence_proxy_async_smem();                       // staging (generic proxy) before the next TMA write (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 0);
      // ---- EPI2: O -> +b1, dropout, residual, mask -> y ----
      uint32_t keep = 0xffffffffu;
      if (a.train) {
        const uint2 bits = dropout_bits(a.seed, a.offset, a.layer_id, (uint32_t)(b * a.T + t));
        keep = s == 0 ? bits.x : bits.y;
      }
      const float m = (t < len) ? 1.f : 0.f;
      const float on = a.train ? 2.f * m : m;         // kept channels: scale 2 (p = 0.5), then the mask
      mbar_wait(bar_g2, p);
      if (it == 0 && etid == 0) TC_STAMP(15);
      tc_fence_after_sync();
      {
        uint32_t v[32];
        tmem_ld32(trow + kColO, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * c + j;
            const float ov = __uint_as_float(v[i]) + bias1[i];
            o[j] = xc[i] * m + (((keep >> i) & 1u) ? ov * on : 0.f);
          }
          *reinterpret_cast<float4*>(stage_y + stage_off(row, s * 8 + c)) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before_sync();
      copy_out_rows(stage_y, a.y + vbase, t0, a.T, q, s, lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 2);
      if (it == 0 && etid == 0) TC_STAMP(16);
      ++it;
    }
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) TC_STAMP(17);
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// =============================================================================================
// Weight-gradient kernel: for every tap k of a tile,
//     D_k[o][c] += sum_t A_k[t][o] * B_k[t][c]           (reduction over the tile's 128 frames)
// as one UMMA chain with BOTH operands MN-major (the channel dimension is the contiguous one in the
// TMA-written tiles, frames are K):  A = [A_hi ; A_lo] stacked along M (128 = 64 channels of hi + 64 of
// lo, so both halves of the split ride one M=128 instruction at full tensor rate), B = rna_tf32(B).
// Lanes 0..63 and 64..127 of the accumulator are added in the epilogue, so A enters exactly and the only
// rounding is the unbiased 2^-12 of B.  Accumulators stay in TMEM for the CTA's whole tile loop; one
// partial per CTA goes to scratch and reduce_partials_kernel sums the grid in fixed order.
// Taps: dWd[k] uses A = gu shifted by -(k-1)d, B = x; dW1 uses A = go = gy*mask*dropout (made by the
// transform warps from the raw gy tile), B = h.  Column sums of A give the bias gradients.
// =============================================================================================
struct TcWgradArgs {
  const int* lens; float* part;
  int B, T, tiles_per_video, num_tiles, ntap;
  int tap_shift[4];        // A tile starts at t0 + tap_shift[k]
  int tap_gy[4];           // 1: A comes from tm_a1 (gy) and is masked / dropped-out in place; B from tm_b1 (h)
  int tap_bias[4];         // 1: emit the column sums of A
  int train; uint32_t layer_id; uint64_t seed, offset;
};
constexpr int kWgStage = 3 * kSlot;                       // A_hi | A_lo | B
constexpr int kWgOffBits = 2 * kWgStage;                  // 128 x uint2 keep-bits
constexpr int kWgOffBars = kWgOffBits + 128 * 8;
constexpr int kWgOffTmemPtr = kWgOffBars + 8 * 8;
constexpr int kTcWgradSmem = kWgOffTmemPtr + 16 + 1024;
constexpr int kWgPartFloats = 4 * 4096 + 4 * 64;          // per CTA: [4][64][64] weight partials | [4][64] bias sums

__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                const __grid_constant__ CUtensorMap tm_b0, const __grid_constant__ CUtensorMap tm_b1, TcWgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint2* sBits = reinterpret_cast<uint2*>(smem + kWgOffBits);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgOffBars);
  uint64_t* bar_full = bars;          // [2] TMA landed
  uint64_t* bar_ready = bars + 2;     // [2] transform done (one arrival per transform warp)
  uint64_t* bar_empty = bars + 4;     // [2] MMAs that read the stage are complete
  uint64_t* bar_done = bars + 6;      // all MMAs complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kWgOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a0); tma_prefetch_desc(&tm_a1); tma_prefetch_desc(&tm_b0); tma_prefetch_desc(&tm_b1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_ready + i, kEpiWarps); mbar_init(bar_empty + i, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);

  // a tap's A tile holds something non-zero only if it overlaps [0, min(T, len)) (gu and go vanish beyond len)
  auto tap_present = [&](int t0, int k, int len) {
    const int tf = t0 + a.tap_shift[k];
    const int lim = len < a.T ? len : a.T;
    return (tf + TM - 1 >= 0) && (tf < lim);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        const int len = __ldg(a.lens + b);
        for (int k = 0; k < a.ntap; ++k) {
          if (!tap_present(t0, k, len)) continue;
          const uint32_t st = n & 1;
          mbar_wait(bar_empty + st, ((n >> 1) & 1) ^ 1);
          uint8_t* base = smem + st * kWgStage;
          const CUtensorMap* ma = a.tap_gy[k] ? &tm_a1 : &tm_a0;
          const CUtensorMap* mb = a.tap_gy[k] ? &tm_b1 : &tm_b0;
          const int tf = t0 + a.tap_shift[k];
          mbar_arrive_expect_tx(bar_full + st, 2 * kSlot);
          tma_load_3d(base, ma, bar_full + st, 0, tf, b);
          tma_load_3d(base + kSubA, ma, bar_full + st, 32, tf, b);
          tma_load_3d(base + 2 * kSlot, mb, bar_full + st, 0, t0, b);
          tma_load_3d(base + 2 * kSlot + kSubA, mb, bar_full + st, 32, t0, b);
          ++n;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t usbase = __reduce_or_sync(0xffffffffu, sbase), utmem = __reduce_or_sync(0xffffffffu, tmem);
    constexpr uint32_t idesc = umma_idesc_tf32(TM, 64) | (1u << 15) | (1u << 16);     // A and B MN-major
    uint32_t n = 0, inited = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
      for (int k = 0; k < a.ntap; ++k) {
        if (!tap_present(t0, k, len)) continue;
        const uint32_t st = n & 1;
        mbar_wait(bar_ready + st, (n >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t ad = umma_desc_lo_mn(usbase + st * kWgStage), bd = umma_desc_lo_mn(usbase + st * kWgStage + 2 * kSlot);
        const uint32_t dk = utmem + k * 64;
        const uint32_t first = (inited >> k) & 1u;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
          umma_tf32_ss(dk, ad + kk * 64, bd + kk * 64, idesc, (kk != 0) | first, 1, kDescHiMn32);
        inited |= 1u << k;
        umma_commit(bar_empty + st, 1);
        ++n;
      }
    }
    umma_commit(bar_done, 1);
    __syncwarp();
  } else {
    // ============ transform warps (then the final epilogue) ============
    const int q = warp & 3, s = (warp - 2) >> 2;
    const int etid = tid - 64;                   // 0..255
    const int j = etid & 15;                     // 16-byte chunk column: channels 4j..4j+3
    const int sub = j >> 3, cq = j & 7;
    float bsum[4][4] = {};
    uint32_t n = 0, used = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= a.ntap || !tap_present(t0, k, len)) continue;
        const uint32_t st = n & 1;
        uint8_t* base = smem + st * kWgStage;
        const bool gy = a.tap_gy[k] != 0;
        if (gy && a.train) {                     // keep-bits of the tile's 128 frames, one Philox call each
          if (etid < TM) sBits[etid] = dropout_bits(a.seed, a.offset, a.layer_id, (uint32_t)(b * a.T + t0 + etid));
          named_bar_sync(5, 32 * kEpiWarps);
        }
        mbar_wait(bar_full + st, (n >> 1) & 1);
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (etid >> 4) + 16 * i;
          const uint32_t off = sub * kSubA + sw32_off(r, cq);
          float4 v = *reinterpret_cast<const float4*>(base + off);
          if (gy) {
            const int t = t0 + r;
            float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < len) {
              sc = make_float4(1.f, 1.f, 1.f, 1.f);
              if (a.train) sc = dropout_scale4(sBits[r], j);
            }
            v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
            *reinterpret_cast<float4*>(base + off) = v;
          }
          cs[0] += v.x; cs[1] += v.y; cs[2] += v.z; cs[3] += v.w;
          uint4 lo = make_uint4(lo_bits(v.x), lo_bits(v.y), lo_bits(v.z), lo_bits(v.w));
          *reinterpret_cast<uint4*>(base + kSlot + off) = lo;
          const float4 w = *reinterpret_cast<const float4*>(base + 2 * kSlot + off);
          *reinterpret_cast<uint4*>(base + 2 * kSlot + off) = make_uint4(tf32_rna(w.x), tf32_rna(w.y), tf32_rna(w.z), tf32_rna(w.w));
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) bsum[k][c] += cs[c];
        fence_proxy_async_smem();                // generic writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready + st);
        if (gy && a.train) named_bar_sync(5, 32 * kEpiWarps);     // sBits may be rewritten for the next tile
        used |= 1u << k;
        ++n;
      }
    }
    // ---- final epilogue: D_k lanes 0..63 (A_hi part) + lanes 64..127 (A_lo part) -> per-CTA partial ----
    // staged through shared memory (swizzled rows) so that the global stores are whole 256-byte rows
    mbar_wait(bar_done, 0);
    tc_fence_after_sync();
    uint8_t* sum = smem;                                     // [4][64 rows][256 B] over the (now idle) stages
    float* part = a.part + (size_t)blockIdx.x * kWgPartFloats;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 32);
    const int orow = (q & 1) * 32 + lane;                    // output channel of this thread's TMEM lane
    if (q >= 2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= a.ntap) continue;
        uint32_t v[32];
        if ((used >> k) & 1u) { tmem_ld32(trow + k * 64, v); tmem_wait_ld(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(sum + k * 16384 + stage_off(orow, s * 8 + c)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    }
    named_bar_sync(5, 32 * kEpiWarps);
    if (q < 2) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= a.ntap) continue;
        uint32_t v[32];
        if ((used >> k) & 1u) { tmem_ld32(trow + k * 64, v); tmem_wait_ld(); }
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4* p4 = reinterpret_cast<float4*>(sum + k * 16384 + stage_off(orow, s * 8 + c));
          float4 lo = *p4;
          *p4 = make_float4(__uint_as_float(v[4 * c]) + lo.x, __uint_as_float(v[4 * c + 1]) + lo.y,
                            __uint_as_float(v[4 * c + 2]) + lo.z, __uint_as_float(v[4 * c + 3]) + lo.w);
        }
      }
    }
    named_bar_sync(5, 32 * kEpiWarps);
    for (int i = etid; i < a.ntap * 64 * 16; i += 32 * kEpiWarps) {     // 16 float4 per row, rows contiguous in `part`
      const int k = i >> 10, r = (i >> 4) & 63, c = i & 15;
      reinterpret_cast<float4*>(part + (k * 64 + r) * 64)[c] = *reinterpret_cast<const float4*>(sum + k * 16384 + stage_off(r, c));
    }
    named_bar_sync(5, 32 * kEpiWarps);
    // bias sums: 16 row-groups x 64 channels per tap -> fixed-order sum
    float* red = reinterpret_cast<float*>(smem);             // [4][16][64]
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < a.ntap)
        *reinterpret_cast<float4*>(red + (k * 16 + (etid >> 4)) * 64 + 4 * j) = make_float4(bsum[k][0], bsum[k][1], bsum[k][2], bsum[k][3]);
    named_bar_sync(5, 32 * kEpiWarps);
    if (etid < 64) {
      for (int k = 0; k < a.ntap; ++k) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) t += red[(k * 16 + g) * 64 + etid];
        part[4 * 4096 + k * 64 + etid] = t;
      }
    }
    tc_fence_before_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// the two instantiations
template __global__ void tc_layer_kernel<0>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);
template __global__ void tc_layer_kernel<1>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);

}  // namespace tc
}  // namespace mstcn
