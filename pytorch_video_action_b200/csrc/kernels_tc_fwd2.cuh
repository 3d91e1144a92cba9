// DilatedResidualLayer.forward (networks.py:343-347), second-generation tensor-core kernel: the tap tiles live in TMEM.
//
// tc_layer_kernel<0> keeps the three tap tiles in shared memory until the tap GEMM has read them and then reuses two of
// the slots as output staging, so a CTA cannot start loading its next tile before the current one has left: consecutive
// tiles of a CTA run back to back (~5.3 us per tile at large batches, against ~1.6 us of tensor time).  Here the epilogue
// warps copy every tap tile into TMEM as it lands -- the raw fp32 words (the tensor core reads them as x_hi) beside the
// x_lo words they already produced -- and BOTH products take their A operand from TMEM:
//     H += x_raw * [W_hi | W_lo]   (m128 n128 k8, A in TMEM)        H += x_lo * W_hi   (m128 n64 k8, A in TMEM)
// so a shared-memory slot is only a landing buffer (free again ~0.3 us after its TMA completes), the output staging has a
// buffer of its own, and the producer prefetches the next tile's taps while this tile is still in its GEMMs / epilogues.
//
// Shared memory (same 226 KB budget): 128 KB weight image | 2 landing slots | 1 staging buffer (h, then y).
// TMEM: X_raw[3 taps] 0..191 | X_lo[3 taps] 192..383 | H 384..511 (h_hi rewritten into 384..447, O = 448..511,
//       h_lo over X_lo[tap 2] = 320..383 once the tap GEMM is complete).
// Same arguments, tensor maps, task order and flag protocol as tc_layer_kernel<0> (results differ in the last bit only: the
// hi and lo products of a K-block are accumulated back to back instead of all hi products first).
#pragma once
#include "kernels_tc.cuh"

namespace mstcn {
namespace tc {

constexpr uint32_t k2ColXraw = 0, k2ColXlo = 192, k2ColH = 384, k2ColO = 448, k2ColHlo = 320;
constexpr int k2OffStage = kOffSlots + 2 * kSlot;

__global__ void __launch_bounds__(kTcLayerThreads, 1)
tc_fwd2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
               const __grid_constant__ CUtensorMap tm_h, TcLayerFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sBias = reinterpret_cast<float*>(smem + kOffBias);            // bd[64] | b1[64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* bar_full = bars;            // [2] landing slot: TMA bytes landed
  uint64_t* bar_sfree = bars + 2;       // [2] landing slot: every epilogue warp has copied its part to TMEM
  uint64_t* bar_parked = bars + 4;      // [3] tap k: x_raw / x_lo parked in TMEM (one arrival per epilogue warp; absent taps too)
  uint64_t* bar_g1 = bars + 7;          // H accumulator complete
  uint64_t* bar_h = bars + 8;           // h_hi / h_lo parked in TMEM
  uint64_t* bar_g2 = bars + 9;          // O accumulator complete
  uint64_t* bar_wd = bars + 10;         // dilated-conv weight images landed
  uint64_t* bar_w1 = bars + 11;         // 1x1 weight images landed
  uint64_t* bar_sh = bars + 12;         // h staged (one arrival per epilogue warp) -> store warp
  uint64_t* bar_sy = bars + 13;         // y staged -> store warp
  uint64_t* bar_stfree = bars + 14;     // the staging buffer has been read by its TMA store
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  int wstep = 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_y);
    if (a.h != nullptr) tma_prefetch_desc(&tm_h);
    for (int k = 0; k < 2; ++k) { mbar_init(bar_full + k, 1); mbar_init(bar_sfree + k, kEpiWarps); }
    for (int k = 0; k < 3; ++k) mbar_init(bar_parked + k, kEpiWarps);
    mbar_init(bar_g1, 1); mbar_init(bar_h, kEpiWarps); mbar_init(bar_g2, 1);
    mbar_init(bar_wd, 1); mbar_init(bar_w1, 1);
    mbar_init(bar_sh, kEpiWarps); mbar_init(bar_sy, kEpiWarps); mbar_init(bar_stfree, 1);
    fence_barrier_init();
    if (a.flags != nullptr) {                      // chain: the weights of this CTA's first compute task
      for (int task = blockIdx.x; task < a.num_tiles * a.nsteps; task += gridDim.x) {
        const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b)) continue;
        wstep = step;
        break;
      }
    }
    const float* wimg_p = a.wimg + (long long)(a.lyr0 + wstep * a.lyr_dir) * a.wimg_stride;
    mbar_arrive_expect_tx(bar_wd, 12 * kSubB);
    for (int i = 0; i < 12; ++i) bulk_load(smem + i * kSubB, wimg_p + i * (kSubB / 4), kSubB, bar_wd);
    mbar_arrive_expect_tx(bar_w1, 4 * kSubB);
    for (int i = 12; i < 16; ++i) bulk_load(smem + i * kSubB, wimg_p + i * (kSubB / 4), kSubB, bar_w1);
  }
  if (tid >= 64 && tid < 128) sBias[tid - 64] = __ldg(a.bd + tid - 64);
  else if (tid >= 128 && tid < 192) sBias[tid - 64] = __ldg(a.b1 + tid - 128);
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (a.flags_in == nullptr) pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);
  const int order[3] = {1, 0, 2};               // centre tap first
  uint8_t* const stage = smem + k2OffStage;
  const int ntasks = a.num_tiles * a.nsteps;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0, nuse = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
        const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
        const int lyr = a.lyr0 + step * a.lyr_dir;
        const int d = a.d_from_layer ? (1 << lyr) : a.d;
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b)) continue;
        // never more than one tile ahead of the epilogue warps: the previous task's centre tap is parked, so its
        // predecessor's GEMMs are complete and every parity wait below is at most one phase behind its barrier.  (A phase
        // that cannot complete before this thread issues the task -- waiting on an OLDER phase could miss it.)
        if (it >= 1) mbar_wait(bar_parked + 1, (it - 1) & 1);
        const bool new_w = step != wstep;
        if (new_w) {
          mbar_wait(bar_g1, (it - 1) & 1);           // the previous task's tap GEMM has read the Wd region
          const float* wp = a.wimg + (long long)lyr * a.wimg_stride;
          mbar_arrive_expect_tx(bar_wd, 12 * kSubB);
          for (int i = 0; i < 12; ++i) bulk_load(smem + i * kSubB, wp + i * (kSubB / 4), kSubB, bar_wd);
        }
        if ((a.flags != nullptr && step > 0) || a.flags_in != nullptr) {
          const int* fl = (step > 0 ? a.flags + (size_t)(step - 1) * a.num_tiles : a.flags_in) + b * a.tiles_per_video;
          int idx[6];
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            const int tf = t0 + (kk - 1) * d;
            const bool pr = (tf + TM - 1 >= 0) && (tf < a.T);
            const int lo_t = tf < 0 ? 0 : tf, hi_t = (tf + TM - 1 < a.T) ? tf + TM - 1 : a.T - 1;
            idx[2 * kk] = pr ? lo_t / TM : t0 / TM;
            idx[2 * kk + 1] = pr ? hi_t / TM : t0 / TM;
          }
          const long long tw0 = clock64();
          if (a.trace != nullptr) a.trace[8 * (size_t)task] = global_ns();
          while (true) {
            int ok = 1;
#pragma unroll
            for (int j = 0; j < 6; ++j) ok &= ld_flag(fl + idx[j]);
            if (ok) break;
            __nanosleep(40);
            if (clock64() - tw0 > 8000000000LL) trap_report(2, task, blockIdx.x);
          }
          if (a.trace != nullptr) a.trace[8 * (size_t)task + 1] = global_ns();
        }
#pragma unroll
        for (int oi = 0; oi < 3; ++oi) {
          const int k = order[oi];
          const int tf = t0 + (k - 1) * d;
          const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
          if (oi == 2 && new_w) {
            mbar_wait(bar_g2, (it - 1) & 1);         // ... and its 1x1 GEMM the W1 region
            const float* wp = a.wimg + (long long)lyr * a.wimg_stride;
            mbar_arrive_expect_tx(bar_w1, 4 * kSubB);
            for (int i = 12; i < 16; ++i) bulk_load(smem + i * kSubB, wp + i * (kSubB / 4), kSubB, bar_w1);
            wstep = step;
          }
          if (!present) continue;
          const uint32_t slot = nuse & 1;
          mbar_wait(bar_sfree + slot, ((nuse >> 1) & 1) ^ 1);
          uint8_t* dst = smem + kOffSlots + slot * kSlot;
          mbar_arrive_expect_tx(bar_full + slot, kSlot);
          tma_load_4d(dst, &tm_x, bar_full + slot, 0, tf, b, lyr + a.cx_off);
          tma_load_4d(dst + kSubA, &tm_x, bar_full + slot, 32, tf, b, lyr + a.cx_off);
          if (a.trace != nullptr && oi == 0) a.trace[8 * (size_t)task + 4] = global_ns();
          ++nuse;
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    const uint32_t leader = lane == 0 ? 1u : 0u;
    const uint32_t usbase = __reduce_or_sync(0xffffffffu, sbase), utmem = __reduce_or_sync(0xffffffffu, tmem);
    const uint32_t wdh = umma_desc_lo(usbase + kOffWd);
    constexpr uint32_t idesc = umma_idesc_tf32(TM, 64), idesc2 = umma_idesc_tf32(TM, 128);
    const uint32_t w1h = umma_desc_lo(usbase + kOffW1Hi), w1l = umma_desc_lo(usbase + kOffW1Lo);
    const uint32_t tH = utmem + k2ColH, tO = utmem + k2ColO, tXr = utmem + k2ColXraw, tXl = utmem + k2ColXlo, tHlo = utmem + k2ColHlo;
    uint32_t it = 0, wgen = 0, wph = 0;
    int cur_step = -1;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
      const int lyr = a.lyr0 + step * a.lyr_dir;
      const int d = a.d_from_layer ? (1 << lyr) : a.d;
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      if (t0 >= __ldg(a.lens + b)) continue;
      const uint32_t p = it & 1;
      const bool new_w = step != cur_step;
      if (new_w) { cur_step = step; wph = wgen & 1; ++wgen; mbar_wait(bar_wd, wph); }
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
        mbar_wait(bar_parked + k, p);
        if (a.trace != nullptr && oi == 0 && lane == 0) a.trace[8 * (size_t)task + 5] = global_ns();
        tc_fence_after_sync();
        if (present) {
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t wo = ((k * 2 + s) * 2 * kSubB + ks * 32) >> 4;
              const uint32_t ac = k * 64 + s * 32 + ks * 8;
              umma_tf32_ts(tH, tXr + ac, wdh + wo, idesc2, (oi | s | ks) != 0, leader);   // x_hi * [W_hi | W_lo]
              umma_tf32_ts(tH, tXl + ac, wdh + wo, idesc, 1, leader);                       // x_lo * W_hi
            }
        }
      }
      umma_commit(bar_g1, leader);
      mbar_wait(bar_h, p);
      if (new_w) mbar_wait(bar_w1, wph);
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
      ++it;
    }
    __syncwarp();
  } else if (warp == 2 + kEpiWarps) {
    // =============================== store warp ==================================
    uint32_t nst = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
      const int lyr = a.lyr0 + step * a.lyr_dir;
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      int* const flag = (a.flags != nullptr && (step + 1 < a.nsteps || a.publish_last)) ? a.flags + (size_t)step * a.num_tiles + tile : nullptr;
      long long* const tr = a.trace != nullptr ? a.trace + 8 * (size_t)task + 3 : nullptr;
      if (t0 >= __ldg(a.lens + b)) {          // padding tile: y = 0 (h is never read there)
        float* const yout = a.y + (long long)lyr * a.plane + (size_t)b * a.T * C;
        for (int i = lane; i < TM * 16; i += 32) {
          const int t = t0 + (i >> 4);
          if (t < a.T) reinterpret_cast<float4*>(yout + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (flag != nullptr) {
          __syncwarp();
          if (lane == 0) {
            fence_release_gpu();
            fence_proxy_async_all();
            st_flag(flag, 1);
            if (tr != nullptr) *tr = global_ns();
          }
        }
        continue;
      }
      if (lane == 0) {
        if (a.h != nullptr) {
          mbar_wait(bar_sh, nst & 1);
          tma_store_4d(&tm_h, stage, 0, t0, b, lyr + a.co0_off);
          tma_store_4d(&tm_h, stage + kSubA, 32, t0, b, lyr + a.co0_off);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(bar_stfree);
        }
        mbar_wait(bar_sy, nst & 1);
        tma_store_4d(&tm_y, stage, 0, t0, b, lyr + a.co1_off);
        tma_store_4d(&tm_y, stage + kSubA, 32, t0, b, lyr + a.co1_off);
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(bar_stfree);
        if (flag != nullptr) {
          bulk_wait0();
          fence_release_gpu();
          st_flag(flag, 1);
          if (tr != nullptr) *tr = global_ns();
        }
      }
      __syncwarp();
      ++nst;
    }
    if (lane == 0) bulk_wait0();
  } else {
    // =============================== epilogue warps ==============================
    const int q = warp & 3, s = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int etid = tid - 64;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 32);
    const float* biasd = sBias + s * 32;
    const float* bias1 = sBias + 64 + s * 32;
    uint32_t it = 0, nuse = 0, nsf = 0;                 // nsf: uses of the staging buffer so far
    int bstep = a.nsteps > 1 ? -1 : 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
      const int lyr = a.lyr0 + step * a.lyr_dir;
      const int d = a.d_from_layer ? (1 << lyr) : a.d;
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
      if (t0 >= len) continue;
      const uint32_t layer_id = a.layer_id + (uint32_t)lyr;
      if (step != bstep) {                    // chain: this layer's biases (every epilogue warp is past the previous task)
        named_bar_sync(6, 32 * kEpiWarps);
        if (etid < 64) sBias[etid] = __ldg(a.bd + (long long)lyr * a.bias_stride + etid);
        else if (etid < 128) sBias[etid] = __ldg(a.b1 + (long long)lyr * a.bias_stride + etid - 64);
        named_bar_sync(6, 32 * kEpiWarps);
        bstep = step;
      }
      const uint32_t p = it & 1;
      const int t = t0 + row;
      float xc[32];
      // ---- park the tap tiles in TMEM: raw words (x_hi to the tensor core) and x_lo = x - trunc_tf32(x) ----
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
        if (present) {
          const uint32_t slot = nuse & 1;
          mbar_wait(bar_full + slot, (nuse >> 1) & 1);
          const uint8_t* sub = smem + kOffSlots + slot * kSlot + s * kSubA;
          uint32_t raw[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(sub + sw128_off(row, c));
            raw[4 * c + 0] = __float_as_uint(v.x); raw[4 * c + 1] = __float_as_uint(v.y);
            raw[4 * c + 2] = __float_as_uint(v.z); raw[4 * c + 3] = __float_as_uint(v.w);
            lo[4 * c + 0] = lo_bits(v.x); lo[4 * c + 1] = lo_bits(v.y);
            lo[4 * c + 2] = lo_bits(v.z); lo[4 * c + 3] = lo_bits(v.w);
            if (k == 1) { xc[4 * c] = v.x; xc[4 * c + 1] = v.y; xc[4 * c + 2] = v.z; xc[4 * c + 3] = v.w; }
          }
          tmem_st32(trow + k2ColXraw + k * 64, raw);
          tmem_st32(trow + k2ColXlo + k * 64, lo);
          tmem_wait_st();
          ++nuse;
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) { mbar_arrive(bar_sfree + slot); mbar_arrive(bar_parked + k); }
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_parked + k);
        }
        if (a.trace != nullptr && oi == 2 && etid == 0) a.trace[8 * (size_t)task + 6] = global_ns();
      }
      // ---- EPI1: H -> +bd, relu -> h ; h_hi / h_lo back into TMEM as the A operand of the 1x1 ----
      mbar_wait(bar_g1, p);
      if (a.trace != nullptr && etid == 0) a.trace[8 * (size_t)task + 2] = global_ns();
      tc_fence_after_sync();
      {
        uint32_t v[32], lo[32];
        {
          uint32_t w[32];
          tmem_ld32(trow + k2ColH, v);
          tmem_ld32(trow + k2ColH + 64, w);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float hv = fmaxf(__uint_as_float(v[i]) + biasd[i], 0.f);
          v[i] = __float_as_uint(hv);
          lo[i] = lo_bits(hv);
        }
        tmem_st32(trow + k2ColH, v);
        tmem_st32(trow + k2ColHlo, lo);
        if (a.h != nullptr) {
          mbar_wait(bar_stfree, (nsf & 1) ^ 1);        // the previous tile's y has left the staging buffer
          ++nsf;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stage + s * kSubA + sw128_off(row, c)) =
                make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                            __uint_as_float(v[4 * c + 3]));
        }
        tmem_wait_st();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h);
      if (a.h != nullptr) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sh);
      }
      // ---- EPI2: O -> +b1, dropout, residual, mask -> y ----
      uint32_t keep = 0xffffffffu;
      if (a.train) {
        const uint2 bits = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), layer_id, a.frame0 + (uint32_t)(b * a.T + t));
        keep = s == 0 ? bits.x : bits.y;
      }
      const float m = (t < len) ? 1.f : 0.f;
      const float on = a.train ? 2.f * m : m;
      mbar_wait(bar_g2, p);
      if (a.trace != nullptr && etid == 0) a.trace[8 * (size_t)task + 7] = global_ns();
      tc_fence_after_sync();
      {
        uint32_t v[32];
        tmem_ld32(trow + k2ColO, v);
        tmem_wait_ld();
        mbar_wait(bar_stfree, (nsf & 1) ^ 1);          // h (or the previous tile's y) has left the staging buffer
        ++nsf;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * c + j;
            const float ov = __uint_as_float(v[i]) + bias1[i];
            o[j] = xc[i] * m + (((keep >> i) & 1u) ? ov * on : 0.f);
          }
          *reinterpret_cast<float4*>(stage + s * kSubA + sw128_off(row, c)) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sy);
      ++it;
    }
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
  if (a.flags_in != nullptr && tid == 0) pdl_wait();      // never complete before the predecessor grid (see tc_layer_kernel)
}

}  // namespace tc
}  // namespace mstcn
