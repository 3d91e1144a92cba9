// Tensor-core path (tcgen05 + TMEM + TMA) for the fused dilated residual layer, sm_100a only.
//
// Precision: the reference's gradients are only reproducible with fp32-equivalent arithmetic (a
// plain TF32 forward flips max-over-stages winners and moves late-stage gradients by several
// percent -- DESIGN.md "precision").  Every GEMM here is therefore error-compensated 3xTF32:
//     x * W  ~=  x_hi*W_hi + x_hi*W_lo + x_lo*W_hi
// x_hi = trunc_tf32(x) is what the tensor core sees when handed the raw fp32 word (low 13 mantissa
// bits ignored); x_lo = rna_tf32(x - x_hi) is produced in registers and parked in TMEM as the
// A operand of the third product; W_hi = rna_tf32(W), W_lo = rna_tf32(W - W_hi) are prepared once
// per optimizer step by tc_pack_all_kernel.  Accumulation is fp32 in TMEM.
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
// per-layer weight image (floats): [Wd: 6 x (hi sub | lo sub) | W1_hi 2 sub | W1_lo 2 sub].  The hi and lo sub-tiles of a
// Wd K-block are adjacent so that one N=128 MMA multiplies x_hi by [W_hi | W_lo] (SS-form tf32 MMAs are bound by the
// shared-memory operand reads: 6 KB per m128 n64 k8 vs 8 KB for twice the work at n128)
constexpr int kWimgFloats = (6 + 6 + 2 + 2) * (kSubB / 4);          // 32768 floats = 128 KB
constexpr int kOffWd = 0, kOffW1Hi = 12 * kSubB, kOffW1Lo = 14 * kSubB;
constexpr int kOffSlots = 16 * kSubB;                                // 131072
constexpr int kOffBias = kOffSlots + 3 * kSlot;                      // 229376
constexpr int kOffBars = kOffBias + 2 * 64 * 4;                      // 229888
constexpr int kNumBars = 20;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kTcFwdSmem = kOffTmemPtr + 16 + 1024;                  // + slack to 1024-align the base
// TMEM columns
// H is 128 columns wide: [x*W_hi (+ x_lo*W_hi) | x_hi*W_lo], summed by the epilogue
constexpr uint32_t kColAlo = 0, kColH = 192, kColHlo = 320, kColO = 384, kTmemCols = 512;

// element (n = output row, kk = K index) of a K-major SWIZZLE_128B operand image made of [rows x 32] sub-tiles
__host__ __device__ inline int wimg_index(int n, int kk, int rows) {
  const int sub = kk >> 5, k32 = kk & 31;
  return sub * rows * 32 + (n >> 3) * 256 + (n & 7) * 32 + ((((k32 >> 2) ^ (n & 7)) & 7) << 2) + (k32 & 3);
}
// same inside the Wd part of a layer image: K-block kk/32 holds its hi sub-tile, then its lo sub-tile (+ kSubB/4 floats)
__host__ __device__ inline int wd_index(int n, int kk) { return (kk >> 5) * (2 * kSubB / 4) + wimg_index(n, kk & 31, 64); }

struct TcLayerFwdArgs {
  const int* lens; const float* wimg; const float* bd; const float* b1;
  float* y; float* h;
  int B, T, d, tiles_per_video, num_tiles;      // d < 0 in backward-gx mode (taps at t -/+ d swap roles)
  int skip_extra;                               // tiles starting at or after len + skip_extra are all zero
  int train; uint32_t layer_id; uint64_t seed, offset;
  const unsigned long long* offset_dev;   // optional device-side step counter added to `offset` (CUDA-graph replay)
  uint32_t frame0;                        // global index of this launch's first frame (video-group launches keep the
                                          // whole-batch frame numbering of the Philox stream)
  long long* dbg;     // optional: SM-clock timestamps of CTA 0's first tile (mstcn_debug_tc_timing)
  // MODE 2 only (see below): gy of this layer and relu output of the NEXT-LOWER layer, read straight from
  // global by the epilogue threads; wimg2 = that layer's backward image (its 1x1 part is used)
  const float* gyp; const float* hprev; const float* wimg2;
  // MODE 4, optional fusion of the top layer's pre-activation gradient (what tc_bwd_gu_kernel computes) behind the tail:
  //   gu(L-1) = (W1(L-1)^T (ga * mask * dropout(L-1))) * [h(L-1) > 0]   -> gu_out (B*T, 64)
  // hprev = h(L-1) (read straight from global by the epilogue threads: only its sign is needed), wimg2 = layer L-1's backward
  // image (its 1x1 part goes to weight sub-tiles 8..11), layer_id / seed / offset = layer L-1's dropout stream
  float* gu_out;
  // MODE 3 (stage tail forward): per-stage masked logits (B*T, K) row-major, class count
  float* logits_out; int K;
  // Chain launches (MODE 0 / 2): nsteps consecutive layers in ONE persistent launch.  Task = (step, tile) in
  // step-major order, dealt round-robin to the CTAs; a task starts when the flags of the previous step's tiles
  // under its three taps are set (dataflow instead of a kernel boundary per layer).  Step j works on layer
  // lyr = lyr0 + j*lyr_dir: dilation +-(1 << lyr) when d_from_layer, tensor-map layer coordinates lyr + c*_off,
  // outputs y/h + lyr*plane, weight images wimg/wimg2 + lyr*wimg_stride, biases bd/b1 + lyr*bias_stride,
  // dropout id layer_id + lyr.  nsteps == 1 is the plain single-layer launch (all of these 0 / NULL).
  int nsteps, lyr0, lyr_dir, d_from_layer, cx_off, cg_off, chp_off;
  int co0_off, co1_off;   // modes 0 / 2: layer-coordinate offsets of the two output tensor maps (see the store warp)
  long long plane, wimg_stride, bias_stride;
  int* flags;          // [nsteps][num_tiles], zeroed before the launch
  // Kernel-to-kernel dataflow (training forward): flags_in = one row of num_tiles flags published by the PREVIOUS kernel
  // for the tiles this launch reads first (chain step 0 / the tail's input); with it the launch skips griddepcontrol.wait
  // and starts on tiles as they are published, overlapping the previous kernel's drain.  publish_last: a chain also
  // publishes its last step (row nsteps-1); the tail (mode 3) publishes into flags[0..num_tiles).
  const int* flags_in; int publish_last;
  long long* trace;    // optional [nsteps*num_tiles][8] %globaltimer stamps per task: poll start, deps satisfied, GEMM1 done, published,
                       // first TMA issued, centre tap landed (MMA warp), x_lo of all taps parked, GEMM2 done
};

// relaxed gpu-scope flag accesses for the chain launches' tile dependencies
__device__ __forceinline__ int ld_flag(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// The release fence in front of a tile flag.  __threadfence() is fence.sc.gpu (SASS: MEMBAR.SC.GPU + ERRBAR + CCTL.IVALL);
// the flag protocol only needs release semantics at gpu scope (fence.acq_rel.gpu = MEMBAR.ALL.GPU).
#ifndef MSTCN_FENCE_MODE
#define MSTCN_FENCE_MODE 0
#endif
// Reader side of the tile-flag protocol, after the flags were seen set and before the TMA loads of the tiles:
//   0  nothing (round 1: "the tiles are only read through the async proxy from L2")
//   1  fence.proxy.async            2  fence.acq_rel.gpu + fence.proxy.async
#ifndef MSTCN_READER_FENCE
#define MSTCN_READER_FENCE 0
#endif
// Writer side, TMA-store path: 1 = a proxy fence between cp.async.bulk.wait_group and the gpu-scope release fence
#ifndef MSTCN_WRITER_PROXY_FENCE
#define MSTCN_WRITER_PROXY_FENCE 0
#endif
// experiment: wait this long after the flags were seen before touching the tiles (separates "the data lags its flag"
// from a protocol error)
// back-off between two polls of a task's dependency flags (chain launches).  Measured (profiles/r02_notes.md): polling is not
// free -- three poll batches in flight cost 5 % of the config-2 step, 500 ns instead of 40 ns of back-off gains 0.3-0.6 % there
// and 2.6 % at config 4 (one long video: many CTAs waiting); 1000 ns and a long first back-off lose again
#ifndef MSTCN_POLL_SLEEP_NS
#define MSTCN_POLL_SLEEP_NS 500
#endif
// back-off after the FIRST unsuccessful poll (a task whose dependencies are not there yet typically waits microseconds)
#ifndef MSTCN_POLL_FIRST_NS
#define MSTCN_POLL_FIRST_NS MSTCN_POLL_SLEEP_NS
#endif
#ifndef MSTCN_POLL_DELAY_NS
#define MSTCN_POLL_DELAY_NS 0
#endif
// Backward chain (MODE 2): the relu output h(l-1) of the FORWARD pass is the one operand of a backward step that comes from
// DRAM (0.8 GB of saved planes have streamed through L2 since it was written); its TMA load can only be issued late in the
// tile (the centre slot is recycled twice), so the producer warms L2 with it at the top of the task, before it polls the
// flags.  Same for the tail backward's q = softmax(z)*mask tile.  1 = on (default; -0.25 % of the config-2 step, same box A/B).
#ifndef MSTCN_PREFETCH_HP
#define MSTCN_PREFETCH_HP 1
#endif
// Which tiles beyond a video's end a backward step computes.  1 (rounds 1-2): every tile that starts before len + d, in
// MODE 1 and MODE 2 alike.  0: MODE 2 only computes tiles that hold valid frames.  gx(l) is non-zero on [len, len + d), but
// for l >= 1 nothing reads it there: the next step takes gy(l-1) = gx(l) times the mask, go(l-1) = gx(l)*mask*dropout, and
// the weight-gradient kernel masks gy as well; gu(l-1) vanishes beyond len either way.  Only layer 0's gx feeds an UNMASKED
// convolution (the stage's input projection), and that is MODE 1, which keeps the rule.  Tiles that are
// not computed are zero-filled and published by the store warp as before.  Default 0: at config 2 the steps of layers 9 / 8 / 7
// lose 22 / 13 / 7 of their 192 / 183 / 177 tiles (-0.45 % of the step; with the prefetch above -0.6 %, profiles/r02_notes.md).
#ifndef MSTCN_BWD_SKIP_EXTRA
#define MSTCN_BWD_SKIP_EXTRA 0
#endif
template <int MODE>
__device__ __forceinline__ int tile_skip_extra(int d) {
  const int ad = d < 0 ? -d : d;
  if (MODE == 1) return ad;
  if (MODE == 2) return MSTCN_BWD_SKIP_EXTRA ? ad : 0;
  return 0;
}
__device__ __forceinline__ void fence_after_flags_seen() {
#if MSTCN_POLL_DELAY_NS > 0
  __nanosleep(MSTCN_POLL_DELAY_NS);
#endif
#if MSTCN_READER_FENCE == 2
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
#if MSTCN_READER_FENCE >= 1
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_release_gpu() {
#if MSTCN_FENCE_MODE == 1
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
  __threadfence();
#endif
}
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

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

// the tap GEMM's accumulator: columns [0,64) hold x*W_hi (+ x_lo*W_hi), [64,128) hold x_hi*W_lo
__device__ __forceinline__ void tmem_ld_h(uint32_t trow, uint32_t (&v)[32]) {
  uint32_t w[32];
  tmem_ld32(trow + kColH, v);
  tmem_ld32(trow + kColH + 64, w);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
}

constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 64 + 32 * kEpiWarps;     // 320
constexpr int kTcLayerThreads = kTcThreads + 32;    // tc_layer_kernel: + one store warp

// MODE 0: DilatedResidualLayer.forward.
// MODE 1: the input-gradient half of its backward, gx[t] = gy[t]*mask + sum_k Wd[:,:,k]^T gu[t-(k-1)d]:
//         same tap GEMM on gu with the transposed weight image (a.d = -dilation), no 1x1, and an epilogue
//         that adds the residual-branch gradient.  tm_x maps gu, tm_g maps gy, a.y receives gx.
// MODE 2: layer l's input gradient fused with layer l-1's pre-activation gradient -- the backward mirror of
//         the forward kernel, one kernel per layer on the critical path instead of two:
//           gx(l)   = gy(l)*mask + sum_k Wd(l)[:,:,k]^T gu(l)[t-(k-1)d]          -> a.h   (GEMM1, EPI1)
//           gu(l-1) = (W1(l-1)^T (gx(l)*mask*dropout(l-1))) * [h(l-1) > 0]        -> a.y   (GEMM2, EPI2)
//         tm_x maps gu(l); a.wimg = layer l's backward image, a.wimg2 = layer l-1's; layer_id = l-1's.
// MODE 3: the stage tail (networks.py:333, 314, 330) with the forward kernel's skeleton and no side taps:
//           z = (Wout a + bout) * mask                  -> a.logits_out (B*T, K)       (GEMM1, EPI1)
//           q = softmax_K(z) * mask                     -> a.h (B*T, 64; kept for the backward)
//           x0' = Wn q + bn  (next stage's unmasked 1x1) -> a.y                          (GEMM2, EPI2)
//         a.wimg = the stage's tail image (Wout in the centre-tap sub-tiles, Wn in the 1x1 part), a.bd = bout
//         (zero-padded to 64), a.b1 = bn; a.y == NULL for the last stage (no GEMM2).  The running max over
//         stages and the winner index are taken by stage_max_kernel from the per-stage logits.
// MODE 4: the stage tail backward with the fused backward kernel's skeleton and no side taps:
//           gq = Wn^T gin  (gin = gradient of the next stage's projection output; absent for the last stage)
//           gz = ( q * (gq - <gq, q>) + gr ) * mask,  q = softmax_K(z_s)*mask as stored by MODE 3,
//                gr = [winner == s] * dL/dout (route_grad_kernel)  -> a.h (B*T, 64; feeds the tail weight gradients)
//           ga = Wout^T gz                                        -> a.y                  (GEMM2, EPI2)
//         tm_x maps gin (absent: a.gyp == NULL), tm_g maps q_s, tm_hp maps gr_s (all (B*T, 64) planes);
//         a.wimg = the stage's backward tail image (both its parts).
// All epilogue threads have issued the tile's global stores: one thread drains them to gpu scope (for the generic and
// the async proxy -- the consumers read through TMA) and sets the tile's flag; the other warps go on meanwhile.
__device__ __forceinline__ void publish_tile(int* flag, int etid) {
  named_bar_sync(6, 32 * kEpiWarps);
  if (etid == 0) {
    fence_release_gpu();
    fence_proxy_async_all();
    st_flag(flag, 1);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kTcLayerThreads, 1)
tc_layer_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                const __grid_constant__ CUtensorMap tm_hp, TcLayerFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // aligned up by an offset added to the __shared__ array itself, so that the compiler keeps the address space
  // (pointer <- integer casts made every access below a generic LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
  // MODE 2 recycles the centre tap slot twice per tile: x tap -> gy tile (for EPI1) -> h(l-1) tile (for EPI2)
  uint64_t* bar_c1 = bars + 14;         // the MMAs that read the centre tap from smem are complete
  uint64_t* bar_gy = bars + 15;         // gy tile landed in the centre slot
  uint64_t* bar_gyfree = bars + 16;     // EPI1 has consumed gy (one arrival per epilogue warp)
  uint64_t* bar_hp = bars + 17;         // h(l-1) tile landed in the centre slot
  uint64_t* bar_s0 = bars + 18;         // modes 0 / 2: slot-0 staging complete (one arrival per epilogue warp) -> store warp
  uint64_t* bar_s2 = bars + 19;         // same for the slot-2 staging
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) TC_STAMP(0);

  // ---- one-time setup (nothing here depends on the previous kernel: it overlaps that kernel's tail
  //      under programmatic dependent launch) ----
  int wstep = 0;                                // producer thread: step whose weight image is resident / in flight
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int k = 0; k < 3; ++k) { mbar_init(bar_full + k, 1); mbar_init(bar_lo + k, kEpiWarps); }
    mbar_init(bar_g1, 1); mbar_init(bar_h, kEpiWarps); mbar_init(bar_g2, 1);
    mbar_init(bar_free + 0, kEpiWarps); mbar_init(bar_free + 1, (MODE == 2 || MODE == 4) ? kEpiWarps : 1); mbar_init(bar_free + 2, kEpiWarps);
    mbar_init(bar_wd, 1); mbar_init(bar_w1, 1);
    mbar_init(bar_c1, 1); mbar_init(bar_gy, 1); mbar_init(bar_gyfree, kEpiWarps); mbar_init(bar_hp, 1);
    mbar_init(bar_s0, kEpiWarps); mbar_init(bar_s2, kEpiWarps);
    if (MODE == 2 || MODE == 4) { tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_hp); }
    fence_barrier_init();
    // chain launch: the weights this CTA needs first are those of its first compute task's step
    if (a.flags != nullptr) {
      for (int task = blockIdx.x; task < a.num_tiles * a.nsteps; task += gridDim.x) {
        const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
        const int lyr = a.lyr0 + step * a.lyr_dir;
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b) + tile_skip_extra<MODE>(a.d_from_layer ? (1 << lyr) : a.d)) continue;
        wstep = step;
        break;
      }
    }
    const int lyr_w = a.lyr0 + wstep * a.lyr_dir;
    const float* wimg_p = a.wimg + (long long)lyr_w * a.wimg_stride;
    // 128 KB operand image: 12 + 4 bulk copies of one 8 KB sub-tile each (async proxy -> no proxy fence)
    if (MODE == 3 || MODE == 4) {                // only the centre tap exists: K-blocks 2,3 = sub-tiles 4..7 (hi, lo, hi, lo)
      mbar_arrive_expect_tx(bar_wd, 4 * kSubB);
      for (int i = 4; i < 8; ++i) bulk_load(smem + i * kSubB, wimg_p + i * (kSubB / 4), kSubB, bar_wd);
    } else {
      mbar_arrive_expect_tx(bar_wd, 12 * kSubB);
      for (int i = 0; i < 12; ++i) bulk_load(smem + i * kSubB, wimg_p + i * (kSubB / 4), kSubB, bar_wd);
    }
    if (MODE != 1) {
      const float* w1src = MODE == 2 ? a.wimg2 + (long long)lyr_w * a.wimg_stride : wimg_p;
      const bool fuse_w = MODE == 4 && a.gu_out != nullptr;      // + the top layer's transposed 1x1 image for the fused gu GEMM
      mbar_arrive_expect_tx(bar_w1, (fuse_w ? 8 : 4) * kSubB);
      for (int i = 12; i < 16; ++i) bulk_load(smem + i * kSubB, w1src + i * (kSubB / 4), kSubB, bar_w1);
      if (fuse_w)
        for (int i = 12; i < 16; ++i) bulk_load(smem + (i - 4) * kSubB, a.wimg2 + i * (kSubB / 4), kSubB, bar_w1);
    } else {
      tma_prefetch_desc(&tm_g);
    }
  }
  if (MODE == 0 || MODE == 3) {
    if (tid >= 64 && tid < 128) sBias[tid - 64] = __ldg(a.bd + tid - 64);
    else if (tid >= 128 && tid < 192) sBias[tid - 64] = (MODE == 3 && a.y == nullptr) ? 0.f : __ldg(a.b1 + tid - 128);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (tid == 0) TC_STAMP(1);
  // A head kernel (no tile flags to follow) waits for its predecessor grid BEFORE it lets its dependents launch: the
  // flag-linked kernels behind it skip griddepcontrol.wait, so this is what keeps every one of them -- and the operand
  // images they fetch in their prologues -- behind everything that preceded the head in the stream.
  if (a.flags_in == nullptr) pdl_wait();        // the previous kernel's activations are complete and visible (else: per-tile flags)
  pdl_launch_dependents();                      // the next kernel's prologue may start as SMs free up
  if (tid == 0) TC_STAMP(2);
  const uint32_t tmem = *tmem_ptr;
  constexpr uint32_t idesc = umma_idesc_tf32(TM, 64);
  const uint32_t sbase = smem_u32(smem);
  const int order[3] = {1, 0, 2};               // centre tap first: it always exists and seeds the accumulator
  // Modes 0 / 2 hand their two output tiles to a store warp: staged in the TMA (SWIZZLE_128B) layout, written by
  // cp.async.bulk.tensor, and -- in a chain launch -- published by that warp, so the epilogue warps never wait for stores.
  constexpr bool kTmaOut = MODE == 0 || MODE == 2;
  uint8_t* const stage_h_ = smem + kOffSlots;               // tap-0 slot doubles as the first output's staging
  uint8_t* const stage_y_ = smem + kOffSlots + 2 * kSlot;   // tap-2 slot doubles as the second output's staging
  const bool has_in = MODE != 4 || a.gyp != nullptr;   // MODE 4, last stage: there is no next-stage gradient to pull back
  const bool fuse_gu = MODE == 4 && a.gu_out != nullptr;   // MODE 4: third GEMM + epilogue for the top layer's gu (see TcLayerFwdArgs)

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int task = blockIdx.x; task < a.num_tiles * a.nsteps; task += gridDim.x) {
        const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
        const int lyr = a.lyr0 + step * a.lyr_dir;
        const int d = a.d_from_layer ? (a.d < 0 ? -(1 << lyr) : (1 << lyr)) : a.d;
        const int skip_extra = tile_skip_extra<MODE>(d);
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b) + skip_extra) continue;
        const bool new_w = step != wstep;            // chain: this task needs another layer's weight image
        if (new_w) {
          mbar_wait(bar_g1, (it - 1) & 1);           // the previous task's tap GEMM has read the Wd region
          const float* wp = a.wimg + (long long)lyr * a.wimg_stride;
          mbar_arrive_expect_tx(bar_wd, 12 * kSubB);
          for (int i = 0; i < 12; ++i) bulk_load(smem + i * kSubB, wp + i * (kSubB / 4), kSubB, bar_wd);
        }
#if MSTCN_PREFETCH_HP
        if (MODE == 2) {                 // h(l-1) of the forward pass: from DRAM, needed two GEMMs from now
          tma_prefetch_4d(&tm_hp, 0, t0, b, lyr + a.chp_off);
          tma_prefetch_4d(&tm_hp, 32, t0, b, lyr + a.chp_off);
        }
        if (MODE == 4 && a.gyp != nullptr) {   // q of the forward pass, needed by the softmax backward right after the first GEMM
          tma_prefetch_4d(&tm_g, 0, t0, b, lyr + a.cg_off);
          tma_prefetch_4d(&tm_g, 32, t0, b, lyr + a.cg_off);
        }
#endif
        if ((a.flags != nullptr && step > 0) || a.flags_in != nullptr) {
          // dataflow dependency: the previous step's (or previous kernel's) tiles under the three taps (<= 2 tiles per
          // tap) are complete
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
          bool first_poll = true;
          while (true) {
            int ok = 1;
#pragma unroll
            for (int j = 0; j < 6; ++j) ok &= ld_flag(fl + idx[j]);
            if (ok) break;
            __nanosleep(first_poll ? MSTCN_POLL_FIRST_NS : MSTCN_POLL_SLEEP_NS);
            first_poll = false;
            if (clock64() - tw0 > 8000000000LL) trap_report(2, task, blockIdx.x);
          }
          fence_after_flags_seen();
          if (a.trace != nullptr) a.trace[8 * (size_t)task + 1] = global_ns();
          // The tiles are read through TMA only (async proxy, served by L2, issued after the flags were seen set); the
          // publishing thread drained the writers' stores to gpu scope and fenced them for the async proxy before it set
          // the flag (publish_tile).  A reader-side MEMBAR.GPU + proxy fence (~1 us on the critical path of every task)
          // would order nothing that is read here.
        }
#pragma unroll
        for (int oi = 0; oi < 3; ++oi) {
          const int k = order[oi];
          const int tf = t0 + (k - 1) * d;
          const bool present = (tf + TM - 1 >= 0) && (tf < a.T);
          const bool with_gy = MODE == 1 && oi == 2;      // gy rides on the last tap's barrier (freed last)
          if (oi == 2 && new_w) {
            if (MODE != 1) {
              mbar_wait(bar_g2, (it - 1) & 1);       // ... and its 1x1 GEMM the W1 region
              const float* wp = (MODE == 2 ? a.wimg2 : a.wimg) + (long long)lyr * a.wimg_stride;
              mbar_arrive_expect_tx(bar_w1, 4 * kSubB);
              for (int i = 12; i < 16; ++i) bulk_load(smem + i * kSubB, wp + i * (kSubB / 4), kSubB, bar_w1);
            }
            wstep = step;
          }
          mbar_wait(bar_free + k, (it & 1) ^ 1);
          if (MODE == 4) {
            // centre: the gin tile (absent for the last stage); tap-2 slot: the routed output gradient gr (B,T,K)
            if (k == 1 && a.gyp != nullptr) {
              mbar_arrive_expect_tx(bar_full + k, kSlot);
              uint8_t* dst = smem + kOffSlots + kSlot;
              tma_load_4d(dst, &tm_x, bar_full + k, 0, t0, b, lyr + a.cx_off);
              tma_load_4d(dst + kSubA, &tm_x, bar_full + k, 32, t0, b, lyr + a.cx_off);
            } else if (k == 2) {
              mbar_arrive_expect_tx(bar_full + k, kSlot);
              uint8_t* dst = smem + kOffSlots + 2 * kSlot;
              tma_load_4d(dst, &tm_hp, bar_full + k, 0, t0, b, lyr + a.chp_off);
              tma_load_4d(dst + kSubA, &tm_hp, bar_full + k, 32, t0, b, lyr + a.chp_off);
            } else {
              mbar_arrive(bar_full + k);
            }
            continue;
          }
          if (present || with_gy) {
            mbar_arrive_expect_tx(bar_full + k, (present ? kSlot : 0) + (with_gy ? kSlot : 0));
            if (present) {
              uint8_t* dst = smem + kOffSlots + k * kSlot;
              tma_load_4d(dst, &tm_x, bar_full + k, 0, tf, b, lyr + a.cx_off);
              tma_load_4d(dst + kSubA, &tm_x, bar_full + k, 32, tf, b, lyr + a.cx_off);
            }
            if (with_gy) {
              tma_load_4d(smem + kOffW1Hi, &tm_g, bar_full + k, 0, t0, b, lyr + a.cg_off);
              tma_load_4d(smem + kOffW1Hi + kSubA, &tm_g, bar_full + k, 32, t0, b, lyr + a.cg_off);
            }
            if (it == 0 && oi == 0) TC_STAMP(3);
            if (a.trace != nullptr && oi == 0) a.trace[8 * (size_t)task + 4] = global_ns();
          } else {
            mbar_arrive(bar_full + k);          // keep the phase in step; the tap contributes exactly 0
          }
        }
        if (MODE == 4 && a.gyp != nullptr) {
          // the centre slot is recycled for the stage's q = softmax(z)*mask tile, needed by the softmax backward
          uint8_t* c1 = smem + kOffSlots + kSlot;
          mbar_wait(bar_c1, it & 1);
          mbar_wait(bar_lo + 1, it & 1);
          mbar_arrive_expect_tx(bar_gy, kSlot);
          tma_load_4d(c1, &tm_g, bar_gy, 0, t0, b, lyr + a.cg_off);
          tma_load_4d(c1 + kSubA, &tm_g, bar_gy, 32, t0, b, lyr + a.cg_off);
        }
        if (MODE == 2) {
          uint8_t* c1 = smem + kOffSlots + kSlot;
          mbar_wait(bar_c1, it & 1);            // tensor core done with the centre tap ...
          mbar_wait(bar_lo + 1, it & 1);        // ... and so are the epilogue warps (x_lo parked)
          mbar_arrive_expect_tx(bar_gy, kSlot);
          tma_load_4d(c1, &tm_g, bar_gy, 0, t0, b, lyr + a.cg_off);
          tma_load_4d(c1 + kSubA, &tm_g, bar_gy, 32, t0, b, lyr + a.cg_off);
          mbar_wait(bar_gyfree, it & 1);
          mbar_arrive_expect_tx(bar_hp, kSlot);
          tma_load_4d(c1, &tm_hp, bar_hp, 0, t0, b, lyr + a.chp_off);
          tma_load_4d(c1 + kSubA, &tm_hp, bar_hp, 32, t0, b, lyr + a.chp_off);
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
    const uint32_t wdh = umma_desc_lo(usbase + kOffWd);
    constexpr uint32_t idesc2 = umma_idesc_tf32(TM, 128);         // x_hi * [W_hi | W_lo] in one instruction
    const uint32_t w1h = umma_desc_lo(usbase + kOffW1Hi), w1l = umma_desc_lo(usbase + kOffW1Lo);
    const uint32_t tH = utmem + kColH, tO = utmem + kColO, tAlo = utmem + kColAlo, tHlo = utmem + kColHlo;
    uint32_t it = 0, wgen = 0, wph = 0;
    int cur_step = -1;
    for (int task = blockIdx.x; task < a.num_tiles * a.nsteps; task += gridDim.x) {
      const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
      const int lyr = a.lyr0 + step * a.lyr_dir;
      const int d = a.d_from_layer ? (a.d < 0 ? -(1 << lyr) : (1 << lyr)) : a.d;
      const int skip_extra = tile_skip_extra<MODE>(d);
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      if (t0 >= __ldg(a.lens + b) + skip_extra) continue;
      const uint32_t p = it & 1;
      const bool new_w = step != cur_step;         // first task, or a chain task of another layer: next weight generation
      if (new_w) { cur_step = step; wph = wgen & 1; ++wgen; mbar_wait(bar_wd, wph); if (it == 0 && lane == 0) TC_STAMP(4); }
      // x_hi * (W_hi + W_lo): needs only the TMA data.  Centre tap first (always present, seeds H).
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T) && has_in;
        mbar_wait(bar_full + k, p);
        if (it == 0 && oi == 0 && lane == 0) TC_STAMP(5);
        if (a.trace != nullptr && oi == 0 && lane == 0) a.trace[8 * (size_t)task + 5] = global_ns();
        tc_fence_after_sync();
        if (present) {
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t ad = a0 + ((k * kSlot + s * kSubA + ks * 32) >> 4);
              const uint32_t wo = ((k * 2 + s) * 2 * kSubB + ks * 32) >> 4;
              umma_tf32_ss(tH, ad, wdh + wo, idesc2, (oi | s | ks) != 0, leader);
            }
        }
        if ((MODE == 2 || MODE == 4) && oi == 0) umma_commit(bar_c1, leader);
      }
      // x_lo * W_hi: A operand from TMEM once the epilogue warps have parked it
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T) && has_in;
        mbar_wait(bar_lo + k, p);
        tc_fence_after_sync();
        if (present) {
#pragma unroll
          for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_tf32_ts(tH, tAlo + k * 64 + s * 32 + ks * 8, wdh + (((k * 2 + s) * 2 * kSubB + ks * 32) >> 4), idesc, 1, leader);
        }
      }
      umma_commit(bar_g1, leader);
      if (it == 0 && lane == 0) TC_STAMP(6);
      if (MODE == 1 || (MODE == 3 && a.y == nullptr)) {
        // No second GEMM: bar_h only says "every epilogue warp has read H out of TMEM".  Without this wait the NEXT tile's
        // first MMA (accumulate = 0) could overwrite H while a slow epilogue warp was still loading it -- the centre slot is
        // handed back right after bar_g1, so the next tap can land and be multiplied within a microsecond (whole tiles of
        // the last stage's logits / of layer 0's gx came out wrong in ~0.5 % of the steps once programmatic launches made
        // all CTAs start their tiles at the same instant; tools/locate_race.py).
        mbar_wait(bar_h, p);
        ++it;
        continue;
      }
      // fused gu (MODE 4): bar_h and bar_g2 complete TWICE per tile, so their parities are 0 then 1 in every tile
      mbar_wait(bar_h, fuse_gu ? 0u : p);
      if (it == 0 && lane == 0) TC_STAMP(7);
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
      if (it == 0 && lane == 0) TC_STAMP(8);
      if (MODE == 4 && fuse_gu) {
        // third GEMM: gh = W1(L-1)^T go, go = ga*mask*dropout parked as hi / lo in the H / Hlo columns by EPI2 (which has
        // read ga out of O before it arrives on bar_h the second time)
        const uint32_t w3h = umma_desc_lo(usbase + 8 * kSubB), w3l = umma_desc_lo(usbase + 10 * kSubB);
        mbar_wait(bar_h, 1);
        tc_fence_after_sync();
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t wo = (s * kSubB + ks * 32) >> 4;
            const uint32_t ah = tH + s * 32 + ks * 8, al = tHlo + s * 32 + ks * 8;
            umma_tf32_ts(tO, ah, w3h + wo, idesc, (s | ks) != 0, leader);
            umma_tf32_ts(tO, ah, w3l + wo, idesc, 1, leader);
            umma_tf32_ts(tO, al, w3h + wo, idesc, 1, leader);
          }
        umma_commit(bar_g2, leader);
      }
      ++it;
    }
    __syncwarp();
  } else if (warp == 2 + kEpiWarps) {
    // =============================== store warp (modes 0 / 2) ===================
    // Output tiles staged by the epilogue warps leave through TMA stores; padding tiles are zero-filled here; in a
    // chain launch this warp publishes the tile once its stores are complete.
    if (kTmaOut) {
      const CUtensorMap* tm_o0 = MODE == 0 ? &tm_hp : &tm_g;     // mode 0: h planes | mode 2: Gl planes (gx)
      const CUtensorMap* tm_o1 = MODE == 0 ? &tm_g : &tm_x;      // mode 0: y planes | mode 2: U planes (gu of the layer below)
      uint32_t it = 0;
      for (int task = blockIdx.x; task < a.num_tiles * a.nsteps; task += gridDim.x) {
        const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
        const int lyr = a.lyr0 + step * a.lyr_dir;
        const int d = a.d_from_layer ? (a.d < 0 ? -(1 << lyr) : (1 << lyr)) : a.d;
        const int skip_extra = tile_skip_extra<MODE>(d);
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        int* const flag = (a.flags != nullptr && (step + 1 < a.nsteps || a.publish_last)) ? a.flags + (size_t)step * a.num_tiles + tile : nullptr;
        long long* const tr = a.trace != nullptr ? a.trace + 8 * (size_t)task + 3 : nullptr;
        if (t0 >= __ldg(a.lens + b) + skip_extra) {
          float* const yout = a.y + (long long)lyr * a.plane + (size_t)b * a.T * C;
          float* const hout = a.h + (long long)lyr * a.plane + (size_t)b * a.T * C;
          for (int i = lane; i < TM * 16; i += 32) {
            const int t = t0 + (i >> 4);
            if (t < a.T) {
              reinterpret_cast<float4*>(yout + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MODE == 2) reinterpret_cast<float4*>(hout + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
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
          const uint32_t p = it & 1;
          if (a.h != nullptr) {
            mbar_wait(bar_s0, p);
            tma_store_4d(tm_o0, stage_h_, 0, t0, b, lyr + a.co0_off);
            tma_store_4d(tm_o0, stage_h_ + kSubA, 32, t0, b, lyr + a.co0_off);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive_n(bar_free + 0, kEpiWarps);
          }
          mbar_wait(bar_s2, p);
          tma_store_4d(tm_o1, stage_y_, 0, t0, b, lyr + a.co1_off);
          tma_store_4d(tm_o1, stage_y_ + kSubA, 32, t0, b, lyr + a.co1_off);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive_n(bar_free + 2, kEpiWarps);
          if (flag != nullptr) {
            // wait_group alone is NOT enough: without the gpu-scope fence the flag (generic proxy) was observed ahead of
            // the tile's async-proxy writes on other SMs (the full-size determinism test caught it).  MEMBAR.GPU costs
            // ~0.7 us of every layer step's critical path; it is the price of the release.
            bulk_wait0();
#if MSTCN_WRITER_PROXY_FENCE
            fence_proxy_async_all();
#endif
            fence_release_gpu();
            st_flag(flag, 1);
            if (tr != nullptr) *tr = global_ns();
          }
          if (it == 0) TC_STAMP(20);
        }
        __syncwarp();
        ++it;
      }
      if (lane == 0) bulk_wait0();
    }
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
    int bstep = a.nsteps > 1 ? -1 : 0;                  // step whose biases sit in sBias
    for (int task = blockIdx.x; task < a.num_tiles * a.nsteps; task += gridDim.x) {
      const int step = task / a.num_tiles, tile = task - step * a.num_tiles;
      const int lyr = a.lyr0 + step * a.lyr_dir;
      const int d = a.d_from_layer ? (a.d < 0 ? -(1 << lyr) : (1 << lyr)) : a.d;
      const int skip_extra = tile_skip_extra<MODE>(d);
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
      const size_t vbase = (size_t)b * a.T * C;
      float* const yout = a.y ? a.y + (long long)lyr * a.plane : nullptr;
      float* const hout = a.h ? a.h + (long long)lyr * a.plane : nullptr;
      const uint32_t layer_id = a.layer_id + (uint32_t)lyr;
      if (MODE == 3 && t0 >= len) {
        // padding tile: z = 0 (mask), q = 0, and the next stage's unmasked 1x1 outputs its bias
        const int rows = (a.T - t0) < TM ? (a.T - t0) : TM;
        float* lg = a.logits_out + ((size_t)b * a.T + t0) * a.K;
        for (int i = etid; i < rows * a.K; i += 32 * kEpiWarps) lg[i] = 0.f;
        for (int i = etid; i < TM * 16; i += 32 * kEpiWarps) {
          const int t = t0 + (i >> 4);
          if (t < a.T) {
            if (a.h != nullptr) reinterpret_cast<float4*>(hout + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.y != nullptr)
              reinterpret_cast<float4*>(yout + vbase + (size_t)t * C)[i & 15] = *reinterpret_cast<const float4*>(sBias + 64 + 4 * (i & 15));
          }
        }
        if (a.flags != nullptr) publish_tile(a.flags + tile, etid);
        continue;
      }
      if (t0 >= len + skip_extra) {           // nothing but zeros reaches this tile: y = 0 (h is never read there)
        if (!kTmaOut) {                       // (modes 0 / 2: written and published by the store warp)
          for (int i = etid; i < TM * 16; i += 32 * kEpiWarps) {
            const int t = t0 + (i >> 4);
            if (t < a.T) {
              reinterpret_cast<float4*>(yout + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MODE == 4) reinterpret_cast<float4*>(hout + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (MODE == 4 && fuse_gu) reinterpret_cast<float4*>(a.gu_out + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (a.flags != nullptr) publish_tile(a.flags + tile, etid);
        }
        continue;
      }
      if (MODE == 0 && step != bstep) {       // chain: this layer's biases (every epilogue warp is past the previous task)
        named_bar_sync(6, 32 * kEpiWarps);
        if (etid < 64) sBias[etid] = __ldg(a.bd + (long long)lyr * a.bias_stride + etid);
        else if (etid < 128) sBias[etid] = __ldg(a.b1 + (long long)lyr * a.bias_stride + etid - 64);
        named_bar_sync(6, 32 * kEpiWarps);
        bstep = step;
      }
      const uint32_t p = it & 1;
      const int t = t0 + row;
      uint32_t hmask = 0;                       // fused gu (MODE 4): [h(L-1) > 0] of this thread's row and column half,
      if (MODE == 4 && fuse_gu && t < a.T) {    // fetched now, used two GEMMs later
        const float4* hp = reinterpret_cast<const float4*>(a.hprev + vbase + (size_t)t * C + s * 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 hv = __ldg(hp + c);
          hmask |= (hv.x > 0.f ? 1u : 0u) << (4 * c) | (hv.y > 0.f ? 1u : 0u) << (4 * c + 1) |
                   (hv.z > 0.f ? 1u : 0u) << (4 * c + 2) | (hv.w > 0.f ? 1u : 0u) << (4 * c + 3);
        }
      }
      float xc[32];
#pragma unroll
      for (int oi = 0; oi < 3; ++oi) {
        const int k = order[oi];
        const int tf = t0 + (k - 1) * d;
        const bool present = (tf + TM - 1 >= 0) && (tf < a.T) && has_in;
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
        if (a.trace != nullptr && oi == 2 && etid == 0) a.trace[8 * (size_t)task + 6] = global_ns();
      }
      // ---- EPI1: H -> +bd, relu -> h ; h_hi / h_lo back into TMEM as the A operand of the 1x1 ----
      mbar_wait(bar_g1, p);
      if (a.trace != nullptr && etid == 0) a.trace[8 * (size_t)task + 2] = global_ns();
      if (it == 0 && etid == 0) TC_STAMP(13);
      tc_fence_after_sync();
      if (MODE != 2 && MODE != 4 && etid == 0) mbar_arrive(bar_free + 1);       // centre slot: every MMA and epilogue read is done
      if (MODE == 1) {
        // gx = (W^T gu) + gy * mask ; gy was TMA-loaded into the (unused) 1x1-weight region
        const float m1 = (t < len) ? 1.f : 0.f;
        const uint8_t* gsub = smem + kOffW1Hi + s * kSubA;
        uint32_t v[32];
        tmem_ld_h(trow, v);
        tmem_wait_ld();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_h);              // H is in registers: the next tile's GEMM may overwrite it
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
        copy_out_rows(stage_h, yout + vbase, t0, a.T, q, s, lane);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 0);
        if (a.flags != nullptr) publish_tile(a.flags + tile, etid);      // the next stage's tail backward starts on this tile
        ++it;
        continue;
      }
      if (MODE == 4) {
        // gz = q * (gq - <gq, q>) + gr, masked: softmax backward from the stored q = softmax(z)*mask, plus the routed dL/dout
        const float m1 = (t < len) ? 1.f : 0.f;
        float* xch = reinterpret_cast<float*>(smem);             // [2 halves][128 rows]: weight sub-tile 0 is unused here
        const uint8_t* gsub = stage_y + s * kSubA;               // routed dL/dout tile (tap-2 slot)
        const uint8_t* qsub = smem + kOffSlots + kSlot + s * kSubA;   // q tile (recycled centre slot)
        float gz[32];
        if (has_in) {
          uint32_t v[32];
          tmem_ld_h(trow, v);
          tmem_wait_ld();
          mbar_wait(bar_gy, p);
          float pq[32], dot = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 qv = *reinterpret_cast<const float4*>(qsub + sw128_off(row, c));
            pq[4 * c] = qv.x; pq[4 * c + 1] = qv.y; pq[4 * c + 2] = qv.z; pq[4 * c + 3] = qv.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_free + 1);              // q consumed: the centre slot may take the next gin tile
#pragma unroll
          for (int i = 0; i < 32; ++i) dot += __uint_as_float(v[i]) * pq[i];
          xch[s * 128 + row] = dot;
          named_bar_sync(1 + q, 64);
          dot += xch[(1 - s) * 128 + row];
#pragma unroll
          for (int i = 0; i < 32; ++i) gz[i] = pq[i] * (__uint_as_float(v[i]) - dot);
          named_bar_sync(1 + q, 64);                             // xch is rewritten by the next tile
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) gz[i] = 0.f;
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_free + 1);
        }
        uint32_t v[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 gr = *reinterpret_cast<const float4*>(gsub + sw128_off(row, c));
          const float4 g = make_float4((gz[4 * c] + gr.x) * m1, (gz[4 * c + 1] + gr.y) * m1, (gz[4 * c + 2] + gr.z) * m1,
                                       (gz[4 * c + 3] + gr.w) * m1);
          v[4 * c] = __float_as_uint(g.x); v[4 * c + 1] = __float_as_uint(g.y);
          v[4 * c + 2] = __float_as_uint(g.z); v[4 * c + 3] = __float_as_uint(g.w);
          lo[4 * c] = lo_bits(g.x); lo[4 * c + 1] = lo_bits(g.y); lo[4 * c + 2] = lo_bits(g.z); lo[4 * c + 3] = lo_bits(g.w);
        }
        tmem_st32(trow + kColH, v);
        tmem_st32(trow + kColHlo, lo);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(stage_h + stage_off(row, s * 8 + c)) =
              make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                          __uint_as_float(v[4 * c + 3]));
        tmem_wait_st();
      } else if (MODE == 3) {
        // z = (acc + bout) * mask -> logits; q = softmax_K(z) * mask -> TMEM (A operand of the next stage's 1x1) and a.h
        const float m1 = (t < len) ? 1.f : 0.f;
        const int K = a.K;
        const bool has_next = a.y != nullptr;
        float* xch = reinterpret_cast<float*>(stage_y);          // [2 exchanges][2 halves][128 rows] in the idle tap-2 slot
        // logits staging in the idle tap-0 slot, ONE 8 KB region PER WARP PAIR ([32 rows][Kp] at byte 8192 * q): the same
        // bytes later hold this pair's 32 rows of the q staging, so every reuse of the region is ordered by the pair's own
        // barriers.  (A [128 rows][Kp] array packed the pairs 6272 bytes apart: pair q's q staging then overlapped pair
        // q+1's logits rows, which nothing ordered -- a late pair overwrote 8 frames of its neighbour's staged q.  The
        // forward never noticed, q goes on through TMEM; the backward read the damaged q plane in ~0.1 % of the steps.)
        float* lstage = reinterpret_cast<float*>(stage_h + 8192 * q);
        const int Kp = K < 64 ? (K | 1) : 64;                    // odd row stride: a warp's 32 rows hit 32 different banks (32 x Kp x 4 <= 8 KB)
        uint32_t v[32];
        tmem_ld_h(trow, v);
        tmem_wait_ld();
        if (!has_next) {                                // last stage: no GEMM2, bar_h = "H has been read" (see the MMA warp)
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_h);
        }
        float z[32];
        float zmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          z[i] = (__uint_as_float(v[i]) + biasd[i]) * m1;
          const int c = s * 32 + i;
          if (c < K) { lstage[lane * Kp + c] = z[i]; zmax = fmaxf(zmax, z[i]); }
        }
        xch[s * 128 + row] = zmax;
        named_bar_sync(1 + q, 64);                               // pair: logits staged, partial maxima exchanged
        {   // this pair's 32 rows of logits are one contiguous block of the (B*T, K) tensor
          const int rows_left = a.T - (t0 + 32 * q);
          const int nrow = rows_left < 32 ? (rows_left < 0 ? 0 : rows_left) : 32;
          float* dst = a.logits_out + ((size_t)b * a.T + t0 + 32 * q) * K;
          const float* src = lstage;
          const float invK = 1.f / (float)K;
          for (int i = s * 32 + lane; i < nrow * K; i += 64) {
            const int r = (int)(((float)i + 0.5f) * invK);         // i / K, exact for these small integers
            dst[i] = src[r * Kp + (i - r * K)];
          }
        }
        if (has_next) {
          zmax = fmaxf(zmax, xch[(1 - s) * 128 + row]);
          float e[32], sum = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) { e[i] = (s * 32 + i < K) ? expf(z[i] - zmax) : 0.f; sum += e[i]; }
          xch[256 + s * 128 + row] = sum;
          named_bar_sync(1 + q, 64);
          sum += xch[256 + (1 - s) * 128 + row];
          const float sc = m1 / sum;
          uint32_t lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) { const float qv = e[i] * sc; v[i] = __float_as_uint(qv); lo[i] = lo_bits(qv); }
          tmem_st32(trow + kColH, v);
          tmem_st32(trow + kColHlo, lo);
          named_bar_sync(1 + q, 64);                             // the pair is done with the logits staging rows
          if (a.h != nullptr) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<float4*>(stage_h + stage_off(row, s * 8 + c)) =
                  make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                              __uint_as_float(v[4 * c + 3]));
          }
          tmem_wait_st();
        }
      } else if (MODE == 2) {
        // gx(l) = acc + gy*mask -> staged for the store; go(l-1) = gx * mask * dropout(l-1) -> hi/lo back into TMEM
        const bool inb = t < a.T;
        const float m1 = (t < len) ? 1.f : 0.f;
        uint32_t keep = 0xffffffffu;
        if (a.train) {
          const uint2 bits = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), layer_id, a.frame0 + (uint32_t)(b * a.T + t));
          keep = s == 0 ? bits.x : bits.y;
        }
        const float on = a.train ? 2.f * m1 : m1;
        const uint8_t* gsub = smem + kOffSlots + kSlot + s * kSubA;     // gy tile, TMA-loaded into the centre slot
        (void)inb;
        uint32_t v[32], lo[32];
        tmem_ld_h(trow, v);
        tmem_wait_ld();
        mbar_wait(bar_gy, p);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 g = *reinterpret_cast<const float4*>(gsub + sw128_off(row, c));
          const float gx0 = __uint_as_float(v[4 * c]) + g.x * m1, gx1 = __uint_as_float(v[4 * c + 1]) + g.y * m1;
          const float gx2 = __uint_as_float(v[4 * c + 2]) + g.z * m1, gx3 = __uint_as_float(v[4 * c + 3]) + g.w * m1;
          *reinterpret_cast<float4*>(stage_h + s * kSubA + sw128_off(row, c)) = make_float4(gx0, gx1, gx2, gx3);
          const float go0 = ((keep >> (4 * c)) & 1u) ? gx0 * on : 0.f, go1 = ((keep >> (4 * c + 1)) & 1u) ? gx1 * on : 0.f;
          const float go2 = ((keep >> (4 * c + 2)) & 1u) ? gx2 * on : 0.f, go3 = ((keep >> (4 * c + 3)) & 1u) ? gx3 * on : 0.f;
          v[4 * c] = __float_as_uint(go0); v[4 * c + 1] = __float_as_uint(go1);
          v[4 * c + 2] = __float_as_uint(go2); v[4 * c + 3] = __float_as_uint(go3);
          lo[4 * c] = lo_bits(go0); lo[4 * c + 1] = lo_bits(go1); lo[4 * c + 2] = lo_bits(go2); lo[4 * c + 3] = lo_bits(go3);
        }
        tmem_st32(trow + kColH, v);
        tmem_st32(trow + kColHlo, lo);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gyfree);       // the centre slot may take the h(l-1) tile
        tmem_wait_st();
      } else {
        uint32_t v[32], lo[32];
        tmem_ld_h(trow, v);
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
            *reinterpret_cast<float4*>(stage_h + s * kSubA + sw128_off(row, c)) =
                make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                            __uint_as_float(v[4 * c + 3]));
        }
        tmem_wait_st();
      }
      if (MODE == 3 && a.y == nullptr) {               // last stage: no softmax, no next projection
        named_bar_sync(1 + q, 64);                     // the pair's logits rows are copied out before the next tile restages
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar_free + 0); mbar_arrive(bar_free + 2); }
        ++it;
        continue;
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_h);
      if (it == 0 && etid == 0) TC_STAMP(14);
      if (kTmaOut) {
        fence_proxy_async_smem();                     // staging (generic proxy) -> visible to the TMA store (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(hout != nullptr ? bar_s0 : bar_free + 0);   // no h output: the slot goes straight back
      } else {
        if (hout != nullptr) copy_out_rows(stage_h, hout + vbase, t0, a.T, q, s, lane);
        fence_proxy_async_smem();                     // staging (generic proxy) before the next TMA write (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 0);
      }
      if (MODE == 4) {
        // ---- EPI2 (MODE 4): ga = Wout^T gz ----
        mbar_wait(bar_g2, fuse_gu ? 0u : p);
        tc_fence_after_sync();
        uint32_t v[32];
        tmem_ld32(trow + kColO, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(stage_y + stage_off(row, s * 8 + c)) =
              make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                          __uint_as_float(v[4 * c + 3]));
        if (fuse_gu) {
          // go = ga * mask * dropout(L-1) (what tc_bwd_gu_kernel makes of the gy tile): hi / lo into the H / Hlo columns,
          // whose gz the second GEMM has finished reading (bar_g2).  O has been read: the third GEMM may overwrite it.
          uint32_t keep = 0xffffffffu;
          if (a.train) {
            const uint2 bits = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), layer_id, a.frame0 + (uint32_t)(b * a.T + t));
            keep = s == 0 ? bits.x : bits.y;
          }
          const float on = (t < len) ? (a.train ? 2.f : 1.f) : 0.f;
          uint32_t lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float g = ((keep >> i) & 1u) ? __uint_as_float(v[i]) * on : 0.f;
            v[i] = __float_as_uint(g);
            lo[i] = lo_bits(g);
          }
          tmem_st32(trow + kColH, v);
          tmem_st32(trow + kColHlo, lo);
          tmem_wait_st();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_h);            // second arrival of this tile (parity 1 at the MMA warp)
        }
        tc_fence_before_sync();
        copy_out_rows(stage_y, yout + vbase, t0, a.T, q, s, lane);
        if (fuse_gu) {
          // ---- EPI3: gu(L-1) = gh * [h(L-1) > 0], through the same staging rows ----
          named_bar_sync(1 + q, 64);                    // the pair has read its ga rows
          mbar_wait(bar_g2, 1);
          tc_fence_after_sync();
          tmem_ld32(trow + kColO, v);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stage_y + stage_off(row, s * 8 + c)) =
                make_float4(((hmask >> (4 * c)) & 1u) ? __uint_as_float(v[4 * c]) : 0.f,
                            ((hmask >> (4 * c + 1)) & 1u) ? __uint_as_float(v[4 * c + 1]) : 0.f,
                            ((hmask >> (4 * c + 2)) & 1u) ? __uint_as_float(v[4 * c + 2]) : 0.f,
                            ((hmask >> (4 * c + 3)) & 1u) ? __uint_as_float(v[4 * c + 3]) : 0.f);
          tc_fence_before_sync();
          copy_out_rows(stage_y, a.gu_out + vbase, t0, a.T, q, s, lane);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 2);
        if (a.flags != nullptr) publish_tile(a.flags + tile, etid);      // gz and ga of this tile are stored
        ++it;
        continue;
      }
      if (MODE == 3) {
        // ---- EPI2 (MODE 3): x0' = acc + bn (unmasked: padded frames carry the bias) ----
        mbar_wait(bar_g2, p);
        tc_fence_after_sync();
        uint32_t v[32];
        tmem_ld32(trow + kColO, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(stage_y + stage_off(row, s * 8 + c)) =
              make_float4(__uint_as_float(v[4 * c]) + bias1[4 * c], __uint_as_float(v[4 * c + 1]) + bias1[4 * c + 1],
                          __uint_as_float(v[4 * c + 2]) + bias1[4 * c + 2], __uint_as_float(v[4 * c + 3]) + bias1[4 * c + 3]);
        tc_fence_before_sync();
        copy_out_rows(stage_y, yout + vbase, t0, a.T, q, s, lane);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + 2);
        if (a.flags != nullptr) publish_tile(a.flags + tile, etid);    // the next stage's chain starts on this tile
        ++it;
        continue;
      }
      if (MODE == 2) {
        // ---- EPI2 (MODE 2): gu(l-1) = gh * [h(l-1) > 0] ----
        const uint8_t* hsub = smem + kOffSlots + kSlot + s * kSubA;     // h(l-1) tile in the centre slot
        mbar_wait(bar_hp, p);
        mbar_wait(bar_g2, p);
        tc_fence_after_sync();
        uint32_t v[32];
        tmem_ld32(trow + kColO, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 hv = *reinterpret_cast<const float4*>(hsub + sw128_off(row, c));
          *reinterpret_cast<float4*>(stage_y + s * kSubA + sw128_off(row, c)) =
              make_float4(hv.x > 0.f ? __uint_as_float(v[4 * c]) : 0.f, hv.y > 0.f ? __uint_as_float(v[4 * c + 1]) : 0.f,
                          hv.z > 0.f ? __uint_as_float(v[4 * c + 2]) : 0.f, hv.w > 0.f ? __uint_as_float(v[4 * c + 3]) : 0.f);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar_free + 1); mbar_arrive(bar_s2); }   // centre slot free; gu(l-1) tile to the store warp
        ++it;
        continue;
      }
      // ---- EPI2: O -> +b1, dropout, residual, mask -> y ----
      uint32_t keep = 0xffffffffu;
      if (a.train) {
        const uint2 bits = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), layer_id, a.frame0 + (uint32_t)(b * a.T + t));
        keep = s == 0 ? bits.x : bits.y;
      }
      const float m = (t < len) ? 1.f : 0.f;
      const float on = a.train ? 2.f * m : m;         // kept channels: scale 2 (p = 0.5), then the mask
      mbar_wait(bar_g2, p);
      if (it == 0 && etid == 0) TC_STAMP(15);
      if (a.trace != nullptr && etid == 0) a.trace[8 * (size_t)task + 7] = global_ns();
      tc_fence_after_sync();
      {
        uint32_t v[32];
        tmem_ld32(trow + kColO, v);
        tmem_wait_ld();
        if (it == 0 && etid == 0) TC_STAMP(18);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * c + j;
            const float ov = __uint_as_float(v[i]) + bias1[i];
            o[j] = xc[i] * m + (((keep >> i) & 1u) ? ov * on : 0.f);
          }
          *reinterpret_cast<float4*>(stage_y + s * kSubA + sw128_off(row, c)) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      if (it == 0 && etid == 0) TC_STAMP(19);
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s2);             // y tile to the store warp
      if (it == 0 && etid == 0) TC_STAMP(16);
      ++it;
    }
  }
  // ---- teardown ----
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) TC_STAMP(17);
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
  // A launch that followed tile flags instead of griddepcontrol.wait must still not COMPLETE before its predecessor grid:
  // everything behind it in the stream that is not flag-linked (events, plain launches, the weight-gradient stream) takes
  // this grid's completion for the completion of all earlier work.  Without this wait the predecessor's last stores
  // (e.g. the zero fill of far padding tiles, which no tile flag consumer waits for) could still be in flight when the
  // weight-gradient kernel read them: 1.4 % of the B=64 eager steps differed in a few gradients (tools/race_stress.py).
  if (a.flags_in != nullptr && tid == 0) pdl_wait();
}

// =============================================================================================
// Weight-gradient kernel: for every tap k of a tile,
//     D_k[o][c] += sum_t A_k[t][o] * B_k[t][c]           (reduction over the tile's 128 frames)
// as one UMMA chain with BOTH operands MN-major (channels are the contiguous dimension of the
// TMA-written tiles, frames are K).  Both operands enter exactly: A = [A_hi ; A_lo] stacked along M
// (64 channels of trunc_tf32 + 64 of the remainder) and B = [B_hi ; B_lo] stacked along N, so one
// m128 n128 k8 instruction produces all four partial products; the four 64x64 quadrants of the
// accumulator are added in the epilogue.  (A single-rounded B moved late-stage conv gradients by
// 4e-3 through cancellation -- not acceptable against the 1e-3 bar.)
// Accumulators stay in TMEM (4 taps x 128 columns = all 512) for the CTA's whole tile loop; one partial
// per CTA goes to scratch and reduce_partials_kernel sums the grid in fixed order.
// Taps 0..2: dWd[k], A = gu shifted by -(k-1)d, B = x (loaded once per tile);  tap 3: dW1,
// A = go = gy*mask*dropout (made in place by the transform warps), B = h.  Column sums of A give the
// bias gradients.
// Per-tile event order, identical in every role:  [B<-x] A0 A1 A2 [B<-h] A3  (absent taps dropped).
// =============================================================================================
struct TcWgradArgs {
  const int* lens; float* part;
  int B, T, d, tiles_per_video, num_tiles;
  // stage mode: the grid is nlayers x ctas_per_layer; CTA c works on layer c / ctas_per_layer (the 4th
  // coordinate of the tensor maps, dilation 1 << layer, dropout id layer0_id + layer) and strides over that
  // layer's tiles by ctas_per_layer.  Single-layer launches use nlayers = 1, ctas_per_layer = gridDim.x.
  int nlayers, ctas_per_layer, layer0_id, dil_from_layer;
  // tap_mask: which of the four taps exist; gy_transform: tap 3's A is gy and gets mask*dropout applied in place;
  // tap3_full_T: tap 3 also visits tiles beyond the video's length (its column sums feed an unmasked bias)
  int tap_mask, gy_transform, tap3_full_T;
  // stage mode may append tail_ctas CTAs for the 1x1 convolutions around the stage ("layer" nlayers: dilation 0, taps
  // tail_tap_mask, no gy transform, tap 3 over all frames; tap 3's B operand comes through tm_q); cg_off is added to the
  // layer coordinate of tm_gy for the real layers (its map starts one plane earlier so that the tail can reach Gl[0])
  int tail_ctas, tail_tap_mask, cg_off;
  // projection mode (the stage-1 input conv's weight gradient, networks.py:325,330): "layer" = a group of four 64-feature
  // chunks of the (B, T, dim) feature tensor, tap k = chunk 4*group + k.  A_k = that feature chunk (through tm_gu, a map over
  // the caller's features), B = the gradient of the projection output (tm_x), shared by the four taps:
  //   part[k][c'][o] = sum over ALL frames of feat[t][64*(4*group+k) + c'] * g0[t][o]  = dW[o][c]^T chunk,
  // and the bias slot of tap 0 holds sum_t g0[t][o] (the conv is unmasked: padded frames count).
  int proj, nchunks;
  long long* dbg;      // optional: CTA 0's per-role wait / work clock totals (slots 32..47 of the timing buffer)
  int train; uint32_t layer_id; uint64_t seed, offset;
  const unsigned long long* offset_dev;   // optional device-side step counter added to `offset` (CUDA-graph replay)
  uint32_t frame0;                        // global index of this launch's first frame (video-group launches keep the
                                          // whole-batch frame numbering of the Philox stream)
};
constexpr int TW = 64;                                    // frames per weight-gradient tile (K of one pipeline stage)
constexpr int kSubW = TW * 128;                           // 8 KB: 64 rows x 32 fp32
constexpr int kWgA = 4 * kSubW;                           // one A stage: A_hi (2 sub) | A_lo (2 sub) = 32 KB
constexpr int kWgAStages = 4, kWgBStages = 3;   // 3 B stages: the next tile's x lands and is split while this tile's x and h are in use
constexpr int kWgOffB = kWgAStages * kWgA;                // B stages: B_hi | B_lo, 32 KB each
constexpr int kWgOffBits = kWgOffB + kWgBStages * kWgA;   // 64 x uint2 keep-bits
constexpr int kWgOffBars = kWgOffBits + TW * 8;
constexpr int kWgNumBars = 3 * kWgAStages + 3 * kWgBStages + 1;
constexpr int kWgOffTmemPtr = kWgOffBars + kWgNumBars * 8;
constexpr int kTcWgradSmem = kWgOffTmemPtr + 16 + 1024;
constexpr int kWgPartFloats = 4 * 4096 + 4 * 64;          // per CTA: [4][64][64] weight partials | [4][64] bias sums

// MN-major operand made of 64-row sub-tiles: blocks of 32 channels are 8 KB apart
__device__ __forceinline__ uint32_t umma_desc_lo_mn64(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | ((8192u >> 4) << 16); }

__global__ void __launch_bounds__(kTcThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_gu, const __grid_constant__ CUtensorMap tm_gy,
                const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h,
                const __grid_constant__ CUtensorMap tm_q, TcWgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // aligned up by an offset added to the __shared__ array itself, so that the compiler keeps the address space
  // (pointer <- integer casts made every access below a generic LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint2* sBits = reinterpret_cast<uint2*>(smem + kWgOffBits);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgOffBars);
  uint64_t* bar_afull = bars;                               // [4] A tile landed
  uint64_t* bar_aready = bars + kWgAStages;                 // [4] A transformed (one arrival per transform warp)
  uint64_t* bar_aempty = bars + 2 * kWgAStages;             // [4] MMAs that read the A stage are complete
  uint64_t* bar_bfull = bars + 3 * kWgAStages;              // [2] B tile landed
  uint64_t* bar_bready = bar_bfull + kWgBStages;            // [2] B split done
  uint64_t* bar_bempty = bar_bready + kWgBStages;           // [2] MMAs that read the B stage are complete
  uint64_t* bar_done = bar_bempty + kWgBStages;             // all MMAs complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kWgOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool prof = a.dbg != nullptr && blockIdx.x == 0;
  long long w0 = 0, w1 = 0, w2 = 0, tstart = prof ? clock64() : 0;
#define WG_TIMED(acc, stmt) do { if (prof) { const long long t_ = clock64(); stmt; acc += clock64() - t_; } else { stmt; } } while (0)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_gu); tma_prefetch_desc(&tm_gy); tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_h); tma_prefetch_desc(&tm_q);
    for (int i = 0; i < kWgAStages; ++i) { mbar_init(bar_afull + i, 1); mbar_init(bar_aready + i, kEpiWarps); mbar_init(bar_aempty + i, 1); }
    for (int i = 0; i < kWgBStages; ++i) { mbar_init(bar_bfull + i, 1); mbar_init(bar_bready + i, kEpiWarps); mbar_init(bar_bempty + i, 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);
  const bool is_tail = (int)blockIdx.x >= a.nlayers * a.ctas_per_layer;
  const int layer = is_tail ? a.nlayers : (int)blockIdx.x / a.ctas_per_layer;
  const int rank = (int)blockIdx.x - layer * a.ctas_per_layer;
  const int nrank = is_tail ? a.tail_ctas : a.ctas_per_layer;
  const int dil = is_tail ? 0 : (a.dil_from_layer ? (1 << layer) : a.d);
  const uint32_t layer_id = a.layer_id + (uint32_t)layer;
  const int proj_left = a.nchunks - 4 * layer;
  const int tap_mask = a.proj ? ((1 << (proj_left < 4 ? proj_left : 4)) - 1) : (is_tail ? a.tail_tap_mask : a.tap_mask);
  const bool proj = a.proj != 0;
  const bool gy_xform = !is_tail && a.gy_transform != 0, tap3_full = is_tail || a.tap3_full_T != 0;
  const int c_gy = is_tail ? 0 : layer + a.cg_off, c_h = is_tail ? 0 : layer;

  // tap k's A tile starts at frame tf and holds something non-zero only if it overlaps [0, min(T, len))
  // (gu and go vanish at and beyond len)
  auto tap_tf = [&](int t0, int k) { return k == 3 ? t0 : t0 - (k - 1) * dil; };
  auto tap_present = [&](int t0, int k, int len) {
    if (!((tap_mask >> k) & 1)) return false;
    const int tf = tap_tf(t0, k);
    const int lim = (len < a.T && !(k == 3 && tap3_full) && !proj) ? len : a.T;
    return (tf + TW - 1 >= 0) && (tf < lim);
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t na = 0, nb = 0;
      for (int tile = rank; tile < a.num_tiles; tile += nrank) {
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TW;
        const int len = __ldg(a.lens + b);
        bool bx = false;
        for (int k = 0; k < 4; ++k) {
          if (!tap_present(t0, k, len)) continue;
          if (proj ? !bx : ((k < 3 && !bx) || k == 3)) {    // B event: x before the first gu tap, h before tap 3
            bx = true;
            const CUtensorMap* mb = (k == 3 && !proj) ? (is_tail ? &tm_q : &tm_h) : &tm_x;
            const int cb = proj ? 0 : (k == 3 ? c_h : layer);
            const uint32_t bs = nb % kWgBStages;
            WG_TIMED(w0, mbar_wait(bar_bempty + bs, ((nb / kWgBStages) & 1) ^ 1));
            mbar_arrive_expect_tx(bar_bfull + bs, 2 * kSubW);
            tma_load_4d(smem + kWgOffB + bs * kWgA, mb, bar_bfull + bs, 0, t0, b, cb);
            tma_load_4d(smem + kWgOffB + bs * kWgA + kSubW, mb, bar_bfull + bs, 32, t0, b, cb);
            ++nb;
          }
          const uint32_t st = na & 3;
          WG_TIMED(w1, mbar_wait(bar_aempty + st, ((na >> 2) & 1) ^ 1));
          const CUtensorMap* ma = (k == 3 && !proj) ? &tm_gy : &tm_gu;
          const int tf = proj ? t0 : tap_tf(t0, k);
          mbar_arrive_expect_tx(bar_afull + st, 2 * kSubW);
          const int ca = proj ? 0 : (k == 3 ? c_gy : layer);
          const int cc = proj ? 64 * (4 * layer + k) : 0;       // feature chunk (columns beyond dim read as zero)
          tma_load_4d(smem + st * kWgA, ma, bar_afull + st, cc, tf, b, ca);
          tma_load_4d(smem + st * kWgA + kSubW, ma, bar_afull + st, cc + 32, tf, b, ca);
          ++na;
        }
      }
      if (prof) { a.dbg[32] = w0; a.dbg[33] = w1; a.dbg[34] = na; a.dbg[35] = nb; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    const uint32_t usbase = __reduce_or_sync(0xffffffffu, sbase), utmem = __reduce_or_sync(0xffffffffu, tmem);
    constexpr uint32_t idesc = umma_idesc_tf32(TM, 128) | (1u << 15) | (1u << 16);     // A and B MN-major
    uint32_t na = 0, nb = 0, inited = 0, bd = 0;
    for (int tile = rank; tile < a.num_tiles; tile += nrank) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TW;
      const int len = __ldg(a.lens + b);
      bool bx = false;
      for (int k = 0; k < 4; ++k) {
        if (!tap_present(t0, k, len)) continue;
        if (proj ? !bx : ((k < 3 && !bx) || k == 3)) {
          bx = true;
          if (nb > 0) umma_commit(bar_bempty + ((nb - 1) % kWgBStages), 1);   // the MMAs that read the previous B stage are all issued
          WG_TIMED(w0, mbar_wait(bar_bready + (nb % kWgBStages), (nb / kWgBStages) & 1));
          bd = umma_desc_lo_mn64(usbase + kWgOffB + (nb % kWgBStages) * kWgA);
          ++nb;
        }
        const uint32_t st = na & 3;
        WG_TIMED(w1, mbar_wait(bar_aready + st, (na >> 2) & 1));
        tc_fence_after_sync();
        const uint32_t ad = umma_desc_lo_mn64(usbase + st * kWgA);
        const uint32_t dk = utmem + k * 128;
        const uint32_t first = (inited >> k) & 1u;
#pragma unroll
        for (int kk = 0; kk < TW / 8; ++kk)
          umma_tf32_ss(dk, ad + kk * 64, bd + kk * 64, idesc, (kk != 0) | first, 1, kDescHiMn32);
        inited |= 1u << k;
        umma_commit(bar_aempty + st, 1);
        ++na;
      }
    }
    umma_commit(bar_done, 1);
    if (prof && lane == 0) { a.dbg[36] = w0; a.dbg[37] = w1; a.dbg[38] = clock64() - tstart; }
    __syncwarp();
  } else {
    // ============ transform warps (then the final epilogue) ============
    const int q = warp & 3, s = (warp - 2) >> 2;
    const int etid = tid - 64;                   // 0..255
    const int j = etid & 15;                     // 16-byte chunk column: channels 4j..4j+3
    const int sub = j >> 3, cq = j & 7;
    float bsum[4][4] = {};
    uint32_t na = 0, nb = 0, used = 0;
    for (int tile = rank; tile < a.num_tiles; tile += nrank) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TW;
      const int len = __ldg(a.lens + b);
      bool bx = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!tap_present(t0, k, len)) continue;
        if (proj ? !bx : ((k < 3 && !bx) || k == 3)) {   // B event: split the x / h tile into hi (as is) and lo
          bx = true;
          const uint32_t bs = nb % kWgBStages;
          WG_TIMED(w0, mbar_wait(bar_bfull + bs, (nb / kWgBStages) & 1));
          uint8_t* bb = smem + kWgOffB + bs * kWgA;
#pragma unroll
          for (int i = 0; i < TW / 16; ++i) {
            const int r = (etid >> 4) + 16 * i;
            const uint32_t off = sub * kSubW + sw32_off(r, cq);
            const float4 w = *reinterpret_cast<const float4*>(bb + off);
            *reinterpret_cast<uint4*>(bb + 2 * kSubW + off) = make_uint4(lo_bits(w.x), lo_bits(w.y), lo_bits(w.z), lo_bits(w.w));
            if (proj) { bsum[0][0] += w.x; bsum[0][1] += w.y; bsum[0][2] += w.z; bsum[0][3] += w.w; }   // bias gradient = column sums of g0
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_bready + bs);
          ++nb;
        }
        const uint32_t st = na & 3;
        uint8_t* base = smem + st * kWgA;
        const bool gy = k == 3 && gy_xform;
        if (gy && a.train) {                     // keep-bits of the tile's frames, one Philox call each
          if (etid < TW) sBits[etid] = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), layer_id, a.frame0 + (uint32_t)(b * a.T + t0 + etid));
          named_bar_sync(5, 32 * kEpiWarps);
        }
        WG_TIMED(w1, mbar_wait(bar_afull + st, (na >> 2) & 1));
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < TW / 16; ++i) {
          const int r = (etid >> 4) + 16 * i;
          const uint32_t off = sub * kSubW + sw32_off(r, cq);
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
          *reinterpret_cast<uint4*>(base + 2 * kSubW + off) = make_uint4(lo_bits(v.x), lo_bits(v.y), lo_bits(v.z), lo_bits(v.w));
        }
        if (!proj) {
#pragma unroll
          for (int c = 0; c < 4; ++c) bsum[k][c] += cs[c];
        }
        fence_proxy_async_smem();                // generic writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_aready + st);
        if (gy && a.train) named_bar_sync(5, 32 * kEpiWarps);     // sBits may be rewritten for the next tile
        used |= 1u << k;
        ++na;
      }
    }
    // ---- final epilogue: the four 64x64 quadrants of D_k (A_hi/A_lo lanes x B_hi/B_lo columns) are
    //      added and staged through swizzled shared-memory rows so the global stores are whole rows ----
    if (prof && etid == 0) { a.dbg[39] = w0; a.dbg[40] = w1; a.dbg[41] = clock64() - tstart; }
    WG_TIMED(w2, mbar_wait(bar_done, 0));
    tc_fence_after_sync();
    uint8_t* sum = smem;                                     // [4][64 rows][256 B] over the (now idle) stages
    float* part = a.part + (size_t)blockIdx.x * kWgPartFloats;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 32);
    const int orow = (q & 1) * 32 + lane;                    // output channel of this thread's TMEM lane
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {                   // pass 0: A_lo lanes park their sums; pass 1: A_hi lanes add
      if ((pass == 0) == (q >= 2)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v[32];
          if ((used >> k) & 1u) {
            uint32_t u0[32], u1[32];
            tmem_ld32(trow + k * 128, u0);
            tmem_ld32(trow + k * 128 + 64, u1);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u0[i]) + __uint_as_float(u1[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4* p4 = reinterpret_cast<float4*>(sum + k * 16384 + stage_off(orow, s * 8 + c));
            float4 o = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            if (pass == 1) { const float4 l = *p4; o.x += l.x; o.y += l.y; o.z += l.z; o.w += l.w; }
            *p4 = o;
          }
        }
      }
      named_bar_sync(5, 32 * kEpiWarps);
    }
    for (int i = etid; i < 4 * 64 * 16; i += 32 * kEpiWarps) {     // 16 float4 per row, rows contiguous in `part`
      const int k = i >> 10, r = (i >> 4) & 63, c = i & 15;
      reinterpret_cast<float4*>(part + (k * 64 + r) * 64)[c] = *reinterpret_cast<const float4*>(sum + k * 16384 + stage_off(r, c));
    }
    named_bar_sync(5, 32 * kEpiWarps);
    // bias sums: 16 row-groups x 64 channels per tap -> fixed-order sum
    float* red = reinterpret_cast<float*>(smem);             // [4][16][64]
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(red + (k * 16 + (etid >> 4)) * 64 + 4 * j) = make_float4(bsum[k][0], bsum[k][1], bsum[k][2], bsum[k][3]);
    named_bar_sync(5, 32 * kEpiWarps);
    if (etid < 64) {
      for (int k = 0; k < 4; ++k) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) t += red[(k * 16 + g) * 64 + etid];
        part[4 * 4096 + k * 64 + etid] = t;
      }
    }
    tc_fence_before_sync();
    if (prof && etid == 0) { a.dbg[42] = w2; a.dbg[43] = clock64() - tstart; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
#undef WG_TIMED
}

// Projection mode of tc_wgrad_kernel: fixed-order sum of its per-CTA partials into the native-layout gradient of the stage-1
// input convolution.  Group g (four 64-feature chunks) owns CTAs g*P .. g*P+P-1; part[cta][k][c'][o] is the partial of
// dW[o][64*(4g+k) + c'], and the bias slot of tap 0 (any group; group 0 is read) is sum_t g0[t][o].
struct ProjWgradReduceArgs { const float* part; float* gw; float* gb; int dim, P, accumulate; };
__global__ void __launch_bounds__(256) proj_wgrad_reduce_kernel(ProjWgradReduceArgs a) {
  // 256 threads = 32 consecutive outputs (o fastest: contiguous in the partials) x 8 partial lanes; lane y sums partials
  // y, y+8, ..., then the 8 lane sums are added in fixed order -> deterministic
  __shared__ float red[8][33];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + tx;
  const int total = a.dim * 64;
  const float* p = nullptr;
  float* dst = nullptr;
  if (idx < total) {
    const int c = idx >> 6, o = idx & 63;
    const int chunk = c >> 6, g = chunk >> 2, k = chunk & 3;
    p = a.part + (size_t)g * a.P * kWgPartFloats + k * 4096 + (c & 63) * 64 + o;
    dst = a.gw + (size_t)o * a.dim + c;
  } else if (idx < total + 64) {
    p = a.part + 4 * 4096 + (idx - total);
    dst = a.gb + (idx - total);
  }
  float v = 0.f;
  if (p != nullptr)
    for (int r = ty; r < a.P; r += 8) v += p[(size_t)r * kWgPartFloats];
  red[ty][tx] = v;
  __syncthreads();
  if (ty == 0 && dst != nullptr) {
    float t = red[0][tx];
#pragma unroll
    for (int y = 1; y < 8; ++y) t += red[y][tx];
    *dst = a.accumulate ? *dst + t : t;
  }
}

// =============================================================================================
// Layer backward, pre-activation gradient:  gu = (W1^T (gy * mask * dropout)) * [h > 0]
// One 128-frame tile at a time: TMA brings the gy and h tiles, the epilogue warps turn gy into
// go = gy*mask*dropout in place (its trunc is the hi operand) and park go_lo in TMEM, 24 tcgen05.mma
// (3xTF32 against the transposed 1x1 image) produce gh in TMEM, and the same warps apply the ReLU
// mask from the h tile and write gu through swizzled staging as whole rows.
// =============================================================================================
struct TcBwdGuArgs {
  const int* lens; const float* wimg_b; float* gu;
  const int* flags_in; int* flags_out;    // per-tile flags of the kernel before / of this launch (see TcLayerFwdArgs::flags_in)
  int B, T, tiles_per_video, num_tiles;
  int train; uint32_t layer_id; uint64_t seed, offset;
  const unsigned long long* offset_dev;   // optional device-side step counter added to `offset` (CUDA-graph replay)
  uint32_t frame0;                        // global index of this launch's first frame (video-group launches keep the
                                          // whole-batch frame numbering of the Philox stream)
};
constexpr int kGuOffW = 0;                               // W1T_hi (2 sub) | W1T_lo (2 sub) = 32 KB
constexpr int kGuOffG = 4 * kSubB;                       // gy / go tile
constexpr int kGuOffH = kGuOffG + kSlot;                 // h tile
constexpr int kGuOffStage = kGuOffH + kSlot;             // gu staging
constexpr int kGuOffBars = kGuOffStage + kSlot;
constexpr int kGuOffTmemPtr = kGuOffBars + 8 * 8;
constexpr int kTcBwdGuSmem = kGuOffTmemPtr + 16 + 1024;

__global__ void __launch_bounds__(kTcThreads, 1)
tc_bwd_gu_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_h, TcBwdGuArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // aligned up by an offset added to the __shared__ array itself, so that the compiler keeps the address space
  // (pointer <- integer casts made every access below a generic LD.E / ST.E)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGuOffBars);
  uint64_t* bar_w = bars;            // weight image landed
  uint64_t* bar_full = bars + 1;     // gy + h tiles landed
  uint64_t* bar_a = bars + 2;        // go in place + go_lo parked (one arrival per epilogue warp)
  uint64_t* bar_g = bars + 3;        // gh accumulator complete
  uint64_t* bar_free = bars + 4;     // gy / h slots may be refilled (one arrival per epilogue warp)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kGuOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_h);
    mbar_init(bar_w, 1); mbar_init(bar_full, 1); mbar_init(bar_a, kEpiWarps); mbar_init(bar_g, 1);
    mbar_init(bar_free, kEpiWarps);
    fence_barrier_init();
    mbar_arrive_expect_tx(bar_w, 4 * kSubB);             // the 1x1 part of the backward image: sub-tiles 12..15
    for (int i = 0; i < 4; ++i) bulk_load(smem + kGuOffW + i * kSubB, a.wimg_b + (12 + i) * (kSubB / 4), kSubB, bar_w);
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (a.flags_in == nullptr) pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);
  constexpr uint32_t kColLo = 0, kColG = 64;
  constexpr uint32_t idesc = umma_idesc_tf32(TM, 64);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
        if (t0 >= __ldg(a.lens + b)) continue;
        if (a.flags_in != nullptr) {            // the tail's gradient tile is published (kernel-to-kernel dataflow)
          const long long tw0 = clock64();
          while (ld_flag(a.flags_in + tile) == 0) {
            __nanosleep(40);
            if (clock64() - tw0 > 8000000000LL) trap_report(3, tile, blockIdx.x);
          }
          fence_after_flags_seen();
        }
        mbar_wait(bar_free, (it & 1) ^ 1);
        mbar_arrive_expect_tx(bar_full, 2 * kSlot);
        tma_load_3d(smem + kGuOffG, &tm_g, bar_full, 0, t0, b);
        tma_load_3d(smem + kGuOffG + kSubA, &tm_g, bar_full, 32, t0, b);
        tma_load_3d(smem + kGuOffH, &tm_h, bar_full, 0, t0, b);
        tma_load_3d(smem + kGuOffH + kSubA, &tm_h, bar_full, 32, t0, b);
        ++it;
      }
    }
  } else if (warp == 1) {
    const uint32_t usbase = __reduce_or_sync(0xffffffffu, sbase), utmem = __reduce_or_sync(0xffffffffu, tmem);
    const uint32_t ag = umma_desc_lo(usbase + kGuOffG);
    const uint32_t wh = umma_desc_lo(usbase + kGuOffW), wl = umma_desc_lo(usbase + kGuOffW + 2 * kSubB);
    const uint32_t tG = utmem + kColG, tLo = utmem + kColLo;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      if (t0 >= __ldg(a.lens + b)) continue;
      if (it == 0) mbar_wait(bar_w, 0);
      mbar_wait(bar_a, it & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ad = ag + ((s * kSubA + ks * 32) >> 4);
          const uint32_t wo = (s * kSubB + ks * 32) >> 4;
          umma_tf32_ss(tG, ad, wh + wo, idesc, (s | ks) != 0, 1);
          umma_tf32_ss(tG, ad, wl + wo, idesc, 1, 1);
          umma_tf32_ts(tG, tLo + s * 32 + ks * 8, wh + wo, idesc, 1, 1);
        }
      umma_commit(bar_g, 1);
      ++it;
    }
    __syncwarp();
  } else {
    const int q = warp & 3, s = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int etid = tid - 64;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 32);
    uint8_t* stage = smem + kGuOffStage;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TM;
      const int len = __ldg(a.lens + b);
      const size_t vbase = (size_t)b * a.T * C;
      if (t0 >= len) {                          // gy is masked away entirely: gu = 0
        for (int i = etid; i < TM * 16; i += 32 * kEpiWarps) {
          const int t = t0 + (i >> 4);
          if (t < a.T) reinterpret_cast<float4*>(a.gu + vbase + (size_t)t * C)[i & 15] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (a.flags_out != nullptr) publish_tile(a.flags_out + tile, etid);
        continue;
      }
      const uint32_t p = it & 1;
      const int t = t0 + row;
      uint32_t keep = 0xffffffffu;
      if (a.train) {
        const uint2 bits = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), a.layer_id, a.frame0 + (uint32_t)(b * a.T + t));
        keep = s == 0 ? bits.x : bits.y;
      }
      const float on = (t < len) ? (a.train ? 2.f : 1.f) : 0.f;
      mbar_wait(bar_full, p);
      {
        uint8_t* sub = smem + kGuOffG + s * kSubA;
        uint32_t lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4* p4 = reinterpret_cast<float4*>(sub + sw128_off(row, c));
          float4 v = *p4;
          v.x = ((keep >> (4 * c)) & 1u) ? v.x * on : 0.f;
          v.y = ((keep >> (4 * c + 1)) & 1u) ? v.y * on : 0.f;
          v.z = ((keep >> (4 * c + 2)) & 1u) ? v.z * on : 0.f;
          v.w = ((keep >> (4 * c + 3)) & 1u) ? v.w * on : 0.f;
          *p4 = v;
          lo[4 * c] = lo_bits(v.x); lo[4 * c + 1] = lo_bits(v.y); lo[4 * c + 2] = lo_bits(v.z); lo[4 * c + 3] = lo_bits(v.w);
        }
        tmem_st32(trow + kColLo, lo);
        tmem_wait_st();
      }
      fence_proxy_async_smem();                 // go (generic-proxy writes) -> readable by the tensor core
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
      mbar_wait(bar_g, p);
      tc_fence_after_sync();
      {
        uint32_t v[32];
        tmem_ld32(trow + kColG, v);
        tmem_wait_ld();
        const uint8_t* hsub = smem + kGuOffH + s * kSubA;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 hv = *reinterpret_cast<const float4*>(hsub + sw128_off(row, c));
          *reinterpret_cast<float4*>(stage + stage_off(row, s * 8 + c)) =
              make_float4(hv.x > 0.f ? __uint_as_float(v[4 * c]) : 0.f, hv.y > 0.f ? __uint_as_float(v[4 * c + 1]) : 0.f,
                          hv.z > 0.f ? __uint_as_float(v[4 * c + 2]) : 0.f, hv.w > 0.f ? __uint_as_float(v[4 * c + 3]) : 0.f);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free);     // gy / h tiles consumed; the next tile's TMA may land
      copy_out_rows(stage, a.gu + vbase, t0, a.T, q, s, lane);
      named_bar_sync(1 + q, 64);                // pair done reading staging before the next tile rewrites it
      if (a.flags_out != nullptr) publish_tile(a.flags_out + tile, etid);
      ++it;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
  if (a.flags_in != nullptr && tid == 0) pdl_wait();      // never complete before the predecessor grid (see tc_layer_kernel)
}

// =============================================================================================
// Tail operand images (same 32768-float shape as a layer image so the layer kernels can consume them):
//   forward : centre-tap sub-tiles <- B[n = class j (zero-padded to 64)][K = channel c] = Wout[j][c];
//             1x1 part            <- B[n = out o][K = class j (padded)]                 = Wn[o][j]   (next stage's conv_1x1)
//   backward: centre-tap sub-tiles <- B[n = class j][K = out o] = Wn[o][j]      (gq = Wn^T gin);
//             1x1 part            <- B[n = channel c][K = class j] = Wout[j][c] (ga = Wout^T gz)
// Only the sub-tiles the tail modes load are written (Wd K-blocks 2,3 and the 1x1 part).
// =============================================================================================
// All tensor-core operand images in ONE launch, destination-major (coalesced stores, gathered reads of the small native
// weight arrays): per thread one (row n, K index) position of one 64x32 sub-tile, written as its hi and its lo word.
// Layer images: forward B[n = out][K = tap*64 + in] = Wd[out][in][tap] and B[n = out][K = in] = W1[out][in]; the backward
// images hold the transposes (B[n = in][K = tap*64 + out], B[n = in][K = out]).  hi = rna_tf32(W), lo = rna_tf32(W - hi).
//   item space: [S*L layers x {fwd, bwd} x 8 sub-tiles (6 Wd K-blocks, 2 W1)] [S tails x {fwd, bwd} x 4 sub-tiles
//   (Wd K-blocks 2,3 = centre tap; 2 W1)] [projection K-blocks], 2048 positions each.
// =============================================================================================
__device__ __forceinline__ void subtile_pos(int r, int& n, int& k32) {      // inverse of wimg_index inside one sub-tile
  n = ((r >> 8) << 3) | ((r >> 5) & 7);
  k32 = ((((r >> 2) & 7) ^ (n & 7)) << 2) | (r & 3);
}
__device__ __forceinline__ void put_hi_lo(float* hi_dst, float* lo_dst, float w) {
  const uint32_t hi = tf32_rna(w);
  *hi_dst = __uint_as_float(hi);
  *lo_dst = __uint_as_float(tf32_rna(w - __uint_as_float(hi)));
}

__global__ void __launch_bounds__(256) tc_pack_all_kernel(Layout lay, const float* __restrict__ params, float* __restrict__ packed) {
  const int S = lay.S, L = lay.L, K = lay.K;
  const long long n_layer = (long long)S * L * 2 * 8 * 2048, n_tail = (long long)S * 2 * 4 * 2048;
  const long long n_proj = (long long)lay.proj_kblocks() * 2048;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_layer + n_tail + n_proj;
       i += (long long)gridDim.x * blockDim.x) {
    int n, k32;
    if (i < n_layer) {
      const int r = (int)(i & 2047), st = (int)((i >> 11) & 7), bwd = (int)((i >> 14) & 1), sl = (int)(i >> 15);
      const int s = sl / L, l = sl - s * L;
      subtile_pos(r, n, k32);
      float* img = packed + (bwd ? lay.p_tcb(s, l) : lay.p_tc(s, l));
      if (st < 6) {                                   // dilated conv, K-block st: K index = tap*64 + channel
        const int kk = st * 32 + k32, tap = kk >> 6, c = kk & 63;
        // forward: B[n = out][K = tap*64 + in] ; backward: B[n = in][K = tap*64 + out]   (both = Wd[out][in][tap])
        const int o = bwd ? c : n, ci = bwd ? n : c;
        const float w = params[lay.wd(s, l) + ((size_t)o * 64 + ci) * 3 + tap];
        put_hi_lo(img + st * 4096 + r, img + st * 4096 + 2048 + r, w);
      } else {                                        // 1x1: forward B[n = out][K = in], backward B[n = in][K = out]
        const int kk = (st - 6) * 32 + k32;
        const int o = bwd ? kk : n, ci = bwd ? n : kk;
        const float w = params[lay.w1(s, l) + (size_t)o * 64 + ci];
        put_hi_lo(img + kOffW1Hi / 4 + (st - 6) * 2048 + r, img + kOffW1Lo / 4 + (st - 6) * 2048 + r, w);
      }
    } else if (i < n_layer + n_tail) {
      const long long j = i - n_layer;
      const int r = (int)(j & 2047), st = (int)((j >> 11) & 3), bwd = (int)((j >> 13) & 1), s = (int)(j >> 14);
      subtile_pos(r, n, k32);
      float* img = packed + (bwd ? lay.p_ttb(s) : lay.p_tt(s));
      const bool has_next = s + 1 < S;
      const int kk = (st & 1) * 32 + k32;             // K index inside the 64-wide centre tap / 1x1
      float w = 0.f;
      if (st < 2) {      // centre-tap K-blocks 2,3.  forward: B[n = class][K = channel] = Wout ; backward: B[n = class][K = out] = Wn
        if (n < K) w = bwd ? (has_next ? params[lay.win_w(s + 1) + (size_t)kk * K + n] : 0.f) : params[lay.wout(s) + (size_t)n * 64 + kk];
        put_hi_lo(img + (2 + st) * 4096 + r, img + (2 + st) * 4096 + 2048 + r, w);
      } else {           // 1x1 part.  forward: B[n = out][K = class] = Wn ; backward: B[n = channel][K = class] = Wout
        if (kk < K) w = bwd ? params[lay.wout(s) + (size_t)kk * 64 + n] : (has_next ? params[lay.win_w(s + 1) + (size_t)n * K + kk] : 0.f);
        put_hi_lo(img + kOffW1Hi / 4 + (st - 2) * 2048 + r, img + kOffW1Lo / 4 + (st - 2) * 2048 + r, w);
      }
    } else {
      const long long j = i - n_layer - n_tail;
      const int r = (int)(j & 2047), kb = (int)(j >> 11);
      subtile_pos(r, n, k32);
      const int c = kb * 32 + k32;
      const float w = c < lay.dim ? params[lay.win_w(0) + (size_t)n * lay.dim + c] : 0.f;
      float* img = packed + lay.p_tp() + (size_t)kb * 4096;
      put_hi_lo(img + r, img + 2048 + r, w);
    }
  }
}

// =============================================================================================
// Stage-1 input projection  y[n] = Win x[n] + bin  (SingleStageModel.conv_1x1, networks.py:330; NOT masked, fact 0.5)
// on the tensor cores.  x is the caller's (B*T, D) feature matrix, read in place through a 2-D tensor map in
// K-blocks of 32 features (the last block's out-of-range columns are zero-filled by TMA); the weight image holds, per
// K-block, [W_hi | W_lo] so that x_hi*[W_hi|W_lo] is one m128 n128 k8 MMA; x_lo is parked in TMEM (ring of 4 blocks) for
// the third product.  Tiles are 128 consecutive rows of the flat frame axis (the op is pointwise); a tile that lies
// wholly in one video's padding (x = 0) gets the bias without touching the tensor core.
// =============================================================================================
struct TcProjArgs {
  const float* wimg; const float* bias; const int* lens; float* y;
  long long n_rows; int T, kblocks, num_tiles;
};
constexpr int kPjStages = 4;
constexpr int kPjStage = kSubA + 2 * kSubB;                       // x block 16 KB | [W_hi | W_lo] 16 KB
constexpr int kPjOffOut = kPjStages * kPjStage;                   // 128 KB: output staging (32 KB)
constexpr int kPjOffBias = kPjOffOut + kSlot;
constexpr int kPjOffBars = kPjOffBias + 256;
constexpr int kPjNumBars = 3 * kPjStages + 2;
constexpr int kPjOffTmemPtr = kPjOffBars + kPjNumBars * 8;
constexpr int kTcProjSmem = kPjOffTmemPtr + 16 + 1024;

__global__ void __launch_bounds__(kTcThreads, 1)
tc_proj_kernel(const __grid_constant__ CUtensorMap tm_x, TcProjArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sBias = reinterpret_cast<float*>(smem + kPjOffBias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPjOffBars);
  uint64_t* bar_full = bars;                         // [4] x block + weight block landed
  uint64_t* bar_lo = bars + kPjStages;               // [4] x_lo of the block parked (4 warps)
  uint64_t* bar_empty = bars + 2 * kPjStages;        // [4] the block's MMAs are complete
  uint64_t* bar_done = bars + 3 * kPjStages;         // the tile's accumulator is complete
  uint64_t* bar_accfree = bar_done + 1;              // the epilogue has read the accumulator (8 warps)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kPjOffTmemPtr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < kPjStages; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_lo + i, 4); mbar_init(bar_empty + i, 1); }
    mbar_init(bar_done, 1); mbar_init(bar_accfree, kEpiWarps);
    fence_barrier_init();
  }
  if (tid >= 64 && tid < 128) sBias[tid - 64] = __ldg(a.bias + tid - 64);
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  pdl_wait();                                   // the operand image comes from the packing kernels launched just before
  pdl_launch_dependents();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t sbase = smem_u32(smem);
  constexpr uint32_t kColAcc = 0, kColLo = 128;

  // a tile that lies wholly inside one video's zero padding needs no GEMM
  auto tile_is_padding = [&](int tile) {
    if (a.lens == nullptr || a.T <= 0) return false;
    const long long r0 = (long long)tile * TM, r1 = r0 + TM - 1 < a.n_rows ? r0 + TM - 1 : a.n_rows - 1;
    const int b0 = (int)(r0 / a.T), b1 = (int)(r1 / a.T);
    return b0 == b1 && (int)(r0 - (long long)b0 * a.T) >= __ldg(a.lens + b0);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        if (tile_is_padding(tile)) continue;
        for (int kb = 0; kb < a.kblocks; ++kb, ++n) {
          const uint32_t st = n % kPjStages;
          mbar_wait(bar_empty + st, ((n / kPjStages) & 1) ^ 1);
          uint8_t* dst = smem + st * kPjStage;
          mbar_arrive_expect_tx(bar_full + st, kPjStage);
          tma_load_2d(dst, &tm_x, bar_full + st, kb * 32, tile * TM);
          bulk_load(dst + kSubA, a.wimg + (size_t)kb * 4096, 2 * kSubB, bar_full + st);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = lane == 0 ? 1u : 0u;
    const uint32_t usbase = __reduce_or_sync(0xffffffffu, sbase), utmem = __reduce_or_sync(0xffffffffu, tmem);
    constexpr uint32_t idesc = umma_idesc_tf32(TM, 64), idesc2 = umma_idesc_tf32(TM, 128);
    uint32_t n = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      if (tile_is_padding(tile)) continue;
      if (it > 0) { mbar_wait(bar_accfree, (it - 1) & 1); tc_fence_after_sync(); }
      for (int kb = 0; kb < a.kblocks; ++kb, ++n) {
        const uint32_t st = n % kPjStages, ph = (n / kPjStages) & 1;
        const uint32_t ad = umma_desc_lo(usbase + st * kPjStage), wd = umma_desc_lo(usbase + st * kPjStage + kSubA);
        mbar_wait(bar_full + st, ph);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_tf32_ss(utmem + kColAcc, ad + ks * 2, wd + ks * 2, idesc2, (kb | ks) != 0, leader);
        mbar_wait(bar_lo + st, ph);
        tc_fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_tf32_ts(utmem + kColAcc, utmem + kColLo + st * 32 + ks * 8, wd + ks * 2, idesc, 1, leader);
        umma_commit(bar_empty + st, leader);
      }
      umma_commit(bar_done, leader);
      ++it;
    }
    __syncwarp();
  } else {
    const int q = warp & 3, s = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int etid = tid - 64;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* stage = smem + kPjOffOut;
    uint32_t n = 0, it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const long long r0 = (long long)tile * TM;
      if (tile_is_padding(tile)) {
        for (int i = etid; i < TM * 16; i += 32 * kEpiWarps) {
          const long long r = r0 + (i >> 4);
          if (r < a.n_rows) reinterpret_cast<float4*>(a.y + (size_t)r * C)[i & 15] = *reinterpret_cast<const float4*>(sBias + 4 * (i & 15));
        }
        continue;
      }
      for (int kb = 0; kb < a.kblocks; ++kb, ++n) {
        if ((kb & 1) != s) continue;                     // the two warp sets take alternate K-blocks
        const uint32_t st = n % kPjStages, ph = (n / kPjStages) & 1;
        mbar_wait(bar_full + st, ph);
        const uint8_t* sub = smem + st * kPjStage;
        uint32_t lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(sub + sw128_off(row, c));
          lo[4 * c] = lo_bits(v.x); lo[4 * c + 1] = lo_bits(v.y); lo[4 * c + 2] = lo_bits(v.z); lo[4 * c + 3] = lo_bits(v.w);
        }
        tmem_st32(tlane + kColLo + st * 32, lo);
        tmem_wait_st();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_lo + st);
      }
      // ---- epilogue: acc = [x*W_hi (+ x_lo*W_hi) | x_hi*W_lo] -> + bias -> y ----
      mbar_wait(bar_done, it & 1);
      tc_fence_after_sync();
      {
        uint32_t v[32], w[32];
        tmem_ld32(tlane + kColAcc + s * 32, v);
        tmem_ld32(tlane + kColAcc + 64 + s * 32, w);
        tmem_wait_ld();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_accfree);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(stage + stage_off(row, s * 8 + c)) =
              make_float4(__uint_as_float(v[4 * c]) + __uint_as_float(w[4 * c]) + sBias[s * 32 + 4 * c],
                          __uint_as_float(v[4 * c + 1]) + __uint_as_float(w[4 * c + 1]) + sBias[s * 32 + 4 * c + 1],
                          __uint_as_float(v[4 * c + 2]) + __uint_as_float(w[4 * c + 2]) + sBias[s * 32 + 4 * c + 2],
                          __uint_as_float(v[4 * c + 3]) + __uint_as_float(w[4 * c + 3]) + sBias[s * 32 + 4 * c + 3]);
      }
      // rows of the flat frame axis: copy_out_rows with "video" = the whole matrix
      named_bar_sync(1 + q, 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = 32 * q + 16 * s + 2 * i + (lane >> 4);
        const float4 v = *reinterpret_cast<const float4*>(stage + stage_off(r, lane & 15));
        if (r0 + r < a.n_rows) reinterpret_cast<float4*>(a.y + (size_t)(r0 + r) * C)[lane & 15] = v;
      }
      named_bar_sync(1 + q, 64);                         // the pair is done with its staging rows
      ++it;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// gr[s][n][j] = [winner[n][j] == s] * gout[n][j] * gscale : the max over stages routes each (frame, class) gradient to
// the winning stage (torch.max backward, networks.py:319).  One pass writes every stage's plane.
__global__ void __launch_bounds__(256) route_grad_kernel(const float* __restrict__ gout, const float* __restrict__ gscale,
                                                         const uint8_t* __restrict__ winner, int S, int K, int64_t frames,
                                                         float* __restrict__ gr0, int64_t stage_stride) {
  // one thread per (frame, 4-class chunk) of the zero-padded (frames, 64) planes
  const float sc = gscale ? __ldg(gscale) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < frames * 16; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = i >> 4;
    const int c0 = (int)(i & 15) * 4;
    if (winner == nullptr) {                     // per-stage gradients (S, frames, K): just pad the rows to 64 classes
      for (int s = 0; s < S; ++s) {
        float g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) g[j] = c0 + j < K ? gout[((size_t)s * frames + f) * K + c0 + j] * sc : 0.f;
        *reinterpret_cast<float4*>(gr0 + (size_t)s * stage_stride + i * 4) = make_float4(g[0], g[1], g[2], g[3]);
      }
      continue;
    }
    float g[4]; int w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = c0 + j < K;
      g[j] = in ? gout[f * K + c0 + j] * sc : 0.f;
      w[j] = in ? winner[f * K + c0 + j] : -1;
    }
    for (int s = 0; s < S; ++s)
      *reinterpret_cast<float4*>(gr0 + (size_t)s * stage_stride + i * 4) =
          make_float4(w[0] == s ? g[0] : 0.f, w[1] == s ? g[1] : 0.f, w[2] == s ? g[2] : 0.f, w[3] == s ? g[3] : 0.f);
  }
}

// out[n][j] = max over stages of the per-stage logits, winner = first stage attaining it
// (torch.cat / permute / torch.max over dim 0, networks.py:312-319, first index on ties).
__global__ void __launch_bounds__(256) stage_max_kernel(const float* __restrict__ logits0, int64_t stage_stride, int S,
                                                        int64_t n, float* __restrict__ out, uint8_t* __restrict__ winner) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float best = logits0[i];
    int w = 0;
    for (int s = 1; s < S; ++s) {
      const float v = logits0[(size_t)s * stage_stride + i];
      if (v > best) { best = v; w = s; }
    }
    out[i] = best;
    winner[i] = (uint8_t)w;
  }
}

// Fused loss head of the reference training step: out = max over stages (networks.py:319), CrossEntropyLoss(ignore_index
// = -1) on it (train.py:267,326) and the backward of both -- dL/dz_s = [winner == s] * (softmax(out) - onehot) / n_valid --
// written straight into the zero-padded (frames, 64) routed-gradient planes the tail backward reads.  One warp per
// frame; same arithmetic and the same block partition as stage_max_kernel + ce_loss_kernel + route_grad_kernel, so the
// loss and every gradient are bit-identical to the unfused path.
__global__ void __launch_bounds__(256) loss_head_kernel(const float* __restrict__ logits0, int64_t stage_stride, int S, int K,
                                                        int64_t n_rows, const int64_t* __restrict__ y, float inv_nvalid,
                                                        float* __restrict__ gr0, int64_t gr_stride, float* __restrict__ out,
                                                        uint8_t* __restrict__ winner, float* __restrict__ scratch) {
  __shared__ float s_sum[8], s_cnt[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float my_sum = 0.f, my_cnt = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n_rows; row += (int64_t)gridDim.x * 8) {
    const bool c0 = lane < K, c1 = lane + 32 < K;
    float v0 = -INFINITY, v1 = -INFINITY;
    int w0 = 0, w1 = 0;
    for (int s = 0; s < S; ++s) {                       // first stage wins ties, like torch.max
      const float* zr = logits0 + (size_t)s * stage_stride + row * K;
      const float a0 = c0 ? zr[lane] : -INFINITY, a1 = c1 ? zr[lane + 32] : -INFINITY;
      if (s == 0 || a0 > v0) { v0 = a0; w0 = s; }
      if (s == 0 || a1 > v1) { v1 = a1; w1 = s; }
    }
    if (out != nullptr) {
      if (c0) { out[row * K + lane] = v0; winner[row * K + lane] = (uint8_t)w0; }
      if (c1) { out[row * K + lane + 32] = v1; winner[row * K + lane + 32] = (uint8_t)w1; }
    }
    const int64_t lab = y[row];
    float g0 = 0.f, g1 = 0.f;
    if (lab >= 0 && lab < K) {
      float mx = fmaxf(v0, v1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e0 = c0 ? expf(v0 - mx) : 0.f, e1 = c1 ? expf(v1 - mx) : 0.f;
      float sum = e0 + e1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      g0 = c0 ? (e0 * inv - (lane == lab ? 1.f : 0.f)) * inv_nvalid : 0.f;
      g1 = c1 ? (e1 * inv - (lane + 32 == lab ? 1.f : 0.f)) * inv_nvalid : 0.f;
      const float zl = __shfl_sync(0xffffffffu, lab < 32 ? v0 : v1, (int)(lab & 31));
      if (lane == 0) { my_sum += (mx + logf(sum)) - zl; my_cnt += 1.f; }
    } else if (lab != -1 && lane == 0) {
      my_sum += __int_as_float(0x7fc00000);               // label outside [0, K) and not ignore_index: the loss turns NaN
    }
    for (int s = 0; s < S; ++s) {
      float* gp = gr0 + (size_t)s * gr_stride + row * 64;
      gp[lane] = (c0 && w0 == s) ? g0 : 0.f;
      gp[lane + 32] = (c1 && w1 == s) ? g1 : 0.f;
    }
  }
  if (lane == 0) { s_sum[warp] = my_sum; s_cnt[warp] = my_cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += s_sum[i]; c += s_cnt[i]; }
    scratch[2 * blockIdx.x] = a; scratch[2 * blockIdx.x + 1] = c;
  }
}

// the two instantiations
template __global__ void tc_layer_kernel<0>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);
template __global__ void tc_layer_kernel<1>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);
template __global__ void tc_layer_kernel<2>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);
template __global__ void tc_layer_kernel<3>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);
template __global__ void tc_layer_kernel<4>(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, TcLayerFwdArgs);

}  // namespace tc
}  // namespace mstcn
