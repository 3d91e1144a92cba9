// Backward kernels of the exact-fp32 path (what autograd replays for loss.backward(),
// train.py:328).  Weight gradients are reduced in registers per CTA across its tiles, written
// as one partial per CTA, then summed across the grid by reduce_partials_kernel in a fixed
// order (deterministic).
#pragma once
#include "common.cuh"

namespace mstcn {

// --------------------------------------------------------------------------------------------
// Layer backward, pass A:  g = gy*mask; go = drop(g); gu = (W1^T go) * [h > 0]  (-> global gu)
//   partials: dW1 (64 out, 64 in) = sum go h^T ; db1 = sum go ; dbd = sum gu
// --------------------------------------------------------------------------------------------
struct LayerBwdAArgs {
  const float* gy; const float* h; float* gu; const int* lens; const float* w1;   // w1 native (out,in)
  float* part;   // per CTA: [4096 dW1 | 64 db1 | 64 dbd]
  int B, T, tiles_per_video, num_tiles;
  int train; uint32_t layer_id; uint64_t seed, offset;
  const unsigned long long* offset_dev;   // optional device-side step counter added to `offset` (CUDA-graph replay)
  uint32_t frame0;                        // global index of this launch's first frame (video-group launches keep the
                                          // whole-batch frame numbering of the Philox stream)
  int gu_only;   // 1: weight / bias gradients come from the tensor-core wgrad kernel
};
constexpr int kBwdAPart = 4096 + 128;
constexpr int kLayerBwdASmem = (3 * TILE + 8 * C) * 4 + 64 * 8;

__global__ void __launch_bounds__(NT, 2) layer_bwd_a_kernel(LayerBwdAArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sW1 = smem;               // (64 out, 64 in)
  float* sG = sW1 + TILE;          // go tile
  float* sH = sG + TILE;           // h tile
  float* sRed = sH + TILE;         // 8*64
  uint2* sBits = reinterpret_cast<uint2*>(sRed + 8 * C);
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  load_weights(sW1, a.w1, TILE / 4, tid);
  float accw[8][4] = {};
  float sb1[4] = {0.f, 0.f, 0.f, 0.f}, sbd[4] = {0.f, 0.f, 0.f, 0.f};

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TF;
    const int len = __ldg(a.lens + b);
    const size_t vbase = (size_t)b * a.T * C;
    if (t0 >= len) { zero_rows(a.gu + vbase, t0, a.T, tid); continue; }
    __syncthreads();
    load_tile(sG, a.gy + vbase, t0, a.T, tid);
    load_tile(sH, a.h + vbase, t0, a.T, tid);
    if (a.train && tid < TF)
      sBits[tid] = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), a.layer_id, a.frame0 + (uint32_t)(b * a.T + t0 + tid));
    __syncthreads();
    // go = gy * mask * dropout, in place (each thread rewrites only the chunks it owns below)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = fg + 8 * j, t = t0 + r;
      float4 g = ld4s(sG + swz(r, og));
      const float m = (t < len) ? 1.f : 0.f;
      float4 s = make_float4(m, m, m, m);
      if (a.train) { float4 d = dropout_scale4(sBits[r], og); s.x *= d.x; s.y *= d.y; s.z *= d.z; s.w *= d.w; }
      g.x *= s.x; g.y *= s.y; g.z *= s.z; g.w *= s.w;
      st4s(sG + swz(r, og), g);
      sb1[0] += g.x; sb1[1] += g.y; sb1[2] += g.z; sb1[3] += g.w;
    }
    __syncthreads();
    float acc[8][4] = {};
    fgemm(sG, sW1, acc, fg, og);       // gh[f][c] = sum_o go[f][o] W1[o][c]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = fg + 8 * j, t = t0 + r;
      const float4 hv = ld4s(sH + swz(r, og));
      float4 gu = make_float4(hv.x > 0.f ? acc[j][0] : 0.f, hv.y > 0.f ? acc[j][1] : 0.f,
                              hv.z > 0.f ? acc[j][2] : 0.f, hv.w > 0.f ? acc[j][3] : 0.f);
      sbd[0] += gu.x; sbd[1] += gu.y; sbd[2] += gu.z; sbd[3] += gu.w;
      if (t < a.T) reinterpret_cast<float4*>(a.gu + vbase + (size_t)t * C)[og] = gu;
    }
    if (!a.gu_only) wgemm(sG, sH, accw, fg, og);       // dW1[o][c] += sum_f go[f][o] h[f][c]
  }
  if (a.gu_only) return;
  float* part = a.part + (size_t)blockIdx.x * kBwdAPart;
  store_wacc(part, accw, fg, og);
  store_colsum(part + 4096, sb1, sRed, fg, og, tid);
  store_colsum(part + 4096 + 64, sbd, sRed, fg, og, tid);
}

// --------------------------------------------------------------------------------------------
// Layer backward, pass B:  gx[t] = gy[t]*mask + sum_k Wd[:,:,k]^T gu[t-(k-1)d]
//   partials: dWd[k] (64 out, 64 in) = sum_t gu[t-(k-1)d] x[t]^T   (x OOB = 0 = conv padding)
// --------------------------------------------------------------------------------------------
struct LayerBwdBArgs {
  const float* gy; const float* gu; const float* x; float* gx; const int* lens; const float* wd_b;  // (3, out, in)
  float* part;   // per CTA: [3][64 out][64 in]
  int B, T, d, tiles_per_video, num_tiles;
  int wgrad_only;   // 1: gx is produced by the tensor-core kernel; only the dWd partials are computed here
};
constexpr int kBwdBPart = 3 * 4096;
constexpr int kLayerBwdBSmem = (3 * TILE + 3 * TILE + TILE) * 4;

__global__ void __launch_bounds__(NT, 2) layer_bwd_b_kernel(LayerBwdBArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                 // (3, 64 out, 64 in)
  float* sU = sW + 3 * TILE;        // gu taps: tap k holds gu[t - (k-1)d]
  float* sX = sU + 3 * TILE;        // x tile (centre)
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  if (!a.wgrad_only) load_weights(sW, a.wd_b, 3 * TILE / 4, tid);
  float accw[3][8][4] = {};

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TF;
    const int len = __ldg(a.lens + b);
    const size_t vbase = (size_t)b * a.T * C;
    // gu is zero at and beyond len, so a tile starting at or after len + d sees only zeros
    if ((long long)t0 - a.d >= len) { if (!a.wgrad_only) zero_rows(a.gx + vbase, t0, a.T, tid); continue; }
    const bool tap0 = (t0 + a.d) < a.T;                 // reads gu[t + d]
    const bool tap2 = (t0 + TF - 1 - a.d) >= 0;         // reads gu[t - d]
    __syncthreads();
    if (tap0) load_tile(sU, a.gu + vbase, t0 + a.d, a.T, tid);
    load_tile(sU + TILE, a.gu + vbase, t0, a.T, tid);
    if (tap2) load_tile(sU + 2 * TILE, a.gu + vbase, t0 - a.d, a.T, tid);
    load_tile(sX, a.x + vbase, t0, a.T, tid);
    __syncthreads();
    if (!a.wgrad_only) {
      float acc[8][4] = {};
      if (tap0) fgemm(sU, sW, acc, fg, og);
      fgemm(sU + TILE, sW + TILE, acc, fg, og);
      if (tap2) fgemm(sU + 2 * TILE, sW + 2 * TILE, acc, fg, og);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = t0 + fg + 8 * j;
        if (t >= a.T) continue;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < len) g = __ldg(reinterpret_cast<const float4*>(a.gy + vbase + (size_t)t * C) + og);
        reinterpret_cast<float4*>(a.gx + vbase + (size_t)t * C)[og] =
            make_float4(acc[j][0] + g.x, acc[j][1] + g.y, acc[j][2] + g.z, acc[j][3] + g.w);
      }
    }
    if (tap0) wgemm(sU, sX, accw[0], fg, og);
    wgemm(sU + TILE, sX, accw[1], fg, og);
    if (tap2) wgemm(sU + 2 * TILE, sX, accw[2], fg, og);
  }
  float* part = a.part + (size_t)blockIdx.x * kBwdBPart;
#pragma unroll
  for (int k = 0; k < 3; ++k) store_wacc(part + k * 4096, accw[k], fg, og);
}

// --------------------------------------------------------------------------------------------
// Stage tail backward.
//   gq = Wn^T gin (next stage's input grad), gp = gq*mask, p = softmax(z):
//   gz = ( p*(gp - <gp,p>) + [winner==stage]*gout*gscale ) * mask
//   ga = Wout^T gz ; partials: dWout (64pad,64) = sum gz a^T, dbout = sum gz,
//                              dWn (64,64pad) = sum gin q^T (q = p*mask), dbn = sum gin (ALL frames)
// --------------------------------------------------------------------------------------------
struct TailBwdArgs {
  const float* a; const float* logits; const float* gout; const float* gscale; const uint8_t* winner;
  const float* gin; const int* lens; const float* wout_b; const float* wn_b;
  float* ga; float* part;   // per CTA: [4096 dWout | 64 dbout | 4096 dWn | 64 dbn]
  int B, T, K, stage, tiles_per_video, num_tiles;
};
constexpr int kTailBwdPart = 2 * (4096 + 64);
constexpr int kTailBwdSmem = (6 * TILE + 8 * C) * 4;

__global__ void __launch_bounds__(NT, 2) tail_bwd_kernel(TailBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sWnb = smem;               // (64 out, 64 class-padded)
  float* sWob = sWnb + TILE;        // (64 class-padded, 64 in)
  float* sGin = sWob + TILE;
  float* sA = sGin + TILE;
  float* sGz = sA + TILE;
  float* sQ = sGz + TILE;
  float* sRed = sQ + TILE;
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  const bool has_next = a.gin != nullptr;
  const int K = a.K;
  if (has_next) load_weights(sWnb, a.wn_b, TILE / 4, tid);
  load_weights(sWob, a.wout_b, TILE / 4, tid);
  const float gscale = a.gscale ? __ldg(a.gscale) : 1.f;
  float accwo[8][4] = {}, accwn[8][4] = {};
  float sbo[4] = {0.f, 0.f, 0.f, 0.f}, sbn[4] = {0.f, 0.f, 0.f, 0.f};

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TF;
    const int len = __ldg(a.lens + b);
    const size_t fbase = (size_t)b * a.T;
    if (t0 >= len) {
      // padding tile: gz = 0, so ga = 0 and no weight gradient -- except the next stage's (unmasked)
      // projection bias, which sums gin over ALL frames
      zero_rows(a.ga + fbase * C, t0, a.T, tid);
      if (has_next) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = t0 + fg + 8 * j;
          if (t < a.T) {
            const float4 gi = __ldg(reinterpret_cast<const float4*>(a.gin + (fbase + t) * C) + og);
            sbn[0] += gi.x; sbn[1] += gi.y; sbn[2] += gi.z; sbn[3] += gi.w;
          }
        }
      }
      continue;
    }
    __syncthreads();
    if (has_next) load_tile(sGin, a.gin + fbase * C, t0, a.T, tid);
    load_tile(sA, a.a + fbase * C, t0, a.T, tid);
    __syncthreads();
    float acc[8][4] = {};
    if (has_next) fgemm(sGin, sWnb, acc, fg, og);     // gq[f][j] = sum_o gin[f][o] Wn[o][j]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = fg + 8 * j, t = t0 + r;
      const bool inb = t < a.T;
      const float m = (t < len) ? 1.f : 0.f;
      const size_t row = (fbase + t) * (size_t)K;
      float gz[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
      if (has_next) {
        float z[4], e[4];
        float zmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * og + i;
          z[i] = (inb && c < K) ? __ldg(a.logits + row + c) : 0.f;
          if (c < K) zmax = fmaxf(zmax, z[i]);
        }
        zmax = row_max16(zmax);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { e[i] = (4 * og + i < K) ? expf(z[i] - zmax) : 0.f; sum += e[i]; }
        sum = row_sum16(sum);
        const float inv = 1.f / sum;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { e[i] *= inv; dot += acc[j][i] * m * e[i]; }
        dot = row_sum16(dot);
#pragma unroll
        for (int i = 0; i < 4; ++i) { gz[i] = e[i] * (acc[j][i] * m - dot); q[i] = e[i] * m; }
      }
      if (inb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * og + i;
          if (c < K && (a.winner == nullptr || __ldg(a.winner + row + c) == (uint8_t)a.stage)) gz[i] += __ldg(a.gout + row + c) * gscale;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) { gz[i] *= m; sbo[i] += gz[i]; }
      st4s(sGz + swz(r, og), make_float4(gz[0], gz[1], gz[2], gz[3]));
      if (has_next) {
        st4s(sQ + swz(r, og), make_float4(q[0], q[1], q[2], q[3]));
        const float4 gi = ld4s(sGin + swz(r, og));
        sbn[0] += gi.x; sbn[1] += gi.y; sbn[2] += gi.z; sbn[3] += gi.w;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
    fgemm(sGz, sWob, acc, fg, og);                   // ga[f][c] = sum_j gz[f][j] Wout[j][c]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = t0 + fg + 8 * j;
      if (t < a.T)
        reinterpret_cast<float4*>(a.ga + (fbase + t) * C)[og] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
    wgemm(sGz, sA, accwo, fg, og);                   // dWout[j][c] += sum_f gz[f][j] a[f][c]
    if (has_next) wgemm(sGin, sQ, accwn, fg, og);    // dWn[o][j]  += sum_f gin[f][o] q[f][j]
  }
  float* part = a.part + (size_t)blockIdx.x * kTailBwdPart;
  store_wacc(part, accwo, fg, og);
  store_colsum(part + 4096, sbo, sRed, fg, og, tid);
  store_wacc(part + 4160, accwn, fg, og);
  store_colsum(part + 4160 + 4096, sbn, sRed, fg, og, tid);
}

// --------------------------------------------------------------------------------------------
// Stage-1 projection weight gradient: dW (64, dim) = sum_n g0[n] x[n]^T, db = sum_n g0[n]
// (all padded frames included: the projection is unmasked).  grid = (kchunks, splits).
// --------------------------------------------------------------------------------------------
struct ProjBwdArgs {
  const float* x; const float* g; float* part;   // per split: [64][kchunks*64] then [64] bias
  int64_t n_frames; int dim, kchunks, num_tiles;
};

__global__ void __launch_bounds__(NT, 4) proj_bwd_kernel(ProjBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sG = smem;
  float* sX = sG + TILE;
  float* sRed = sX + TILE;
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  const int kc = blockIdx.x, sp = blockIdx.y, splits = gridDim.y;
  float accw[8][4] = {};
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tile = sp; tile < a.num_tiles; tile += splits) {
    const int64_t n0 = (int64_t)tile * TF;
    __syncthreads();
    load_tile_cols(sG, a.g, n0, a.n_frames, C, 0, tid);
    load_tile_cols(sX, a.x, n0, a.n_frames, a.dim, kc * 64, tid);
    __syncthreads();
    if (kc == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g = ld4s(sG + swz(fg + 8 * j, og));
        sb[0] += g.x; sb[1] += g.y; sb[2] += g.z; sb[3] += g.w;
      }
    }
    wgemm(sG, sX, accw, fg, og);     // dW[o][kc*64+k] += sum_f g[f][o] x[f][kc*64+k]
  }
  const int ldp = a.kchunks * 64;
  float* part = a.part + (size_t)sp * (64 * (size_t)ldp + 64);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    reinterpret_cast<float4*>(part + (size_t)(8 * fg + i) * ldp + kc * 64)[og] =
        make_float4(accw[i][0], accw[i][1], accw[i][2], accw[i][3]);
  if (kc == 0) store_colsum(part + 64 * (size_t)ldp, sb, sRed, fg, og, tid);
}

// --------------------------------------------------------------------------------------------
// Cross-grid reduction of per-CTA partials into the flat native-layout gradient buffer.
// --------------------------------------------------------------------------------------------
struct ReduceSeg {
  const float* src; float* dst;
  int64_t stride;            // floats between consecutive partials
  int P, rows, cols_src, cols_dst;
  int mode;                  // 0: dst[r*cols_dst + c];  1: conv_dilated: src (3,out,in) -> dst[(o*64 + c)*3 + k]
};
struct ReduceArgs { ReduceSeg seg[6]; int nseg; int accumulate; };

// 256 threads = 32 consecutive output elements x 8 partial lanes: lane y sums partials y, y+8, ...
// (coalesced across x), then the 8 lane sums are added in fixed order -> deterministic.
// The reductions are launched programmatically behind the kernel that wrote their partials (when programmatic launches are
// on): they let their own dependents start, then wait for that grid -- the launch gap disappears, the ordering stays.
__device__ __forceinline__ void reduce_pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(ReduceArgs a) {
  __shared__ float red[8][33];
  reduce_pdl_prologue();
  const ReduceSeg& s = a.seg[blockIdx.y];
  const int total = s.rows * s.cols_dst;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
    const int i = base + tx;
    float v = 0.f;
    int r = 0, c = 0;
    if (i < total) {
      r = i / s.cols_dst; c = i - r * s.cols_dst;
      const float* p = s.src + (size_t)r * s.cols_src + c;
      int k = ty;
      for (; k + 24 < s.P; k += 32) {
        const float v0 = p[(size_t)k * s.stride], v1 = p[(size_t)(k + 8) * s.stride];
        const float v2 = p[(size_t)(k + 16) * s.stride], v3 = p[(size_t)(k + 24) * s.stride];
        v += (v0 + v1) + (v2 + v3);
      }
      for (; k < s.P; k += 8) v += p[(size_t)k * s.stride];
    }
    red[ty][tx] = v;
    __syncthreads();
    if (ty == 0 && i < total) {
      float t = red[0][tx];
#pragma unroll
      for (int g = 1; g < 8; ++g) t += red[g][tx];
      size_t di = i;
      if (s.mode == 1) { const int tap = r >> 6, o = r & 63; di = ((size_t)o * 64 + c) * 3 + tap; }
      s.dst[di] = a.accumulate ? s.dst[di] + t : t;
    }
    __syncthreads();
  }
}

// All dilated layers of one stage at once (tensor-core path): every layer left one tc_wgrad partial
// per CTA at src0 + l*layer_src_stride; blockIdx.y = layer, one thread per gradient element
// (conv_dilated.weight, conv_1x1.weight, conv_dilated.bias, conv_1x1.bias).
struct ReduceLayersArgs {
  const float* src0; float* dst0;          // dst0 = the stage's first conv_dilated.weight inside the flat gradient buffer
  int64_t layer_src_stride, layer_dst_stride, part_stride;
  int P, accumulate;
};

__global__ void __launch_bounds__(256) reduce_layers_kernel(ReduceLayersArgs a) {
  // one thread per gradient element of the layer (12288 conv_dilated.weight | 4096 conv_1x1.weight | 64 + 64 biases):
  // the P partials are summed in fixed order (deterministic), consecutive threads read consecutive addresses
  reduce_pdl_prologue();
  const int l = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 12288 + 4096 + 128) return;
  int src_off, i;
  size_t di;
  if (e < 12288) {                         // partial rows are [tap][out][in]; native layout is (out, in, tap)
    i = e; src_off = 0;
    const int r = i >> 6, c = i & 63, tap = r >> 6, o = r & 63;
    di = ((size_t)o * 64 + c) * 3 + tap;
  } else if (e < 12288 + 4096) {
    i = e - 12288; src_off = 3 * 4096; di = 12288 + 64 + i;
  } else if (e < 12288 + 4096 + 64) {
    i = e - 12288 - 4096; src_off = 4 * 4096 + 64; di = 12288 + i;                  // conv_dilated.bias <- tap 1's column sums
  } else {
    i = e - 12288 - 4096 - 64; src_off = 4 * 4096 + 192; di = 12288 + 64 + 4096 + i;   // conv_1x1.bias <- tap 3's
  }
  const float* p = a.src0 + (size_t)l * a.layer_src_stride + src_off + i;
  float v = 0.f;
  for (int k = 0; k < a.P; ++k) v += p[(size_t)k * a.part_stride];
  float* dst = a.dst0 + (size_t)l * a.layer_dst_stride + di;
  *dst = a.accumulate ? *dst + v : v;
}

}  // namespace mstcn
