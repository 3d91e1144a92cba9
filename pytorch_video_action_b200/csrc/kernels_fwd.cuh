// Forward kernels of the exact-fp32 path: stage-1 projection, fused dilated residual layer,
// fused stage tail.  One persistent CTA loop per kernel (grid = min(tiles, 2 x SMs)).
#pragma once
#include "common.cuh"

namespace mstcn {

// --------------------------------------------------------------------------------------------
// Stage-1 input projection: y[n][:] = W x[n][:] + b   (SingleStageModel.conv_1x1, networks.py:325,330)
// NOT masked: padded frames get the bias.  x is the caller's (B,T,dim)
// batch-first tensor read in place (no transpose copy, networks.py:306).
// --------------------------------------------------------------------------------------------
struct ProjFwdArgs {
  const float* x; const float* w_t; const float* bias; float* y;
  int64_t n_frames; int dim; int num_tiles;
};

__global__ void __launch_bounds__(NT, 4) proj_fwd_kernel(ProjFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sX = smem;            // swizzled (64 frames, 64 k)
  float* sW = smem + TILE;     // (64 k, 64 out)
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias) + og);
  const int kchunks = (a.dim + 63) / 64;
  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int64_t n0 = (int64_t)tile * TF;
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w; }
    for (int kc = 0; kc < kchunks; ++kc) {
      __syncthreads();
      load_tile_cols(sX, a.x, n0, a.n_frames, a.dim, kc * 64, tid);
      for (int i = tid; i < 1024; i += NT) {          // W chunk rows kc*64 + r; zero beyond dim
        int r = i >> 4, k = kc * 64 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < a.dim) v = __ldg(reinterpret_cast<const float4*>(a.w_t + (size_t)k * C) + (i & 15));
        st4s(sW + 4 * i, v);
      }
      __syncthreads();
      fgemm(sX, sW, acc, fg, og);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int64_t n = n0 + fg + 8 * j;
      if (n < a.n_frames)
        reinterpret_cast<float4*>(a.y + n * C)[og] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
  }
}

// --------------------------------------------------------------------------------------------
// DilatedResidualLayer.forward (networks.py:343-347), one kernel:
//   u = bd + sum_k Wd[:,:,k] x[t+(k-1)d]   (3 halo tiles in smem; OOB rows = conv zero padding)
//   h = relu(u) (kept for backward if h_out), o = W1 h + b1, y = (x + drop(o)) * [t < len]
// --------------------------------------------------------------------------------------------
struct LayerFwdArgs {
  const float* x; float* y; float* h; const int* lens;
  const float* wd_t; const float* bd; const float* w1_t; const float* b1;
  int B, T, d, tiles_per_video, num_tiles;
  int train; uint32_t layer_id; uint64_t seed, offset;
  const unsigned long long* offset_dev;   // optional device-side step counter added to `offset` (CUDA-graph replay)
  uint32_t frame0;                        // global index of this launch's first frame (video-group launches keep the
                                          // whole-batch frame numbering of the Philox stream)
};

constexpr int kLayerFwdSmem = (3 * TILE + TILE + 3 * TILE) * 4 + 64 * 8;   // Wd, W1, 3 taps, keep-bits

__global__ void __launch_bounds__(NT, 2) layer_fwd_kernel(LayerFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sWd = smem;                  // (3, 64 in, 64 out)
  float* sW1 = sWd + 3 * TILE;        // (64 in, 64 out)
  float* sX = sW1 + TILE;             // 3 swizzled tap tiles; tap 0 is reused for h
  uint2* sBits = reinterpret_cast<uint2*>(sX + 3 * TILE);
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  load_weights(sWd, a.wd_t, 3 * TILE / 4, tid);
  load_weights(sW1, a.w1_t, TILE / 4, tid);
  const float4 bd4 = __ldg(reinterpret_cast<const float4*>(a.bd) + og);
  const float4 b14 = __ldg(reinterpret_cast<const float4*>(a.b1) + og);

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TF;
    const int len = __ldg(a.lens + b);
    const size_t vbase = (size_t)b * a.T * C;
    if (t0 >= len) {                  // whole tile is padding: y = 0 (mask), nothing else is read later
      zero_rows(a.y + vbase, t0, a.T, tid);
      continue;
    }
    // a tap whose 64 rows all fall outside [0,T) contributes exactly 0 (fact 0.4: d >= T)
    const bool tap0 = (t0 + TF - 1 - a.d) >= 0;
    const bool tap2 = (t0 + a.d) < a.T;
    __syncthreads();
    if (tap0) load_tile(sX, a.x + vbase, t0 - a.d, a.T, tid);
    load_tile(sX + TILE, a.x + vbase, t0, a.T, tid);
    if (tap2) load_tile(sX + 2 * TILE, a.x + vbase, t0 + a.d, a.T, tid);
    if (a.train && tid < TF)
      sBits[tid] = dropout_bits(a.seed, a.offset + (a.offset_dev ? __ldg(a.offset_dev) : 0ull), a.layer_id, a.frame0 + (uint32_t)(b * a.T + t0 + tid));
    __syncthreads();

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = bd4.x; acc[j][1] = bd4.y; acc[j][2] = bd4.z; acc[j][3] = bd4.w; }
    if (tap0) fgemm(sX, sWd, acc, fg, og);
    fgemm(sX + TILE, sWd + TILE, acc, fg, og);
    if (tap2) fgemm(sX + 2 * TILE, sWd + 2 * TILE, acc, fg, og);
    __syncthreads();                  // everyone is done reading tap 0
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = fg + 8 * j, t = t0 + r;
      float4 hv = make_float4(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f), fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f));
      st4s(sX + swz(r, og), hv);
      if (a.h != nullptr && t < a.T) reinterpret_cast<float4*>(a.h + vbase + (size_t)t * C)[og] = hv;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = b14.x; acc[j][1] = b14.y; acc[j][2] = b14.z; acc[j][3] = b14.w; }
    fgemm(sX, sW1, acc, fg, og);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = fg + 8 * j, t = t0 + r;
      if (t >= a.T) continue;
      float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      if (a.train) {
        float4 s = dropout_scale4(sBits[r], og);
        o.x *= s.x; o.y *= s.y; o.z *= s.z; o.w *= s.w;
      }
      const float4 xc = ld4s(sX + TILE + swz(r, og));
      const float m = (t < len) ? 1.f : 0.f;
      reinterpret_cast<float4*>(a.y + vbase + (size_t)t * C)[og] =
          make_float4((xc.x + o.x) * m, (xc.y + o.y) * m, (xc.z + o.z) * m, (xc.w + o.w) * m);
    }
  }
}

// --------------------------------------------------------------------------------------------
// Stage tail: z = (Wout a + bout) * mask (networks.py:333); running max over stages + winner
// (torch.cat / permute / torch.max, :312-319, without the copies); q = softmax(z) * mask (:314);
// next stage's unmasked 1x1: x0' = Wn q + bn (:330).
// --------------------------------------------------------------------------------------------
struct TailFwdArgs {
  const float* a; const int* lens;
  const float* wout_t; const float* bout; float* logits; float* out; uint8_t* winner;
  const float* wn_t; const float* bn; float* next_x0;
  int B, T, K, stage, tiles_per_video, num_tiles;
};

constexpr int kTailFwdSmem = 3 * TILE * 4;

__global__ void __launch_bounds__(NT, 4) tail_fwd_kernel(TailFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sWo = smem;             // (64 in, 64 class-padded)
  float* sWn = sWo + TILE;       // (64 class-padded, 64 out)
  float* sA = sWn + TILE;        // swizzled tile: a, then q
  const int tid = threadIdx.x, fg = tid >> 4, og = tid & 15;
  const bool has_next = a.next_x0 != nullptr;
  load_weights(sWo, a.wout_t, TILE / 4, tid);
  if (has_next) load_weights(sWn, a.wn_t, TILE / 4, tid);
  const float4 bo4 = __ldg(reinterpret_cast<const float4*>(a.bout) + og);
  float4 bn4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has_next) bn4 = __ldg(reinterpret_cast<const float4*>(a.bn) + og);
  const int K = a.K;

  for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_video, t0 = (tile - b * a.tiles_per_video) * TF;
    const int len = __ldg(a.lens + b);
    const size_t fbase = (size_t)b * a.T;
    if (t0 >= len) {
      // padding tile: z = 0 (mask), q = 0, so the next stage's unmasked 1x1 outputs its bias
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = t0 + fg + 8 * j;
        if (t >= a.T) continue;
        const size_t row = (fbase + t) * (size_t)K;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * og + i;
          if (c < K) {
            a.logits[row + c] = 0.f;
            if (a.stage == 0) { a.out[row + c] = 0.f; a.winner[row + c] = 0; }
            else if (0.f > a.out[row + c]) { a.out[row + c] = 0.f; a.winner[row + c] = (uint8_t)a.stage; }
          }
        }
        if (has_next) reinterpret_cast<float4*>(a.next_x0 + (fbase + t) * C)[og] = bn4;
      }
      continue;
    }
    __syncthreads();
    load_tile(sA, a.a + fbase * C, t0, a.T, tid);
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = bo4.x; acc[j][1] = bo4.y; acc[j][2] = bo4.z; acc[j][3] = bo4.w; }
    fgemm(sA, sWo, acc, fg, og);
    __syncthreads();               // a no longer needed; sA becomes q
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = fg + 8 * j, t = t0 + r;
      const float m = (t < len) ? 1.f : 0.f;
      float z[4], e[4];
      float zmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        z[i] = acc[j][i] * m;
        if (4 * og + i < K) zmax = fmaxf(zmax, z[i]);
      }
      if (t < a.T) {
        const size_t row = (fbase + t) * (size_t)K;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 4 * og + i;
          if (c < K) {
            a.logits[row + c] = z[i];
            if (a.stage == 0) { a.out[row + c] = z[i]; a.winner[row + c] = 0; }
            else if (z[i] > a.out[row + c]) { a.out[row + c] = z[i]; a.winner[row + c] = (uint8_t)a.stage; }
          }
        }
      }
      if (has_next) {
        zmax = row_max16(zmax);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { e[i] = (4 * og + i < K) ? expf(z[i] - zmax) : 0.f; sum += e[i]; }
        sum = row_sum16(sum);
        const float sc = m / sum;
        st4s(sA + swz(r, og), make_float4(e[0] * sc, e[1] * sc, e[2] * sc, e[3] * sc));
      }
    }
    if (!has_next) continue;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = bn4.x; acc[j][1] = bn4.y; acc[j][2] = bn4.z; acc[j][3] = bn4.w; }
    fgemm(sA, sWn, acc, fg, og);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = t0 + fg + 8 * j;
      if (t < a.T)
        reinterpret_cast<float4*>(a.next_x0 + (fbase + t) * C)[og] = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
    }
  }
}

}  // namespace mstcn
