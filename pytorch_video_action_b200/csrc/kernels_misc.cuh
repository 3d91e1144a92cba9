// Loss, post-processing, optimizer and parameter packing kernels.
#pragma once
#include "common.cuh"
#include "layout.h"

namespace mstcn {

// --------------------------------------------------------------------------------------------
// Parameter packing: native (out,in,tap) state_dict layout -> GEMM-friendly operands.
// grid.y = stage; every packed element of the stage is produced by exactly one thread.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_params_kernel(Layout lay, const float* __restrict__ p, float* __restrict__ q) {
  const int s = blockIdx.y;
  const int K = lay.K, din = lay.din(s), dinp = lay.dinp(s);
  const int64_t n = lay.pstage_size(s);
  const float* ps = p;              // absolute native offsets below
  float* qs = q + lay.pstage_off(s);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    float v = 0.f;
    const int64_t n_win_t = 64LL * dinp;
    if (r < n_win_t) {                       // win_t (dinp, 64): [k][o] = W[o][k]
      int k = (int)(r >> 6), o = (int)(r & 63);
      if (k < din) v = ps[lay.win_w(s) + (int64_t)o * din + k];
    } else if ((r -= n_win_t) < 64) {        // bin
      v = ps[lay.win_b(s) + r];
    } else if ((r -= 64) < 4096) {           // win_b (64 out, 64 class-padded): native rows padded
      int o = (int)(r >> 6), j = (int)(r & 63);
      if (s > 0 && j < din) v = ps[lay.win_w(s) + (int64_t)o * din + j];
    } else if ((r -= 4096) < (int64_t)lay.L * Layout::kLayerPacked) {
      const int l = (int)(r / Layout::kLayerPacked);
      r -= (int64_t)l * Layout::kLayerPacked;
      if (r < 12288) {                       // wd_t (3, in, out)
        int k = (int)(r >> 12), c = (int)((r >> 6) & 63), o = (int)(r & 63);
        v = ps[lay.wd(s, l) + ((int64_t)o * 64 + c) * 3 + k];
      } else if ((r -= 12288) < 64) {
        v = ps[lay.bd(s, l) + r];
      } else if ((r -= 64) < 4096) {         // w1_t (in, out)
        int c = (int)(r >> 6), o = (int)(r & 63);
        v = ps[lay.w1(s, l) + (int64_t)o * 64 + c];
      } else if ((r -= 4096) < 64) {
        v = ps[lay.b1(s, l) + r];
      } else if ((r -= 64) < 12288) {        // wd_b (3, out, in)
        int k = (int)(r >> 12), o = (int)((r >> 6) & 63), c = (int)(r & 63);
        v = ps[lay.wd(s, l) + ((int64_t)o * 64 + c) * 3 + k];
      } else {                               // w1_n (out, in) native copy
        r -= 12288;
        v = ps[lay.w1(s, l) + r];
      }
    } else {
      r -= (int64_t)lay.L * Layout::kLayerPacked;
      if (r < 4096) {                        // wout_t (64 in, 64 class-padded)
        int c = (int)(r >> 6), j = (int)(r & 63);
        if (j < K) v = ps[lay.wout(s) + (int64_t)j * 64 + c];
      } else if ((r -= 4096) < 64) {         // bout padded
        if (r < K) v = ps[lay.bout(s) + r];
      } else {                               // wout_b (64 class-padded, 64 in)
        r -= 64;
        int j = (int)(r >> 6), c = (int)(r & 63);
        if (j < K) v = ps[lay.wout(s) + (int64_t)j * 64 + c];
      }
    }
    qs[i] = v;
  }
}

// --------------------------------------------------------------------------------------------
// nn.CrossEntropyLoss(ignore_index=-1) forward + backward (train.py:266-267,326).
// One warp per row; gout = softmax - onehot on valid rows (unnormalised), 0 on ignored rows.
// Block partials (sum nll, count) -> scratch; ce_finalize_kernel reduces them in fixed order.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ z, const int64_t* __restrict__ y,
                                                      int64_t n_rows, int K, float* __restrict__ gout,
                                                      float* __restrict__ scratch) {
  __shared__ float s_sum[8], s_cnt[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float my_sum = 0.f, my_cnt = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n_rows; row += (int64_t)gridDim.x * 8) {
    const int64_t lab = y[row];
    const float* zr = z + row * K;
    float* gr = gout + row * K;
    if (lab < 0 || lab >= K) {                  // ignore_index (-1): no loss, zero gradient
      for (int c = lane; c < K; c += 32) gr[c] = 0.f;
      continue;
    }
    float v0 = lane < K ? zr[lane] : -INFINITY;
    float v1 = lane + 32 < K ? zr[lane + 32] : -INFINITY;
    float mx = fmaxf(v0, v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e0 = lane < K ? expf(v0 - mx) : 0.f, e1 = lane + 32 < K ? expf(v1 - mx) : 0.f;
    float sum = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    if (lane < K) gr[lane] = e0 * inv - (lane == lab ? 1.f : 0.f);
    if (lane + 32 < K) gr[lane + 32] = e1 * inv - (lane + 32 == lab ? 1.f : 0.f);
    if (lane == 0) { my_sum += (mx + logf(sum)) - zr[lab]; my_cnt += 1.f; }
  }
  if (lane == 0) { s_sum[warp] = my_sum; s_cnt[warp] = my_cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a += s_sum[w]; c += s_cnt[w]; }
    scratch[2 * blockIdx.x] = a;
    scratch[2 * blockIdx.x + 1] = c;
  }
}

__global__ void __launch_bounds__(256) ce_finalize_kernel(const float* __restrict__ scratch, int nblocks,
                                                          int64_t n_valid_override, float* __restrict__ result) {
  __shared__ double sa[256], sc[256];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { a += scratch[2 * i]; c += scratch[2 * i + 1]; }
  sa[threadIdx.x] = a; sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {          // fixed-shape tree: deterministic
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  a = sa[0]; c = sc[0];
  const double div = n_valid_override > 0 ? (double)n_valid_override : c;
  result[0] = div > 0 ? (float)(a / div) : 0.f;
  result[1] = div > 0 ? (float)(1.0 / div) : 0.f;
  result[2] = (float)c;
}

// --------------------------------------------------------------------------------------------
// Per-frame argmax: torch.max(outputs.data, 1) (train.py:157, inference.py:123), first index on ties.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_argmax_kernel(const float* __restrict__ z, int64_t n_rows, int K,
                                                           int64_t* __restrict__ idx, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const float* zr = z + row * K;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int c = lane; c < K; c += 32) {
    const float v = zr[c];
    if (v > best || bi == 0x7fffffff) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if (lane == 0) { idx[row] = bi; if (val) val[row] = best; }
}

// --------------------------------------------------------------------------------------------
// Segment majority vote: argmax(bincount(pred[s:e])), lowest class wins ties (train.py:161-170);
// inference_fallback adds inference.py:147-151 (class 0 -> argsort(bincount)[1], stable ascending
// over classes 0..max(pred)).  One CTA per segment, shared-memory histogram.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) segment_vote_kernel(const int64_t* __restrict__ pred, const int* __restrict__ bounds,
                                                           int K, int fallback, int* __restrict__ labels) {
  __shared__ int hist[KMAX];
  const int seg = blockIdx.x;
  for (int c = threadIdx.x; c < KMAX; c += blockDim.x) hist[c] = 0;
  __syncthreads();
  const int s = bounds[seg], e = bounds[seg + 1];
  for (int t = s + threadIdx.x; t < e; t += blockDim.x) {
    const int64_t p = pred[t];
    if (p >= 0 && p < K) atomicAdd(&hist[(int)p], 1);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  int best = 0, bc = -1, maxc = 0;
  for (int c = 0; c < K; ++c) {
    if (hist[c] > bc) { bc = hist[c]; best = c; }
    if (hist[c] > 0) maxc = c;
  }
  if (fallback && best == 0 && maxc > 0) {
    // second element of a stable ascending sort of hist[0..maxc]
    int i0 = 0;
    for (int c = 1; c <= maxc; ++c) if (hist[c] < hist[i0]) i0 = c;
    int i1 = -1;
    for (int c = 0; c <= maxc; ++c) {
      if (c == i0) continue;
      if (i1 < 0 || hist[c] < hist[i1]) i1 = c;
    }
    best = i1;
  }
  labels[seg] = best;
}

// --------------------------------------------------------------------------------------------
// torch.optim.Adam (train.py:273,329): lerp first moment, addcmul second, eps outside the sqrt.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float step_size, float one_minus_b1, float b2, float one_minus_b2,
                                                   float bc2_sqrt, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + one_minus_b1 * (gi - m[i]);
    const float vi = v[i] * b2 + one_minus_b2 * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

__global__ void dropout_scale_kernel(uint64_t seed, uint64_t offset, const unsigned long long* offset_dev, uint32_t layer,
                                     int64_t n, float* __restrict__ out) {
  if (offset_dev) offset += __ldg(offset_dev);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (frame, 4-channel group)
  if (i >= n * 16) return;
  const int64_t f = i >> 4; const int og = (int)(i & 15);
  const uint2 bits = dropout_bits(seed, offset, layer, (uint32_t)f);
  reinterpret_cast<float4*>(out + f * C)[og] = dropout_scale4(bits, og);
}

}  // namespace mstcn
