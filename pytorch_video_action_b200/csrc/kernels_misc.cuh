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

// the biases only (what the tensor-core path reads from the fp32 operand area): bin | bd, b1 per layer | bout (padded)
__global__ void __launch_bounds__(256) pack_biases_kernel(Layout lay, const float* __restrict__ p, float* __restrict__ q) {
  const int per_stage = 64 * (2 + 2 * lay.L);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lay.S * per_stage; i += gridDim.x * blockDim.x) {
    const int s = i / per_stage, r = i - s * per_stage, g = r >> 6, c = r & 63;
    if (g == 0) q[lay.p_bin(s) + c] = p[lay.win_b(s) + c];
    else if (g == 1) q[lay.p_bout(s) + c] = c < lay.K ? p[lay.bout(s) + c] : 0.f;
    else {
      const int l = (g - 2) >> 1;
      if ((g & 1) == 0) q[lay.p_bd(s, l) + c] = p[lay.bd(s, l) + c];
      else q[lay.p_b1(s, l) + c] = p[lay.b1(s, l) + c];
    }
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
      // any other label outside [0, K) is a caller error (nn.CrossEntropyLoss asserts on the device): the loss turns NaN
      if (lab != -1 && lane == 0) my_sum += __int_as_float(0x7fc00000);
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
  // every row ignored: NaN, like torch's mean over an empty set (0 / 0)
  result[0] = div > 0 ? (float)(a / div) : __int_as_float(0x7fc00000);
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

// The same step with its state on the device, so that a captured CUDA graph can replay it: the step count lives in
// *step_dev (advanced by adam_bump_kernel behind this kernel) and the learning rate in *lr_dev (rewritten by the host,
// in stream order, when a scheduler changes it).  Bias corrections in double, once per block.
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                       const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                                       const long long* __restrict__ step_dev) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double step = (double)(*step_dev + 1);
    s_step_size = (float)((double)*lr_dev / (1.0 - pow((double)b1, step)));
    s_bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt, one_minus_b1 = 1.f - b1, one_minus_b2 = 1.f - b2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + one_minus_b1 * (gi - m[i]);
    const float vi = v[i] * b2 + one_minus_b2 * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}
__global__ void adam_bump_kernel(long long* step_dev) { *step_dev += 1; }

__global__ void dropout_scale_kernel(uint64_t seed, uint64_t offset, const unsigned long long* offset_dev, uint32_t layer,
                                     int64_t n, float* __restrict__ out) {
  if (offset_dev) offset += __ldg(offset_dev);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (frame, 4-channel group)
  if (i >= n * 16) return;
  const int64_t f = i >> 4; const int og = (int)(i & 15);
  const uint2 bits = dropout_bits(seed, offset, layer, (uint32_t)f);
  reinterpret_cast<float4*>(out + f * C)[og] = dropout_scale4(bits, og);
}


// --------------------------------------------------------------------------------------------
// pad_batch on the device (train.py:183-205, inference.py:32-44): the dataset's frames live concatenated in HBM
// (feats (sum T_i, D), labels (sum T_i,), offsets (V+1,)); a batch is gathered straight into the padded
// (B, T, D) feature tensor (zeros beyond a video's length) and the flat (B*T,) target (-1 beyond it).
// One warp per output frame row; float4 copies (D % 4 == 0).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pad_batch_kernel(const float* __restrict__ feats, const int64_t* __restrict__ labels,
                                                        const int64_t* __restrict__ offsets, const int* __restrict__ video_idx,
                                                        int B, int T, int D, float* __restrict__ x, int64_t* __restrict__ y,
                                                        int* __restrict__ lens_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= (int64_t)B * T) return;
  const int b = (int)(row / T), t = (int)(row - (int64_t)b * T);
  const int v = video_idx[b];
  const int64_t o0 = offsets[v], len = offsets[v + 1] - o0;
  float4* dst = reinterpret_cast<float4*>(x + row * D);
  if (t < len) {
    const float4* src = reinterpret_cast<const float4*>(feats + (o0 + t) * D);
    for (int i = lane; i < D / 4; i += 32) dst[i] = __ldg(src + i);
  } else {
    for (int i = lane; i < D / 4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (lane == 0) {
    if (y != nullptr) y[row] = (t < len && labels != nullptr) ? labels[o0 + t] : -1;
    if (lens_out != nullptr && t == 0) lens_out[b] = (int)(len < T ? len : T);
  }
}


// --------------------------------------------------------------------------------------------
// Canonical MS-TCN loss (Farha & Gall, CVPR 2019; NOT in the reference, SURVEY.md 0.3 / 8a-L2 -- parity unpinned):
//   sum over stages s of  CE(z_s, y; ignore -1)  +  lam * mean_{b,t>=1,c}( clamp((logp_s[t] - logp_s[t-1].detach())^2, 0, tau^2) * m[b,t] )
// forward + backward in one pass over the per-stage logits (S, B*T, K): one warp per (stage, frame) row.
// part[2*block] = sum of -logp[y], part[2*block+1] = sum of clamped squares; finalize scales and adds.
// --------------------------------------------------------------------------------------------
struct PaperLossArgs {
  const float* z; const int64_t* labels; const int* lens; float* g; float* part;
  int S, B, T, K;
  float inv_nvalid, tmse_scale, tau2;     // tmse_scale = lam / (B * (T-1) * K)
};

__global__ void __launch_bounds__(256) paper_loss_kernel(PaperLossArgs a) {
  __shared__ float s_ce[8], s_tm[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t N = (int64_t)a.B * a.T, rows = N * a.S;
  const int K = a.K;
  float my_ce = 0.f, my_tm = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const int64_t n = row % N;
    const int b = (int)(n / a.T), t = (int)(n - (int64_t)b * a.T);
    const float* zr = a.z + row * K;
    float* gr = a.g + row * K;
    const bool c0 = lane < K, c1 = lane + 32 < K;
    // log-softmax of this frame
    const float v0 = c0 ? zr[lane] : -INFINITY, v1 = c1 ? zr[lane + 32] : -INFINITY;
    float mx = fmaxf(v0, v1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = (c0 ? expf(v0 - mx) : 0.f) + (c1 ? expf(v1 - mx) : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = mx + logf(sum);
    const float lp0 = v0 - lse, lp1 = v1 - lse;
    const float p0 = c0 ? expf(lp0) : 0.f, p1 = c1 ? expf(lp1) : 0.f;
    float g0 = 0.f, g1 = 0.f;
    const int64_t lab = a.labels[n];
    if (lab >= 0 && lab < K) {                       // cross-entropy term, ignore_index = -1
      g0 = (p0 - (lane == lab ? 1.f : 0.f)) * a.inv_nvalid;
      g1 = (p1 - (lane + 32 == lab ? 1.f : 0.f)) * a.inv_nvalid;
      if (lane == (int)(lab & 31)) my_ce -= (lab < 32 ? lp0 : lp1);
    } else if (lab != -1 && lane == 0) {
      my_ce += __int_as_float(0x7fc00000);           // label outside [0, K) and not ignore_index: the loss turns NaN
    }
    if (t >= 1 && t < __ldg(a.lens + b)) {           // truncated-MSE smoothing term, masked by m[b, t]
      const float* zp = zr - K;
      const float u0 = c0 ? zp[lane] : -INFINITY, u1 = c1 ? zp[lane + 32] : -INFINITY;
      float mp = fmaxf(u0, u1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mp = fmaxf(mp, __shfl_xor_sync(0xffffffffu, mp, o));
      float sp = (c0 ? expf(u0 - mp) : 0.f) + (c1 ? expf(u1 - mp) : 0.f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, o);
      const float lsp = mp + logf(sp);
      const float d0 = c0 ? lp0 - (u0 - lsp) : 0.f, d1 = c1 ? lp1 - (u1 - lsp) : 0.f;
      const float q0 = d0 * d0, q1 = d1 * d1;
      float tm = fminf(q0, a.tau2) + fminf(q1, a.tau2);
      // d clamp(x, 0, tau^2)/dx = 1 on [0, tau^2] (torch.clamp passes the gradient on the closed interval)
      const float e0 = (c0 && q0 <= a.tau2) ? 2.f * d0 * a.tmse_scale : 0.f;
      const float e1 = (c1 && q1 <= a.tau2) ? 2.f * d1 * a.tmse_scale : 0.f;
      float es = e0 + e1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { es += __shfl_xor_sync(0xffffffffu, es, o); tm += __shfl_xor_sync(0xffffffffu, tm, o); }
      g0 += e0 - p0 * es;                            // back through log_softmax of frame t (frame t-1 is detached)
      g1 += e1 - p1 * es;
      if (lane == 0) my_tm += tm;
    }
    if (c0) gr[lane] = g0;
    if (c1) gr[lane + 32] = g1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_ce += __shfl_xor_sync(0xffffffffu, my_ce, o);
  if (lane == 0) { s_ce[warp] = my_ce; s_tm[warp] = my_tm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.f, m = 0.f;
    for (int i = 0; i < 8; ++i) { c += s_ce[i]; m += s_tm[i]; }
    a.part[2 * blockIdx.x] = c; a.part[2 * blockIdx.x + 1] = m;
  }
}

__global__ void __launch_bounds__(256) paper_loss_finalize_kernel(const float* __restrict__ part, int nblocks, float inv_nvalid,
                                                                  float tmse_scale, float* __restrict__ result) {
  __shared__ double sa[256], sc[256];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { a += part[2 * i]; c += part[2 * i + 1]; }
  sa[threadIdx.x] = a; sc[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {          // fixed-shape tree: deterministic
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const double ce = sa[0] * inv_nvalid, tm = sc[0] * tmse_scale;
  result[0] = (float)(ce + tm); result[1] = (float)ce; result[2] = (float)tm;
}

}  // namespace mstcn
