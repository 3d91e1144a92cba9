// Shared device helpers for the exact-fp32 MS-TCN kernels (sm_100a).
//
// Tiling vocabulary: a TILE is TF=64 consecutive frames of ONE video, 64 channels wide,
// kept in shared memory as 64 rows x 16 float4 chunks with an XOR swizzle
// (chunk q of row r lives at chunk q ^ (r & 15)), so both the row-wise loaders/writers and
// the GEMM readers below are bank-conflict free without padding bytes.
// A CTA has NT=128 threads; thread (fg = tid>>4, og = tid&15) owns the 8x4 register tile
// {frames fg+8j, j<8} x {channels 4og..4og+3}.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mstcn {

constexpr int C = 64;            // num_f_maps
constexpr int TF = 64;           // frames per tile
constexpr int NT = 128;          // threads per CTA
constexpr int TILE = TF * C;     // floats per tile (16 KB)
constexpr int KMAX = 64;         // largest n_class

__device__ __forceinline__ int swz(int row, int q) { return row * C + (((q ^ row) & 15) << 2); }

__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4s(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- Philox4x32-10: dropout keep-bits (mirrors oracle/mstcn_oracle.py::dropout_keep_bits) ----
__device__ __forceinline__ uint2 dropout_bits(uint64_t seed, uint64_t offset, uint32_t layer, uint32_t frame) {
  uint32_t c0 = frame, c1 = layer, c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint2(c0, c1);
}
// multiplier {0,2} for the 4 channels 4og..4og+3 of a frame whose keep words are `bits`
__device__ __forceinline__ float4 dropout_scale4(uint2 bits, int og) {
  uint32_t w = (og < 8) ? bits.x : bits.y;
  uint32_t s = w >> ((og & 7) << 2);
  return make_float4((s & 1u) ? 2.f : 0.f, (s & 2u) ? 2.f : 0.f, (s & 4u) ? 2.f : 0.f, (s & 8u) ? 2.f : 0.f);
}

// ---- tile movers ---------------------------------------------------------------------------
// rows t_first + r (r < TF) of one video (base pointer `vid`, T frames, 64 floats per frame);
// rows outside [0, T) read as zero -- this IS the conv's zero padding (networks.py:339).
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ vid,
                                          int t_first, int T, int tid) {
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = tid + i * NT, r = idx >> 4, q = idx & 15, t = t_first + r;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T) v[i] = __ldg(reinterpret_cast<const float4*>(vid + (size_t)t * C) + q);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = tid + i * NT, r = idx >> 4, q = idx & 15;
    st4s(dst + swz(r, q), v[i]);
  }
}

// generic: rows n_first + r of a flat (n_rows, ld) matrix, columns col0 + 4q.. ; zero outside
__device__ __forceinline__ void load_tile_cols(float* __restrict__ dst, const float* __restrict__ src,
                                               int64_t n_first, int64_t n_rows, int ld, int col0, int tid) {
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = tid + i * NT, r = idx >> 4, q = idx & 15;
    int64_t n = n_first + r;
    int c = col0 + 4 * q;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < n_rows && c < ld) v[i] = __ldg(reinterpret_cast<const float4*>(src + n * ld + c));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = tid + i * NT, r = idx >> 4, q = idx & 15;
    st4s(dst + swz(r, q), v[i]);
  }
}

// plain (unswizzled) copy of n4 float4 from global to shared
__device__ __forceinline__ void load_weights(float* __restrict__ dst, const float* __restrict__ src, int n4, int tid) {
  for (int i = tid; i < n4; i += NT)
    st4s(dst + 4 * i, __ldg(reinterpret_cast<const float4*>(src) + i));
}

__device__ __forceinline__ void zero_rows(float* __restrict__ vid, int t0, int T, int tid) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int idx = tid + i * NT, r = idx >> 4, q = idx & 15, t = t0 + r;
    if (t < T) reinterpret_cast<float4*>(vid + (size_t)t * C)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- register-tile GEMMs ---------------------------------------------------------------------
// frame GEMM: acc[j][i] += sum_k A[fg+8j][k] * W[k][4og+i],  A = swizzled tile, W = (64,64) row-major
__device__ __forceinline__ void fgemm(const float* __restrict__ As, const float* __restrict__ Ws,
                                      float (&acc)[8][4], int fg, int og) {
#pragma unroll 2
  for (int k4 = 0; k4 < 16; ++k4) {
    float4 w0 = ld4s(Ws + (4 * k4 + 0) * C + 4 * og);
    float4 w1 = ld4s(Ws + (4 * k4 + 1) * C + 4 * og);
    float4 w2 = ld4s(Ws + (4 * k4 + 2) * C + 4 * og);
    float4 w3 = ld4s(Ws + (4 * k4 + 3) * C + 4 * og);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 a = ld4s(As + swz(fg + 8 * j, k4));
      acc[j][0] = fmaf(a.x, w0.x, acc[j][0]); acc[j][1] = fmaf(a.x, w0.y, acc[j][1]);
      acc[j][2] = fmaf(a.x, w0.z, acc[j][2]); acc[j][3] = fmaf(a.x, w0.w, acc[j][3]);
      acc[j][0] = fmaf(a.y, w1.x, acc[j][0]); acc[j][1] = fmaf(a.y, w1.y, acc[j][1]);
      acc[j][2] = fmaf(a.y, w1.z, acc[j][2]); acc[j][3] = fmaf(a.y, w1.w, acc[j][3]);
      acc[j][0] = fmaf(a.z, w2.x, acc[j][0]); acc[j][1] = fmaf(a.z, w2.y, acc[j][1]);
      acc[j][2] = fmaf(a.z, w2.z, acc[j][2]); acc[j][3] = fmaf(a.z, w2.w, acc[j][3]);
      acc[j][0] = fmaf(a.w, w3.x, acc[j][0]); acc[j][1] = fmaf(a.w, w3.y, acc[j][1]);
      acc[j][2] = fmaf(a.w, w3.z, acc[j][2]); acc[j][3] = fmaf(a.w, w3.w, acc[j][3]);
    }
  }
}

// weight-gradient GEMM (reduction over the tile's frames):
//   acc[i][n] += sum_f A[f][8mg+i] * B[f][4ng+n],   A, B swizzled tiles; mg = tid>>4, ng = tid&15
__device__ __forceinline__ void wgemm(const float* __restrict__ As, const float* __restrict__ Bs,
                                      float (&acc)[8][4], int mg, int ng) {
#pragma unroll 4
  for (int f = 0; f < TF; ++f) {
    float4 a0 = ld4s(As + swz(f, 2 * mg));
    float4 a1 = ld4s(As + swz(f, 2 * mg + 1));
    float4 b = ld4s(Bs + swz(f, ng));
    float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
      acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
    }
  }
}

// store a wgemm accumulator as a (64,64) row-major partial: row 8mg+i, cols 4ng..4ng+3
__device__ __forceinline__ void store_wacc(float* __restrict__ dst, const float (&acc)[8][4], int mg, int ng) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    reinterpret_cast<float4*>(dst + (8 * mg + i) * C)[ng] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// column sums held per thread (4 channels, summed over the thread's frames) -> (64) partial.
// red: shared scratch of 8*64 floats.  Deterministic order.
__device__ __forceinline__ void store_colsum(float* __restrict__ dst, const float (&s)[4], float* red, int fg, int og, int tid) {
  __syncthreads();
  st4s(red + fg * C + 4 * og, make_float4(s[0], s[1], s[2], s[3]));
  __syncthreads();
  if (tid < C) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g * C + tid];
    dst[tid] = t;
  }
}

// reductions across the 16 lanes that share a frame row (lanes differ in og = tid & 15)
__device__ __forceinline__ float row_max16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mstcn
