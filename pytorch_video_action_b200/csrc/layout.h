// Flat parameter layout (state_dict order, native nn.Conv1d (out,in,tap) layout -- SURVEY.md 8b)
// and the packed kernel-side layout.  Shared by host and device code.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MSTCN_HD __host__ __device__ __forceinline__
#else
#define MSTCN_HD inline
#endif

namespace mstcn {

struct Layout {
  int dim, S, L, K;

  // ---- native flat buffer --------------------------------------------------------------
  // stage: conv_1x1.weight (64,din,1) | conv_1x1.bias (64) | L x { conv_dilated.weight (64,64,3) |
  //        conv_dilated.bias (64) | conv_1x1.weight (64,64,1) | conv_1x1.bias (64) } |
  //        conv_out.weight (K,64,1) | conv_out.bias (K)
  static constexpr int64_t kLayerParams = 12288 + 64 + 4096 + 64;  // 16512
  MSTCN_HD int din(int s) const { return s == 0 ? dim : K; }
  MSTCN_HD int64_t stage_size(int s) const { return 64LL * din(s) + 64 + (int64_t)L * kLayerParams + 64LL * K + K; }
  MSTCN_HD int64_t stage_off(int s) const { return s == 0 ? 0 : stage_size(0) + (int64_t)(s - 1) * stage_size(1); }
  MSTCN_HD int64_t total() const { return stage_off(S - 1) + stage_size(S - 1); }
  MSTCN_HD int64_t win_w(int s) const { return stage_off(s); }
  MSTCN_HD int64_t win_b(int s) const { return stage_off(s) + 64LL * din(s); }
  MSTCN_HD int64_t layer(int s, int l) const { return win_b(s) + 64 + (int64_t)l * kLayerParams; }
  MSTCN_HD int64_t wd(int s, int l) const { return layer(s, l); }
  MSTCN_HD int64_t bd(int s, int l) const { return layer(s, l) + 12288; }
  MSTCN_HD int64_t w1(int s, int l) const { return layer(s, l) + 12288 + 64; }
  MSTCN_HD int64_t b1(int s, int l) const { return layer(s, l) + 12288 + 64 + 4096; }
  MSTCN_HD int64_t wout(int s) const { return layer(s, L); }
  MSTCN_HD int64_t bout(int s) const { return wout(s) + 64LL * K; }
  MSTCN_HD int tensors() const { return S * (4 + 4 * L); }

  // ---- packed buffer -------------------------------------------------------------------
  // stage: win_t (dinp,64) | bin (64) | win_b (64,64pad) | L x { wd_t (3,64in,64out) | bd (64) |
  //        w1_t (64in,64out) | b1 (64) | wd_b (3,64out,64in) | w1_n (64out,64in) } |
  //        wout_t (64in,64pad) | bout (64pad) | wout_b (64pad,64in)
  // dinp = dim for stage 0, 64 (zero-padded class rows) for later stages.
  static constexpr int64_t kLayerPacked = 12288 + 64 + 4096 + 64 + 12288 + 4096;  // 32896
  MSTCN_HD int dinp(int s) const { return s == 0 ? dim : 64; }
  MSTCN_HD int64_t pstage_size(int s) const { return 64LL * dinp(s) + 64 + 4096 + (int64_t)L * kLayerPacked + 4096 + 64 + 4096; }
  MSTCN_HD int64_t pstage_off(int s) const { return s == 0 ? 0 : pstage_size(0) + (int64_t)(s - 1) * pstage_size(1); }
  MSTCN_HD int64_t ptotal() const { return pstage_off(S - 1) + pstage_size(S - 1); }
  MSTCN_HD int64_t p_win_t(int s) const { return pstage_off(s); }
  MSTCN_HD int64_t p_bin(int s) const { return pstage_off(s) + 64LL * dinp(s); }
  MSTCN_HD int64_t p_win_b(int s) const { return p_bin(s) + 64; }
  MSTCN_HD int64_t p_layer(int s, int l) const { return p_win_b(s) + 4096 + (int64_t)l * kLayerPacked; }
  MSTCN_HD int64_t p_wd_t(int s, int l) const { return p_layer(s, l); }
  MSTCN_HD int64_t p_bd(int s, int l) const { return p_layer(s, l) + 12288; }
  MSTCN_HD int64_t p_w1_t(int s, int l) const { return p_layer(s, l) + 12288 + 64; }
  MSTCN_HD int64_t p_b1(int s, int l) const { return p_layer(s, l) + 12288 + 64 + 4096; }
  MSTCN_HD int64_t p_wd_b(int s, int l) const { return p_layer(s, l) + 12288 + 64 + 4096 + 64; }
  MSTCN_HD int64_t p_w1_n(int s, int l) const { return p_layer(s, l) + 12288 + 64 + 4096 + 64 + 12288; }
  MSTCN_HD int64_t p_wout_t(int s) const { return p_layer(s, L); }
  MSTCN_HD int64_t p_bout(int s) const { return p_wout_t(s) + 4096; }
  MSTCN_HD int64_t p_wout_b(int s) const { return p_bout(s) + 64; }
  // tensor-core operand images (hi/lo TF32 split, UMMA SWIZZLE_128B layout), one per dilated layer,
  // appended after the fp32 operands: [Wd_hi | Wd_lo | W1_hi | W1_lo] = 32768 floats
  // followed by the backward images [WdT_hi | WdT_lo | W1T_hi | W1T_lo] of the same size
  static constexpr int64_t kTcLayerImage = 2 * 32768;
  MSTCN_HD int64_t p_tc(int s, int l) const { return ptotal() + ((int64_t)s * L + l) * kTcLayerImage; }
  MSTCN_HD int64_t p_tcb(int s, int l) const { return p_tc(s, l) + 32768; }
  // then, per stage, the tail's forward and backward images (same shape)
  MSTCN_HD int64_t p_tt(int s) const { return ptotal() + (int64_t)S * L * kTcLayerImage + (int64_t)s * kTcLayerImage; }
  MSTCN_HD int64_t p_ttb(int s) const { return p_tt(s) + 32768; }
  // then the stage-1 input projection's image: per 32-feature K-block [W_hi 64x32 | W_lo 64x32] = 4096 floats
  MSTCN_HD int proj_kblocks() const { return (dim + 31) / 32; }
  MSTCN_HD int64_t p_tp() const { return ptotal() + (int64_t)S * (L + 1) * kTcLayerImage; }
  MSTCN_HD int64_t ptotal_with_tc() const { return p_tp() + (int64_t)proj_kblocks() * 4096; }
};

}  // namespace mstcn
