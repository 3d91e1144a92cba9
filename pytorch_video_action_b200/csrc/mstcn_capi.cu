// C ABI of libmstcn_b200.so (declared in include/mstcn_b200.h).  Host side: argument checks,
// grid sizing (persistent CTAs, 2 per SM), workspace carving, and the whole-model launch
// sequences that stand in for MultiStageModel.forward (networks.py:305-320) and for what
// autograd replays on loss.backward() (train.py:328).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>

#include "../../include/mstcn_b200.h"
#include "kernels_bwd.cuh"
#include "kernels_fwd.cuh"
#include "kernels_misc.cuh"
#include "layout.h"
#include "kernels_tc.cuh"
#include "kernels_dp.cuh"
#include "kernels_tc_fwd2.cuh"

using namespace mstcn;

namespace {

thread_local std::string g_err;

int fail(const char* fmt, const char* a = "") {
  char buf[512];
  snprintf(buf, sizeof buf, fmt, a);
  g_err = buf;
  return 1;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

// Everything cached below is PER DEVICE (a process may drive several GPUs): SM count, the max-dynamic-smem function
// attribute, the internal streams / events and the chain-launch lane.
constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

int sm_count() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0) return -1;
  if (cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached[dev] = n;
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: set it once per (kernel, device)
int set_smem_impl(const void* k, int bytes) {
  struct Done { const void* k; int dev; };
  static Done done[512];
  static int n_done = 0;
  static std::mutex mu;
  const int dev = current_device();
  if (dev < 0) return fail("no current CUDA device");
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_done; ++i)
    if (done[i].k == k && done[i].dev == dev) return 0;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    g_err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    return 1;
  }
  if (n_done < 512) done[n_done++] = Done{k, dev};
  return 0;
}
template <typename KernelT>
int set_smem(KernelT k, int bytes) { return set_smem_impl(reinterpret_cast<const void*>(k), bytes); }

int check_dims(const mstcn_dims* d) {
  if (!d) return fail("dims is NULL");
  if (d->num_f_maps != MSTCN_C) return fail("num_f_maps must be 64 (kernels are specialised; no fallback)");
  if (d->n_class < 1 || d->n_class > MSTCN_KMAX) return fail("n_class must be in [1, 64]");
  if (d->dim < 4 || d->dim % 4 != 0) return fail("dim must be a positive multiple of 4");
  if (d->num_stages < 1 || d->num_layers < 1 || d->num_layers > 30) return fail("bad num_stages / num_layers");
  return 0;
}

bool use_tc(const mstcn_dims* d) { return (d->flags & MSTCN_FLAG_TENSOR_CORES) != 0; }
bool use_tc_bwd(const mstcn_dims* d) { return use_tc(d) && (d->flags & MSTCN_FLAG_FFMA_BACKWARD) == 0; }

Layout make_layout(const mstcn_dims* d) { return Layout{d->dim, d->num_stages, d->num_layers, d->n_class}; }

int tiles_per_video(int T) { return (T + TF - 1) / TF; }
int persistent_grid(int tiles, int per_sm) {
  int g = sm_count() * per_sm;
  if (g < 1) g = 1;
  return tiles < g ? tiles : g;
}

cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// partial-scratch capacity (floats) every backward kernel may use: grid <= 2*SMs CTAs
int64_t bwd_grid_cap() { return 2LL * (sm_count() > 0 ? sm_count() : 148); }
int64_t layer_bwd_scratch() { return bwd_grid_cap() * (kBwdAPart > kBwdBPart ? kBwdAPart : kBwdBPart); }
int64_t tail_bwd_scratch() { return bwd_grid_cap() * kTailBwdPart; }
int proj_kchunks(int dim) { return (dim + 63) / 64; }
int proj_splits(int dim) {
  int s = (int)(bwd_grid_cap() / proj_kchunks(dim));
  return s < 1 ? 1 : s;
}
int64_t proj_bwd_scratch(int dim) { return (int64_t)proj_splits(dim) * (64LL * proj_kchunks(dim) * 64 + 64); }

int pdl_enabled();
// launch with the programmatic-stream-serialization attribute (when programmatic launches are on): the kernel may start
// while its stream predecessor drains and must call griddepcontrol.wait before it touches that kernel's results
template <typename KernelT, typename... Args>
int launch_pdl_grid(const char* name, KernelT kernel, dim3 grid, dim3 block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled();
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) {
    g_err = std::string(name) + ": " + cudaGetErrorString(e);
    return 1;
  }
  return check_launch(name);
}

int launch_reduce(const ReduceArgs& ra, cudaStream_t st) {
  int maxel = 0;
  for (int i = 0; i < ra.nseg; ++i) {
    int el = ra.seg[i].rows * ra.seg[i].cols_dst;
    if (el > maxel) maxel = el;
  }
  dim3 grid((maxel + 31) / 32, ra.nseg);
  return launch_pdl_grid("reduce_partials_kernel", reduce_partials_kernel, grid, dim3(256), st, ra);
}

ReduceSeg seg(const float* src, float* dst, int64_t stride, int P, int rows, int cols_src, int cols_dst, int mode = 0) {
  ReduceSeg s;
  s.src = src; s.dst = dst; s.stride = stride; s.P = P; s.rows = rows; s.cols_src = cols_src; s.cols_dst = cols_dst; s.mode = mode;
  return s;
}

// ---- workspace carving ---------------------------------------------------------------------
struct Ws {
  int64_t N;             // B*T
  int S, L, K;
  bool training;
  float* base;
  int64_t act_stage;     // floats per stage of activations
  // inference keeps one plane per layer, shared by all stages (the chain launch runs the layers of a stage as a
  // dataflow: a layer's output must not overwrite a plane an earlier layer's straggling tile still reads)
  float* act(int s, int l) const {
    if (training) return base + s * act_stage + (int64_t)l * N * 64;
    return base + (int64_t)l * N * 64;
  }
  float* h(int s, int l) const { return training ? base + s * act_stage + (int64_t)(L + 1 + l) * N * 64 : nullptr; }
  // q(s) = softmax(z_s) * mask, (N, 64) zero-padded classes: the next stage's input, kept for the backward
  float* q(int s) const { return training ? base + s * act_stage + (int64_t)(2 * L + 1) * N * 64 : nullptr; }
  int64_t lg() const { return (N * K + 63) / 64 * 64; }
  float* logits(int s) const {
    return training ? base + s * act_stage + (int64_t)(2 * L + 2) * N * 64 : base + (int64_t)(L + 1) * N * 64 + s * lg();
  }
  // tile flags: [direction 0 fwd / 1 bwd][stage][layer step | tail][tile], at the very end of the workspace
  int64_t num_tiles;
  // per stage L+2 rows.  Forward: rows 0..L-1 = the chain's layer steps, row L = the tail.  Backward: row 0 = tail,
  // row 1 = top layer's gu, rows 2..L = the chain's steps, row L+1 = layer 0's input gradient.
  int64_t flag_count() const { return (2LL * S * (L + 2) * num_tiles + 63) / 64 * 64; }
  float* flag_base;
  int* flags(int bwd, int s) const { return reinterpret_cast<int*>(flag_base) + ((int64_t)bwd * S + s) * (L + 2) * num_tiles; }
  // backward planes, two sets (stage parity: a stage's weight-gradient kernel still reads its set while the
  // next stage's chain fills the other).  Per set: Gl[j], j = 0..L, with Gl[l+1] = gy(l) = dL/d(output of layer l)
  // and Gl[0] = gradient w.r.t. the stage's projection output; then U[l] = gu(l) = dL/d(pre-ReLU of layer l).
  int64_t gset() const { return (int64_t)(2 * L + 2) * N * 64; }
  float* gz(int p) const { return gl(p, 2 * L + 1); }         // tail's dL/dz (N, 64 zero-padded classes)
  float* gr(int s) const { return base + S * act_stage + 2 * gset() + (int64_t)s * N * 64; }   // routed dL/dout per stage (N, 64)
  float* gl(int p, int j) const { return base + S * act_stage + p * gset() + (int64_t)j * N * 64; }
  float* gu(int p, int l) const { return gl(p, L + 1 + l); }
  float* scratch() const { return base + S * act_stage + 2 * gset() + (int64_t)S * N * 64; }
};

// tensor-core backward keeps one tc_wgrad partial set per layer until the stage's batched reduction
constexpr int kMaxGroupsScratch = 4;     // == kMaxGroups
int64_t tc_layer_part_stride() { return (int64_t)kMaxGroupsScratch * (sm_count() > 0 ? sm_count() : 148) * tc::kWgPartFloats; }

// backward scratch = [tail partials of every group | per-layer wgrad partials of every group | proj partials of every group]
int64_t scratch_tail_region() { return (int64_t)kMaxGroupsScratch * tail_bwd_scratch(); }
int64_t scratch_layer_region(const mstcn_dims* d) {
  int64_t a = (int64_t)d->num_layers * tc_layer_part_stride(), b = layer_bwd_scratch();
  return a > b ? a : b;
}
int64_t scratch_proj_region(const mstcn_dims* d) {      // FFMA partials of every group, or the tensor-core launch's per-CTA partials
  const int64_t a = (int64_t)kMaxGroupsScratch * proj_bwd_scratch(d->dim);
  const int64_t b = (int64_t)(sm_count() > 0 ? sm_count() : 148) * tc::kWgPartFloats;
  return a > b ? a : b;
}
constexpr int kTailWgradCtas = 18;      // CTAs the stage weight-gradient launch spends on the 1x1 convolutions around the stage
int64_t scratch_floats(const mstcn_dims* d) {
  return scratch_tail_region() + scratch_layer_region(d) + scratch_proj_region(d) + (int64_t)kTailWgradCtas * tc::kWgPartFloats;
}

Ws carve(const mstcn_dims* d, int B, int T, bool training, float* base) {
  Ws w;
  w.N = (int64_t)B * T; w.S = d->num_stages; w.L = d->num_layers; w.K = d->n_class; w.training = training; w.base = base;
  // round the logits plane up to a multiple of 64 floats so every plane stays 256-byte aligned
  w.act_stage = training ? (int64_t)(2 * w.L + 2) * w.N * 64 + w.lg() : 0;
  w.num_tiles = (int64_t)B * ((T + tc::TM - 1) / tc::TM);
  const int64_t body = training ? w.S * w.act_stage + 2 * w.gset() + (int64_t)w.S * w.N * 64 + scratch_floats(d)
                                : (int64_t)(w.L + 1) * w.N * 64 + w.S * w.lg();
  w.flag_base = base + body;
  return w;
}

// ---- single-kernel launchers ---------------------------------------------------------------
int do_proj_fwd(const float* x, int64_t n, int dim, const float* w_t, const float* bias, float* y, cudaStream_t st) {
  ProjFwdArgs a{x, w_t, bias, y, n, dim, (int)((n + TF - 1) / TF)};
  if (a.num_tiles == 0) return 0;
  proj_fwd_kernel<<<a.num_tiles, NT, 2 * TILE * 4, st>>>(a);
  return check_launch("proj_fwd_kernel");
}

int do_layer_fwd(const float* x, float* y, float* h, const int* lens, int B, int T, int d,
                 const float* wd_t, const float* bd, const float* w1_t, const float* b1,
                 const mstcn_dropout* drop, int layer_id, cudaStream_t st, uint32_t frame0 = 0) {
  LayerFwdArgs a;
  a.frame0 = frame0;
  a.x = x; a.y = y; a.h = h; a.lens = lens; a.wd_t = wd_t; a.bd = bd; a.w1_t = w1_t; a.b1 = b1;
  a.B = B; a.T = T; a.d = d; a.tiles_per_video = tiles_per_video(T); a.num_tiles = a.tiles_per_video * B;
  a.train = drop && drop->enabled; a.layer_id = (uint32_t)layer_id;
  a.seed = drop ? drop->seed : 0; a.offset = drop ? drop->offset : 0;
  a.offset_dev = drop ? reinterpret_cast<const unsigned long long*>(drop->offset_dev) : nullptr;
  if (a.num_tiles == 0) return 0;
  if (set_smem(layer_fwd_kernel, kLayerFwdSmem)) return 1;
  layer_fwd_kernel<<<persistent_grid(a.num_tiles, 2), NT, kLayerFwdSmem, st>>>(a);
  return check_launch("layer_fwd_kernel");
}

int do_layer_bwd_gx_tc(const float* gu, const float* gy, float* gx, const int* lens, int B, int T, int d,
                       const float* wimg_b, cudaStream_t st, const int* flags_in = nullptr, int* flags_out = nullptr);
int do_wgrad_tc(const float* gu, const float* gy, const float* x, const float* h, const int* lens, int B, int T, int d,
                const mstcn_dropout* drop, int layer_id, float* part, int* grid_out, cudaStream_t st, uint32_t frame0);
int do_bwd_gu_tc(const float* gy, const float* h, float* gu, const int* lens, int B, int T, const float* wimg_b,
                 const mstcn_dropout* drop, int layer_id, cudaStream_t st, uint32_t frame0, const int* flags_in = nullptr,
                 int* flags_out = nullptr);


// tc_wimg_b != NULL: the input gradient comes from the tensor-core kernel and the FFMA pass B only
// accumulates the dilated-conv weight gradient
int do_layer_bwd(const float* x, const float* h, const float* gy, float* gx, float* gu, const int* lens,
                 int B, int T, int d, const float* wd_b, const float* w1, const mstcn_dropout* drop, int layer_id,
                 float* gwd, float* gbd, float* gw1, float* gb1, float* scratch, int accumulate, cudaStream_t st,
                 const float* tc_wimg_b = nullptr, int* deferred_grid = nullptr, uint32_t frame0 = 0) {
  const int tpv = tiles_per_video(T), tiles = tpv * B;
  if (tiles == 0) return 0;
  if (set_smem(layer_bwd_a_kernel, kLayerBwdASmem) || set_smem(layer_bwd_b_kernel, kLayerBwdBSmem)) return 1;
  const int grid = persistent_grid(tiles, 2);
  const bool tcp = tc_wimg_b != nullptr;
  LayerBwdAArgs a;
  a.gy = gy; a.h = h; a.gu = gu; a.lens = lens; a.w1 = w1; a.part = scratch;
  a.B = B; a.T = T; a.tiles_per_video = tpv; a.num_tiles = tiles;
  a.train = drop && drop->enabled; a.layer_id = (uint32_t)layer_id;
  a.seed = drop ? drop->seed : 0; a.offset = drop ? drop->offset : 0;
  a.offset_dev = drop ? reinterpret_cast<const unsigned long long*>(drop->offset_dev) : nullptr;
  a.gu_only = tcp;
  a.frame0 = frame0;
  if (tcp) {
    if (do_bwd_gu_tc(gy, h, gu, lens, B, T, tc_wimg_b, drop, layer_id, st, frame0)) return 1;
  } else {
    layer_bwd_a_kernel<<<grid, NT, kLayerBwdASmem, st>>>(a);
    if (check_launch("layer_bwd_a_kernel")) return 1;
  }
  if (tcp) {
    // tensor-core path: gx and all four weight-gradient taps (dWd[0..2], dW1) + bias sums
    if (do_layer_bwd_gx_tc(gu, gy, gx, lens, B, T, d, tc_wimg_b, st)) return 1;
    int wg = 0;
    if (do_wgrad_tc(gu, gy, x, h, lens, B, T, d, drop, layer_id, scratch, &wg, st, frame0)) return 1;
    if (deferred_grid != nullptr) { *deferred_grid = wg; return 0; }    // the caller reduces the whole stage at once
    ReduceArgs r; r.accumulate = accumulate; r.nseg = 4;
    r.seg[0] = seg(scratch, gwd, tc::kWgPartFloats, wg, 192, 64, 64, 1);
    r.seg[1] = seg(scratch + 3 * 4096, gw1, tc::kWgPartFloats, wg, 64, 64, 64);
    r.seg[2] = seg(scratch + 4 * 4096 + 64, gbd, tc::kWgPartFloats, wg, 1, 64, 64);
    r.seg[3] = seg(scratch + 4 * 4096 + 192, gb1, tc::kWgPartFloats, wg, 1, 64, 64);
    return launch_reduce(r, st);
  }
  ReduceArgs ra; ra.accumulate = accumulate; ra.nseg = 3;
  ra.seg[0] = seg(scratch, gw1, kBwdAPart, grid, 64, 64, 64);
  ra.seg[1] = seg(scratch + 4096, gb1, kBwdAPart, grid, 1, 64, 64);
  ra.seg[2] = seg(scratch + 4160, gbd, kBwdAPart, grid, 1, 64, 64);
  if (launch_reduce(ra, st)) return 1;

  LayerBwdBArgs b;
  b.gy = gy; b.gu = gu; b.x = x; b.gx = gx; b.lens = lens; b.wd_b = wd_b; b.part = scratch;
  b.B = B; b.T = T; b.d = d; b.tiles_per_video = tpv; b.num_tiles = tiles;
  b.wgrad_only = 0;
  layer_bwd_b_kernel<<<grid, NT, kLayerBwdBSmem, st>>>(b);
  if (check_launch("layer_bwd_b_kernel")) return 1;
  ReduceArgs rb; rb.accumulate = accumulate; rb.nseg = 1;
  rb.seg[0] = seg(scratch, gwd, kBwdBPart, grid, 192, 64, 64, 1);
  return launch_reduce(rb, st);
}

int do_tail_fwd(const float* a_, const int* lens, int B, int T, int K, int stage, const float* wout_t, const float* bout,
                float* logits, float* out, uint8_t* winner, const float* wn_t, const float* bn, float* next_x0,
                cudaStream_t st) {
  TailFwdArgs a;
  a.a = a_; a.lens = lens; a.wout_t = wout_t; a.bout = bout; a.logits = logits; a.out = out; a.winner = winner;
  a.wn_t = wn_t; a.bn = bn; a.next_x0 = next_x0;
  a.B = B; a.T = T; a.K = K; a.stage = stage; a.tiles_per_video = tiles_per_video(T); a.num_tiles = a.tiles_per_video * B;
  if (a.num_tiles == 0) return 0;
  if (set_smem(tail_fwd_kernel, kTailFwdSmem)) return 1;
  tail_fwd_kernel<<<persistent_grid(a.num_tiles, 2), NT, kTailFwdSmem, st>>>(a);
  return check_launch("tail_fwd_kernel");
}

int do_tail_bwd(const float* a_, const float* logits, const float* gout, const float* gscale, const uint8_t* winner,
                const float* gin, const int* lens, int B, int T, int K, int stage, const float* wout_b, const float* wn_b,
                float* ga, float* gwout, float* gbout, float* gwn, float* gbn, float* scratch, int accumulate,
                cudaStream_t st, int* deferred_grid = nullptr) {
  TailBwdArgs a;
  a.a = a_; a.logits = logits; a.gout = gout; a.gscale = gscale; a.winner = winner; a.gin = gin; a.lens = lens;
  a.wout_b = wout_b; a.wn_b = wn_b; a.ga = ga; a.part = scratch;
  a.B = B; a.T = T; a.K = K; a.stage = stage; a.tiles_per_video = tiles_per_video(T); a.num_tiles = a.tiles_per_video * B;
  if (a.num_tiles == 0) return 0;
  if (set_smem(tail_bwd_kernel, kTailBwdSmem)) return 1;
  const int grid = persistent_grid(a.num_tiles, 2);
  tail_bwd_kernel<<<grid, NT, kTailBwdSmem, st>>>(a);
  if (check_launch("tail_bwd_kernel")) return 1;
  if (deferred_grid != nullptr) { *deferred_grid = grid; return 0; }
  ReduceArgs ra; ra.accumulate = accumulate; ra.nseg = 2;
  ra.seg[0] = seg(scratch, gwout, kTailBwdPart, grid, K, 64, 64);
  ra.seg[1] = seg(scratch + 4096, gbout, kTailBwdPart, grid, 1, 64, K);
  if (gin != nullptr) {
    ra.seg[2] = seg(scratch + 4160, gwn, kTailBwdPart, grid, 64, 64, K);
    ra.seg[3] = seg(scratch + 4160 + 4096, gbn, kTailBwdPart, grid, 1, 64, 64);
    ra.nseg = 4;
  }
  return launch_reduce(ra, st);
}

// measurement only: MSTCN_DBG_SKIP=projbwd,wgrad,... drops the named launches (results are then garbage) -- the step-time
// difference is what that kernel costs on the critical path, overlap included
bool dbg_skip(const char* what) {
  static const char* env = getenv("MSTCN_DBG_SKIP");
  return env != nullptr && strstr(env, what) != nullptr;
}

int do_proj_bwd(const float* x, const float* g, int64_t n, int dim, float* gw, float* gb, float* scratch,
                int accumulate, cudaStream_t st, int* deferred_splits = nullptr) {
  if (dbg_skip("projbwd")) { if (deferred_splits) *deferred_splits = 1; return 0; }
  ProjBwdArgs a;
  a.x = x; a.g = g; a.part = scratch; a.n_frames = n; a.dim = dim; a.kchunks = proj_kchunks(dim);
  a.num_tiles = (int)((n + TF - 1) / TF);
  int splits = proj_splits(dim);
  if (splits > a.num_tiles) splits = a.num_tiles > 0 ? a.num_tiles : 1;
  dim3 grid(a.kchunks, splits);
  proj_bwd_kernel<<<grid, NT, (2 * TILE + 8 * C) * 4, st>>>(a);
  if (check_launch("proj_bwd_kernel")) return 1;
  if (deferred_splits != nullptr) { *deferred_splits = splits; return 0; }
  const int ldp = a.kchunks * 64;
  const int64_t stride = 64LL * ldp + 64;
  ReduceArgs ra; ra.accumulate = accumulate; ra.nseg = 2;
  ra.seg[0] = seg(scratch, gw, stride, splits, 64, ldp, dim);
  ra.seg[1] = seg(scratch + 64LL * ldp, gb, stride, splits, 1, 64, 64);
  return launch_reduce(ra, st);
}


// ---- tensor-core path ----------------------------------------------------------------------
long long* g_tc_dbg = nullptr;
long long* g_tc_trace = nullptr;
// mstcn_debug_backward_timing: per stage [chain ms, wgrad ms] of the latest backward call, measured alone
int g_bwd_timing = 0;
float g_bwd_times[2 * 16] = {};
// keeps the stream busy for `ns` so that the host can queue [event, launch, event] behind it: the interval between the
// two events is then GPU time only (no tensor-map encoding / launch latency of the host in it)
__global__ void timer_delay_kernel(long long ns) {
  long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  do { __nanosleep(200); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
struct StageTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t st;
  bool on;
  StageTimer(cudaStream_t s, cudaStream_t other) : st(s), on(g_bwd_timing != 0) {
    if (!on) return;
    cudaStreamSynchronize(s); cudaStreamSynchronize(other);
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    timer_delay_kernel<<<1, 1, 0, st>>>(300000);
    cudaEventRecord(e0, st);
  }
  void stop(float* out) {
    if (!on) return;
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
};

int fwd_v2_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSTCN_FWD_V2"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

// kernel-to-kernel tile dataflow (consumers skip griddepcontrol.wait and follow the producer kernel's tile flags);
// MSTCN_DF=0 keeps programmatic dependent launch but makes every kernel wait for its predecessor grid
int df_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSTCN_DF"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

// Programmatic dependent launch between consecutive kernels plus the kernel-to-kernel tile dataflow that rides on it
// (3 % of the config-2 step).  MSTCN_PDL=0 turns both off (every launch then waits for its predecessor grid to drain).
// History: for part of round 2 this was off by default because ~0.1-0.6 % of the steps came out with a few wrong
// gradients under it.  The cause was not the launch mechanism: in the two single-GEMM kernel modes (layer-0 gx, last-stage
// tail) nothing ordered the next tile's first MMA after the epilogue's load of the accumulator, and the synchronized tile
// starts under programmatic launches made that window reachable (tools/locate_race.py found whole 128-frame tiles of the
// last stage's logits / of layer 0's gx wrong; fixed in tc_layer_kernel, 0 mismatches in 21 500 alternating-batch replays).
int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSTCN_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}
// L2 promotion of every tensor map: a 128-byte box row of an activation / feature tile always has its other half (or the next
// K-block) fetched next, so 256-byte promotion saves a DRAM / L2 transaction per row: 0.5 % of the step at configs 2, 3 and 4
// (profiles/r02_notes.md).  MSTCN_L2PROMO=128 restores the round-1 setting.
CUtensorMapL2promotion l2_promotion() {
  static const int v = getenv("MSTCN_L2PROMO") ? atoi(getenv("MSTCN_L2PROMO")) : 256;
  return v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// (B, T, 64) fp32 activations seen as a 3-D tensor (channel, frame, video); box = 32 channels x 128
// frames, SWIZZLE_128B; out-of-range frames read as zero.
int encode_act_tensor_map(CUtensorMap* tm, const float* base, int B, int T, int atom32, int nlayers, int64_t layer_stride,
                          int box_rows);

// Activation planes live in a reused workspace, so the same (pointer, B, T) triples come back every
// step: keep the encoded maps in a small per-thread direct-mapped cache.
// atom32 = 1 selects SWIZZLE_128B_ATOM_32B (what a transposed / MN-major tf32 UMMA operand needs)
// nlayers > 0 adds a 4th (layer) dimension: nlayers planes layer_stride floats apart
int make_act_tensor_map(CUtensorMap* tm, const float* base, int B, int T, int atom32 = 0, int nlayers = 0,
                        int64_t layer_stride = 0, int box_rows = tc::TM) {
  struct Entry { const float* base; int B, T, atom32, nlayers, rows; int64_t stride; CUtensorMap tm; };
  constexpr int kEntries = 512;
  thread_local Entry cache[kEntries] = {};
  const uintptr_t key = ((reinterpret_cast<uintptr_t>(base) >> 8) * 4 + atom32 * 2 + (nlayers > 0)) * 0x9E3779B97F4A7C15ull;
  Entry& e = cache[(key >> 40) & (kEntries - 1)];
  if (e.base != base || e.B != B || e.T != T || e.atom32 != atom32 || e.nlayers != nlayers || e.stride != layer_stride ||
      e.rows != box_rows) {
    if (encode_act_tensor_map(&e.tm, base, B, T, atom32, nlayers, layer_stride, box_rows)) { e.base = nullptr; return 1; }
    e.base = base; e.B = B; e.T = T; e.atom32 = atom32; e.nlayers = nlayers; e.stride = layer_stride; e.rows = box_rows;
  }
  *tm = e.tm;
  return 0;
}

int encode_act_tensor_map(CUtensorMap* tm, const float* base, int B, int T, int atom32, int nlayers, int64_t layer_stride,
                          int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {64, (cuuint64_t)T, (cuuint64_t)B, (cuuint64_t)(nlayers > 0 ? nlayers : 1)};
  if (layer_stride <= 0) layer_stride = (int64_t)B * T * 64;
  cuuint64_t strides[3] = {256, (cuuint64_t)T * 256, (cuuint64_t)layer_stride * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, nlayers > 0 ? 4 : 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  l2_promotion(),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof buf, "%d", (int)r);
    return fail("cuTensorMapEncodeTiled failed with CUresult %s", buf);
  }
  return 0;
}

// Chain launch description (see TcLayerFwdArgs): nsteps layers in one persistent launch; nx / ng / nhp = number of
// planes (`plane` floats apart) behind the three tensor maps.
struct TcChain {
  int nsteps = 1, lyr0 = 0, dir = 0, nx = 1, ng = 1, nhp = 1, cx_off = 0, cg_off = 0, chp_off = 0;
  int64_t plane = 0, wimg_stride = 0, bias_stride = 0;
  int* flags = nullptr;
  const int* flags_in = nullptr;     // previous kernel's per-tile flags (skips the grid dependency wait)
  int publish_last = 0;
};

template <int MODE>
int launch_tc_layer(const float* xin, const float* gy, float* yout, float* h, const int* lens, int B, int T, int d,
                    const float* wimg, const float* bd, const float* b1, const mstcn_dropout* drop, int layer_id,
                    cudaStream_t st, uint32_t frame0 = 0, const float* hprev = nullptr, const float* wimg2 = nullptr,
                    float* logits_out = nullptr, int K = 0, const TcChain& ch = TcChain()) {
  if ((reinterpret_cast<uintptr_t>(xin) & 15) != 0) return fail("tc layer: activations must be 16-byte aligned");
  CUtensorMap tm, tg, thp;
  if (make_act_tensor_map(&tm, xin, B, T, 0, ch.nx, ch.plane)) return 1;
  if (MODE != 0) { if (make_act_tensor_map(&tg, gy, B, T, 0, ch.ng, ch.plane)) return 1; }
  else if (make_act_tensor_map(&tg, yout, B, T, 0, ch.nsteps, ch.plane)) return 1;        // mode 0: the y output map
  if (MODE == 2) { if (make_act_tensor_map(&thp, hprev, B, T, 0, ch.nhp, ch.plane)) return 1; }
  else if (MODE == 0 && h != nullptr) { if (make_act_tensor_map(&thp, h, B, T, 0, ch.nsteps, ch.plane)) return 1; }   // h output map
  else { thp = tm; }
  tc::TcLayerFwdArgs a = {};
  a.lens = lens; a.wimg = wimg; a.bd = bd; a.b1 = b1; a.y = yout; a.h = h;
  a.B = B; a.T = T; a.d = (MODE == 0 || MODE == 3) ? d : -d; a.skip_extra = (MODE == 0 || MODE == 3) ? 0 : d;
  a.nsteps = ch.nsteps; a.lyr0 = ch.lyr0; a.lyr_dir = ch.dir; a.d_from_layer = (MODE == 0 || MODE == 2) && ch.flags != nullptr;
  a.cx_off = ch.cx_off; a.cg_off = ch.cg_off; a.chp_off = ch.chp_off;
  a.plane = ch.plane; a.wimg_stride = ch.wimg_stride; a.bias_stride = ch.bias_stride; a.flags = ch.flags;
  a.flags_in = ch.flags_in; a.publish_last = ch.publish_last;
  a.co0_off = 0; a.co1_off = MODE == 2 ? -1 : 0;     // mode 2 writes gx to Gl[l] (tm_g) and gu to U[l-1] (tm_x)
  a.gyp = gy; a.hprev = hprev; a.wimg2 = wimg2; a.logits_out = logits_out; a.K = K;
  a.tiles_per_video = (T + tc::TM - 1) / tc::TM; a.num_tiles = a.tiles_per_video * B;
  a.train = drop && drop->enabled; a.layer_id = (uint32_t)layer_id;
  a.seed = drop ? drop->seed : 0; a.offset = drop ? drop->offset : 0;
  a.offset_dev = drop ? reinterpret_cast<const unsigned long long*>(drop->offset_dev) : nullptr;
  a.dbg = MODE == 0 ? g_tc_dbg : nullptr;
  a.trace = (MODE == 0 && ch.flags != nullptr) ? g_tc_trace : nullptr;
  a.frame0 = frame0;
  if (a.num_tiles == 0) return 0;
  // mode 0: MSTCN_FWD_V2=1 selects the second-generation forward kernel (tap tiles in TMEM, prefetching producer).  Measured
  // (profiles/r02_notes.md): 4.18 instead of 4.48 us per tile at config 3, no gain at configs 2 / 4 -> opt-in
  const bool v2 = MODE == 0 && fwd_v2_enabled();
  if (v2 ? set_smem(tc::tc_fwd2_kernel, tc::kTcFwdSmem) : set_smem(tc::tc_layer_kernel<MODE>, tc::kTcFwdSmem)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(persistent_grid(a.num_tiles * a.nsteps, 1));
  cfg.blockDim = dim3(tc::kTcLayerThreads);
  cfg.dynamicSmemBytes = tc::kTcFwdSmem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled();
  cudaError_t e = v2 ? cudaLaunchKernelEx(&cfg, tc::tc_fwd2_kernel, tm, tg, thp, a)
                     : cudaLaunchKernelEx(&cfg, tc::tc_layer_kernel<MODE>, tm, tg, thp, a);
  if (e != cudaSuccess) {
    g_err = std::string("tc_layer_kernel: ") + cudaGetErrorString(e);
    return 1;
  }
  return check_launch("tc_layer_kernel");
}

int do_layer_fwd_tc(const float* x, float* y, float* h, const int* lens, int B, int T, int d, const float* wimg,
                    const float* bd, const float* b1, const mstcn_dropout* drop, int layer_id, cudaStream_t st,
                    uint32_t frame0 = 0) {
  return launch_tc_layer<0>(x, nullptr, y, h, lens, B, T, d, wimg, bd, b1, drop, layer_id, st, frame0);
}

template <typename KernelT, typename... Args>
int launch_pdl_threads(const char* name, KernelT kernel, int grid, int threads, int smem_bytes, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous kernel's tail
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl_enabled();
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) {
    g_err = std::string(name) + ": " + cudaGetErrorString(e);
    return 1;
  }
  return check_launch(name);
}

template <typename KernelT, typename... Args>
int launch_pdl(const char* name, KernelT kernel, int grid, int smem_bytes, cudaStream_t st, Args... args) {
  return launch_pdl_threads(name, kernel, grid, tc::kTcThreads, smem_bytes, st, args...);
}

int do_bwd_gu_tc(const float* gy, const float* h, float* gu, const int* lens, int B, int T, const float* wimg_b,
                 const mstcn_dropout* drop, int layer_id, cudaStream_t st, uint32_t frame0, const int* flags_in,
                 int* flags_out) {
  CUtensorMap tg, th;
  if (make_act_tensor_map(&tg, gy, B, T) || make_act_tensor_map(&th, h, B, T)) return 1;
  tc::TcBwdGuArgs a;
  a.lens = lens; a.wimg_b = wimg_b; a.gu = gu; a.B = B; a.T = T; a.frame0 = frame0;
  a.flags_in = flags_in; a.flags_out = flags_out;
  a.tiles_per_video = (T + tc::TM - 1) / tc::TM; a.num_tiles = a.tiles_per_video * B;
  a.train = drop && drop->enabled; a.layer_id = (uint32_t)layer_id;
  a.seed = drop ? drop->seed : 0; a.offset = drop ? drop->offset : 0;
  a.offset_dev = drop ? reinterpret_cast<const unsigned long long*>(drop->offset_dev) : nullptr;
  if (set_smem(tc::tc_bwd_gu_kernel, tc::kTcBwdGuSmem)) return 1;
  return launch_pdl("tc_bwd_gu_kernel", tc::tc_bwd_gu_kernel, persistent_grid(a.num_tiles, 1), tc::kTcBwdGuSmem, st, tg, th, a);
}

// nlayers == 1: one layer (dilation d), grid = min(tiles, #SMs) CTAs.  nlayers > 1: all layers of a stage in one
// launch -- the four pointers address layer 0's plane and consecutive layers are `*_stride` floats apart;
// ctas_per_layer CTAs share each layer's tiles, dilation = 1 << layer, dropout id = layer_id + layer.
// Stage mode with tail_ctas > 0 appends CTAs for the 1x1 convolutions around the stage: plane nlayers of gu (= gz) and of
// x (= the stage's last activation) give dWout (tap 0), gy0 (= Gl[0], one plane before gy) and q_prev give the stage's
// input-projection gradient (tap 3, tail_tap_mask bit 3).
int do_wgrad_tc_multi(const float* gu, int64_t gu_stride, const float* gy, int64_t gy_stride, const float* x, int64_t x_stride,
                      const float* h, int64_t h_stride, const int* lens, int B, int T, int d, int nlayers,
                      int ctas_per_layer, const mstcn_dropout* drop, int layer_id, float* part, cudaStream_t st,
                      uint32_t frame0, int tail_ctas = 0, int tail_tap_mask = 0, const float* q_prev = nullptr) {
  if (dbg_skip("wgrad")) return 0;
  CUtensorMap ta0, ta1, tb0, tb1, tq;
  const int nl = nlayers, tl = tail_ctas > 0 ? 1 : 0;
  // with a tail, tm_gy starts at Gl[0] = gy - stride (coordinate layer + 1 for the real layers)
  if (make_act_tensor_map(&ta0, gu, B, T, 1, nl + tl, gu_stride, tc::TW) ||
      make_act_tensor_map(&ta1, gy - (tl ? gy_stride : 0), B, T, 1, nl + tl, gy_stride, tc::TW) ||
      make_act_tensor_map(&tb0, x, B, T, 1, nl + tl, x_stride, tc::TW) || make_act_tensor_map(&tb1, h, B, T, 1, nl, h_stride, tc::TW) ||
      make_act_tensor_map(&tq, q_prev ? q_prev : h, B, T, 1, 1, 0, tc::TW))
    return 1;
  tc::TcWgradArgs a = {};
  a.lens = lens; a.part = part; a.B = B; a.T = T; a.frame0 = frame0;
  a.tiles_per_video = (T + tc::TW - 1) / tc::TW; a.num_tiles = a.tiles_per_video * B; a.d = d;
  a.nlayers = nlayers; a.ctas_per_layer = ctas_per_layer; a.layer0_id = layer_id; a.dil_from_layer = nlayers > 1;
  a.tap_mask = 0xF; a.gy_transform = 1; a.tap3_full_T = 0;
  a.tail_ctas = tail_ctas; a.tail_tap_mask = tail_tap_mask; a.cg_off = tl;
  a.dbg = g_tc_dbg;
  a.train = drop && drop->enabled; a.layer_id = (uint32_t)layer_id;
  a.seed = drop ? drop->seed : 0; a.offset = drop ? drop->offset : 0;
  a.offset_dev = drop ? reinterpret_cast<const unsigned long long*>(drop->offset_dev) : nullptr;
  if (set_smem(tc::tc_wgrad_kernel, tc::kTcWgradSmem)) return 1;
  return launch_pdl("tc_wgrad_kernel", tc::tc_wgrad_kernel, nlayers * ctas_per_layer + tail_ctas, tc::kTcWgradSmem, st, ta0, ta1,
                    tb0, tb1, tq, a);
}

int do_wgrad_tc(const float* gu, const float* gy, const float* x, const float* h, const int* lens, int B, int T, int d,
                const mstcn_dropout* drop, int layer_id, float* part, int* grid_out, cudaStream_t st, uint32_t frame0) {
  const int tiles = (T + tc::TW - 1) / tc::TW * B;
  const int grid = persistent_grid(tiles, 1);
  *grid_out = grid;
  return do_wgrad_tc_multi(gu, 0, gy, 0, x, 0, h, 0, lens, B, T, d, 1, grid, drop, layer_id, part, st, frame0);
}

// Stage-1 input projection, weight / bias gradient on the tensor cores (networks.py:325,330 under loss.backward()):
//   dW[o][c] = sum over ALL frames of g0[t][o] * x[t][c],  db[o] = sum_t g0[t][o]   (the conv is unmasked)
// = tc_wgrad_kernel in projection mode: the caller's (B, T, dim) features are read in place through a 4-D tensor map in
// 64-feature chunks (columns beyond dim are zero-filled by TMA), four chunks ("taps") per CTA group share the g0 tile.
int proj_wgrad_tc_ctas(int dim, int tiles, int* groups_out) {
  const int nchunks = (dim + 63) / 64, groups = (nchunks + 3) / 4;
  int cpl = sm_count() / groups;
  if (cpl < 1) cpl = 1;
  if (cpl > tiles) cpl = tiles > 0 ? tiles : 1;
  if (groups_out) *groups_out = groups;
  return cpl;
}
int64_t proj_wgrad_tc_scratch(int dim) {
  int groups;
  const int cpl = proj_wgrad_tc_ctas(dim, 1 << 30, &groups);
  return (int64_t)groups * cpl * tc::kWgPartFloats;
}
int do_proj_wgrad_reduce(int dim, int tiles, float* gw, float* gb, const float* part, int accumulate, cudaStream_t st) {
  if (dbg_skip("projbwd")) return 0;
  tc::ProjWgradReduceArgs ra;
  ra.part = part; ra.gw = gw; ra.gb = gb; ra.dim = dim; ra.P = proj_wgrad_tc_ctas(dim, tiles, nullptr); ra.accumulate = accumulate;
  return launch_pdl_grid("proj_wgrad_reduce_kernel", tc::proj_wgrad_reduce_kernel, dim3((dim * 64 + 64 + 31) / 32), dim3(256), st, ra);
}
// reduce_now = false: the caller issues do_proj_wgrad_reduce later on the same stream (other work in between)
int do_proj_wgrad_tc(const float* x, const float* g0, const int* lens, int B, int T, int dim, float* gw, float* gb, float* part,
                     int accumulate, cudaStream_t st, bool reduce_now = true) {
  if (dbg_skip("projbwd")) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail("proj_wgrad_tc: features must be 16-byte aligned");
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is not available from this driver");
  struct Entry { const float* x; int B, T, dim; CUtensorMap tm; };
  thread_local Entry cache = {};
  if (cache.x != x || cache.B != B || cache.T != T || cache.dim != dim) {
    cuuint64_t dims[4] = {(cuuint64_t)dim, (cuuint64_t)T, (cuuint64_t)B, 1};
    cuuint64_t strides[3] = {(cuuint64_t)dim * 4, (cuuint64_t)T * dim * 4, (cuuint64_t)B * T * dim * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)tc::TW, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(&cache.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, l2_promotion(),
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      cache.x = nullptr;
      char buf[64];
      snprintf(buf, sizeof buf, "%d", (int)r);
      return fail("proj_wgrad_tc: cuTensorMapEncodeTiled failed with CUresult %s", buf);
    }
    cache.x = x; cache.B = B; cache.T = T; cache.dim = dim;
  }
  CUtensorMap tg;
  if (make_act_tensor_map(&tg, g0, B, T, 1, 1, 0, tc::TW)) return 1;
  tc::TcWgradArgs a = {};
  a.lens = lens; a.part = part; a.B = B; a.T = T;
  a.tiles_per_video = (T + tc::TW - 1) / tc::TW; a.num_tiles = a.tiles_per_video * B;
  a.proj = 1; a.nchunks = (dim + 63) / 64;
  int groups;
  const int cpl = proj_wgrad_tc_ctas(dim, a.num_tiles, &groups);
  a.nlayers = groups; a.ctas_per_layer = cpl; a.tap_mask = 0xF;
  if (set_smem(tc::tc_wgrad_kernel, tc::kTcWgradSmem)) return 1;
  if (launch_pdl("tc_wgrad_kernel(proj)", tc::tc_wgrad_kernel, groups * cpl, tc::kTcWgradSmem, st, cache.tm, tg, tg, tg, tg, a)) return 1;
  return reduce_now ? do_proj_wgrad_reduce(dim, a.num_tiles, gw, gb, part, accumulate, st) : 0;
}

// stage tail backward on the tensor cores (tc_layer_kernel<4>): gin (NULL for the last stage), q_s, gr_s -> gz, ga.
// Optional fusion (gu_out != NULL): the top layer's pre-activation gradient gu(L-1) = (W1(L-1)^T (ga*mask*dropout)) * [h(L-1) > 0]
// -- tc_bwd_gu_kernel's work, tile-local like the tail itself -- as a third GEMM + epilogue of the same launch: one
// dependent layer step and one kernel ramp less per stage.  h_top = h(s, L-1), wimg_top = layer L-1's backward image.
int do_tail_bwd_tc(const float* gin, const float* q, const float* gr, float* gz, float* ga, const int* lens, int B,
                   int T, int K, const float* timg_b, cudaStream_t st, const int* flags_in = nullptr, int* flags_out = nullptr,
                   float* gu_out = nullptr, const float* h_top = nullptr, const float* wimg_top = nullptr,
                   const mstcn_dropout* drop = nullptr, int layer_id_top = 0) {
  CUtensorMap tm, tg, thp;
  if (make_act_tensor_map(&tm, gin ? gin : gr, B, T, 0, 1) || make_act_tensor_map(&tg, gin ? q : gr, B, T, 0, 1) ||
      make_act_tensor_map(&thp, gr, B, T, 0, 1))
    return 1;
  tc::TcLayerFwdArgs a = {};
  a.nsteps = 1;
  a.lens = lens; a.wimg = timg_b; a.y = ga; a.h = gz;
  a.B = B; a.T = T; a.d = -(T + 2 * tc::TM); a.skip_extra = 0;
  a.tiles_per_video = (T + tc::TM - 1) / tc::TM; a.num_tiles = a.tiles_per_video * B;
  a.gyp = gin; a.K = K;
  a.flags_in = gin ? flags_in : nullptr; a.flags = flags_out;
  if (gu_out != nullptr) {
    a.gu_out = gu_out; a.hprev = h_top; a.wimg2 = wimg_top;
    a.train = drop && drop->enabled; a.layer_id = (uint32_t)layer_id_top;
    a.seed = drop ? drop->seed : 0; a.offset = drop ? drop->offset : 0;
    a.offset_dev = drop ? reinterpret_cast<const unsigned long long*>(drop->offset_dev) : nullptr;
  }
  if (set_smem(tc::tc_layer_kernel<4>, tc::kTcFwdSmem)) return 1;
  return launch_pdl_threads("tc_layer_kernel<4>", tc::tc_layer_kernel<4>, persistent_grid(a.num_tiles, 1), tc::kTcLayerThreads, tc::kTcFwdSmem, st,
                            tm, tg, thp, a);
}

// stage tail forward on the tensor cores (tc_layer_kernel<3>): a -> logits (B*T,K), q (optional), next_x0 (NULL for the
// last stage).  timg = the stage's forward tail image; bout zero-padded to 64; bn = next stage's conv_1x1 bias.
// stage-1 input projection on the tensor cores: x (n, dim) read in place through a 2-D tensor map
int do_proj_fwd_tc(const float* x, int64_t n, int dim, const float* wimg, const float* bias, const int* lens, int T, float* y,
                   cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail("proj_fwd_tc: features must be 16-byte aligned");
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is not available from this driver");
  struct Entry { const float* x; int64_t n; int dim; CUtensorMap tm; };
  thread_local Entry cache[8] = {};
  Entry& e = cache[(reinterpret_cast<uintptr_t>(x) >> 8) & 7];
  if (e.x != x || e.n != n || e.dim != dim) {
    cuuint64_t dims[2] = {(cuuint64_t)dim, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)dim * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)tc::TM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&e.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(),
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { e.x = nullptr; return fail("cuTensorMapEncodeTiled (features) failed"); }
    e.x = x; e.n = n; e.dim = dim;
  }
  tc::TcProjArgs a;
  a.wimg = wimg; a.bias = bias; a.lens = lens; a.y = y; a.n_rows = n; a.T = T;
  a.kblocks = (dim + 31) / 32; a.num_tiles = (int)((n + tc::TM - 1) / tc::TM);
  if (a.num_tiles == 0) return 0;
  if (set_smem(tc::tc_proj_kernel, tc::kTcProjSmem)) return 1;
  // ordinary launch: the kernel has no griddepcontrol.wait, it must not start before the work ahead of it is done
  tc::tc_proj_kernel<<<persistent_grid(a.num_tiles, 1), tc::kTcThreads, tc::kTcProjSmem, st>>>(e.tm, a);
  return check_launch("tc_proj_kernel");
}

// flags_in: the chain's last-step flags of the tiles of a_in (NULL: ordinary grid dependency); flags_out: this
// launch's own per-tile flags for the next stage's chain (NULL: none)
int do_tail_fwd_tc(const float* a_in, const int* lens, int B, int T, int K, const float* timg, const float* bout,
                   const float* bn, float* logits, float* q_out, float* next_x0, cudaStream_t st,
                   const int* flags_in = nullptr, int* flags_out = nullptr) {
  TcChain ch;
  ch.flags_in = flags_in; ch.flags = next_x0 != nullptr ? flags_out : nullptr;
  return launch_tc_layer<3>(a_in, nullptr, next_x0, q_out, lens, B, T, T + 2 * tc::TM, timg, bout, bn ? bn : bout, nullptr, 0, st,
                            0, nullptr, nullptr, logits, K, ch);
}

// layer l's input gradient fused with layer l-1's pre-activation gradient (tc_layer_kernel<2>):
// gu_l, gy_l -> gx_l ; gx_l, h_{l-1} -> gu_{l-1}.  drop / layer_id are layer l-1's.

// gx = gy*mask + sum_k Wd[:,:,k]^T gu[t-(k-1)d] on the tensor cores (wimg_b = the layer's backward image)
int do_layer_bwd_gx_tc(const float* gu, const float* gy, float* gx, const int* lens, int B, int T, int d,
                       const float* wimg_b, cudaStream_t st, const int* flags_in, int* flags_out) {
  TcChain ch;
  ch.flags_in = flags_in; ch.flags = flags_out;
  return launch_tc_layer<1>(gu, gy, gx, nullptr, lens, B, T, d, wimg_b, nullptr, nullptr, nullptr, 0, st, 0, nullptr, nullptr,
                            nullptr, 0, ch);
}


// ---- video groups on concurrent streams --------------------------------------------------------
// A batch of a few thousand frames gives each layer kernel ~1.1 waves of 128-frame tiles, so most SMs
// idle while a few run a second tile -- and the layer chain is strictly sequential.  Every op of the
// model is per-video, so the batch is cut into contiguous video groups whose kernel chains run on
// separate streams: the tiles of one group's layer l fill the SMs another group's layer l' leaves idle.
constexpr int kMaxGroups = 4;

struct StreamPool {
  cudaStream_t side[2 * kMaxGroups];
  cudaEvent_t ev[512];
  cudaEvent_t ev_stage[16];      // weight gradients of stage s reduced (recorded on the wgrad stream)
  int next_ev = 0;
  bool ready = false;
  int init() {
    if (ready) return 0;
    // side[0 .. kMaxGroups) carry layer chains of video groups, side[kMaxGroups ..) the weight-gradient kernels.
    // (Stream priorities were measured and made things slower: 2.66 vs 2.40 ms per step.)
    for (int i = 0; i < 2 * kMaxGroups; ++i)
      if (cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
    for (auto& e : ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate failed");
    for (auto& e : ev_stage)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate failed");
    ready = true;
    return 0;
  }
  cudaEvent_t event() { cudaEvent_t e = ev[next_ev]; next_ev = (next_ev + 1) % 512; return e; }
};
// one pool per (thread, device): streams and events belong to the device that was current when they were created
StreamPool& pool() {
  static thread_local StreamPool p[kMaxDevices];
  const int dev = current_device();
  return p[dev < 0 ? 0 : dev];
}

// ---- chain lane ------------------------------------------------------------------------------
// The chain launches (and the flag-linked kernels around them) spin on per-tile flags written by other CTAs, so every
// CTA of such a launch must become resident while its producers run.  Two of these launch sequences on DIFFERENT
// streams of one GPU (two models, eval under train, a concurrent ensemble) can each hold SMs the other needs and
// starve until the bounded wait traps.  The lane serialises them: a sequence first makes its stream wait for the
// previous sequence's end (an event recorded on whatever stream that ran on), and records its own end when done.
// On the same stream this is a no-op.  Stream capture is left alone (a captured step is one stream-ordered unit;
// graph.GraphedTrainStep orders its own replays the same way on the Python side).
struct ChainLane {
  std::mutex mu;
  cudaEvent_t ev[kMaxDevices] = {};
  cudaStream_t last[kMaxDevices] = {};
  bool used[kMaxDevices] = {};
};
ChainLane& lane() { static ChainLane l; return l; }

bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}

int lane_enter(cudaStream_t st) {
  if (stream_capturing(st)) return 0;
  const int dev = current_device();
  if (dev < 0) return fail("no current CUDA device");
  ChainLane& l = lane();
  std::lock_guard<std::mutex> lock(l.mu);
  if (l.used[dev] && l.last[dev] != st && cudaStreamWaitEvent(st, l.ev[dev], 0) != cudaSuccess)
    return fail("chain lane: cudaStreamWaitEvent failed");
  return 0;
}

int lane_exit(cudaStream_t st) {
  if (stream_capturing(st)) return 0;
  const int dev = current_device();
  if (dev < 0) return fail("no current CUDA device");
  ChainLane& l = lane();
  std::lock_guard<std::mutex> lock(l.mu);
  if (!l.ev[dev] && cudaEventCreateWithFlags(&l.ev[dev], cudaEventDisableTiming) != cudaSuccess)
    return fail("chain lane: cudaEventCreate failed");
  if (cudaEventRecord(l.ev[dev], st) != cudaSuccess) return fail("chain lane: cudaEventRecord failed");
  l.last[dev] = st; l.used[dev] = true;
  return 0;
}

// contiguous video ranges [gb[g], gb[g+1]) with ~equal numbers of valid 128-frame tiles
int plan_groups(const int32_t* lens_host, int B, int want, int* gb) {
  int G = want < 1 ? 1 : want > kMaxGroups ? kMaxGroups : want;
  if (!lens_host || G > B) G = lens_host ? (B < G ? B : G) : 1;
  if (G <= 1) { gb[0] = 0; gb[1] = B; return 1; }
  long long total = 0;
  for (int b = 0; b < B; ++b) total += (lens_host[b] + tc::TM - 1) / tc::TM;
  int g = 0; long long acc = 0;
  gb[0] = 0;
  for (int b = 0; b < B; ++b) {
    acc += (lens_host[b] + tc::TM - 1) / tc::TM;
    const int remaining_videos = B - (b + 1), remaining_groups = G - (g + 1);
    if (g < G - 1 && (acc * G >= total * (g + 1) || remaining_videos == remaining_groups)) gb[++g] = b + 1;
  }
  gb[G] = B;
  return G;
}

struct Fork {
  cudaStream_t main; cudaStream_t st[kMaxGroups]; int G;
  // own_streams: every group (also a single one) runs on a high-priority internal stream forked off the
  // caller's; otherwise group 0 stays on the caller's stream
  int first = 1;
  int begin(cudaStream_t m, int groups, bool own_streams = false) {
    main = m; G = groups; st[0] = m;
    first = own_streams ? 0 : 1;
    if (G - first <= 0) return 0;
    if (pool().init()) return 1;
    cudaEvent_t e = pool().event();
    if (cudaEventRecord(e, m) != cudaSuccess) return fail("cudaEventRecord failed");
    for (int g = first; g < G; ++g) {
      st[g] = pool().side[g];
      if (cudaStreamWaitEvent(st[g], e, 0) != cudaSuccess) return fail("cudaStreamWaitEvent failed");
    }
    return 0;
  }
  int join() {
    for (int g = first; g < G; ++g) {
      cudaEvent_t e = pool().event();
      if (cudaEventRecord(e, st[g]) != cudaSuccess || cudaStreamWaitEvent(main, e, 0) != cudaSuccess)
        return fail("stream join failed");
    }
    return 0;
  }
};

}  // namespace

// =============================================================================================
extern "C" {

int mstcn_abi_version(void) { return MSTCN_ABI_VERSION; }
const char* mstcn_last_error(void) { return g_err.c_str(); }
int mstcn_sm_count(void) { return sm_count(); }

int64_t mstcn_param_count(const mstcn_dims* d) { return check_dims(d) ? -1 : make_layout(d).total(); }
int64_t mstcn_packed_count(const mstcn_dims* d) { return check_dims(d) ? -1 : make_layout(d).ptotal_with_tc(); }
int32_t mstcn_param_tensors(const mstcn_dims* d) { return check_dims(d) ? -1 : make_layout(d).tensors(); }

int64_t mstcn_param_offset(const mstcn_dims* d, int32_t index) {
  if (check_dims(d)) return -1;
  Layout lay = make_layout(d);
  if (index < 0 || index >= lay.tensors()) return -1;
  const int per_stage = 4 + 4 * lay.L;
  const int s = index / per_stage, r = index % per_stage;
  if (r == 0) return lay.win_w(s);
  if (r == 1) return lay.win_b(s);
  if (r >= 2 + 4 * lay.L) return r == 2 + 4 * lay.L ? lay.wout(s) : lay.bout(s);
  const int l = (r - 2) / 4, w = (r - 2) % 4;
  return w == 0 ? lay.wd(s, l) : w == 1 ? lay.bd(s, l) : w == 2 ? lay.w1(s, l) : lay.b1(s, l);
}

int64_t mstcn_packed_offset(const mstcn_dims* d, int32_t stage, int32_t layer, int32_t which) {
  if (check_dims(d)) return -1;
  Layout lay = make_layout(d);
  if (stage < 0 || stage >= lay.S) return -1;
  if (which >= 3 && which <= 8 && (layer < 0 || layer >= lay.L)) return -1;
  switch (which) {
    case 0: return lay.p_win_t(stage);
    case 1: return lay.p_bin(stage);
    case 2: return lay.p_win_b(stage);
    case 3: return lay.p_wd_t(stage, layer);
    case 4: return lay.p_bd(stage, layer);
    case 5: return lay.p_w1_t(stage, layer);
    case 6: return lay.p_b1(stage, layer);
    case 7: return lay.p_wd_b(stage, layer);
    case 8: return lay.p_w1_n(stage, layer);
    case 9: return lay.p_wout_t(stage);
    case 10: return lay.p_bout(stage);
    case 11: return lay.p_wout_b(stage);
    case 12: return (layer < 0 || layer >= lay.L) ? -1 : lay.p_tc(stage, layer);
    case 13: return (layer < 0 || layer >= lay.L) ? -1 : lay.p_tcb(stage, layer);
    case 14: return lay.p_tp();
    default: return -1;
  }
}

int64_t mstcn_bucket_boundary(const mstcn_dims* d, int32_t stage) {
  if (check_dims(d)) return -1;
  Layout lay = make_layout(d);
  if (stage < 0 || stage > lay.S) return -1;
  return stage == lay.S ? lay.total() : lay.layer(stage, 0);
}

int mstcn_pack_params(const mstcn_dims* d, const float* params, float* packed, void* stream) {
  if (check_dims(d)) return 1;
  if (!params || !packed) return fail("pack_params: NULL pointer");
  Layout lay = make_layout(d);
  if (use_tc_bwd(d) && (d->flags & MSTCN_FLAG_PACK_TC_ONLY)) {
    const int n = lay.S * 64 * (2 + 2 * lay.L);
    pack_biases_kernel<<<(n + 255) / 256, 256, 0, S(stream)>>>(lay, params, packed);
    if (check_launch("pack_biases_kernel")) return 1;
  } else {
    dim3 grid(64, lay.S);
    pack_params_kernel<<<grid, 256, 0, S(stream)>>>(lay, params, packed);
    if (check_launch("pack_params_kernel")) return 1;
  }
  if (use_tc(d)) {
    const long long items = ((long long)lay.S * lay.L * 16 + lay.S * 8 + lay.proj_kblocks()) * 2048;
    tc::tc_pack_all_kernel<<<(unsigned)((items + 255) / 256), 256, 0, S(stream)>>>(lay, params, packed);
    return check_launch("tc_pack_all_kernel");
  }
  return 0;
}

int64_t mstcn_workspace_floats(const mstcn_dims* d, int32_t B, int32_t T, int32_t training) {
  if (check_dims(d)) return -1;
  if (B < 1 || T < 1) { fail("B and T must be >= 1"); return -1; }
  Ws w = carve(d, B, T, training != 0, nullptr);
  return (w.flag_base - static_cast<float*>(nullptr)) + w.flag_count();
}

int64_t mstcn_workspace_offset(const mstcn_dims* d, int32_t B, int32_t T, int32_t training, int32_t what, int32_t stage,
                               int32_t layer) {
  if (check_dims(d)) return -1;
  if (B < 1 || T < 1 || stage < 0 || stage >= d->num_stages) return -1;
  Ws w = carve(d, B, T, training != 0, nullptr);
  const float* p = nullptr;
  switch (what) {
    case 0: if (layer < 0 || layer > w.L) return -1; p = w.act(stage, layer); break;
    case 1: if (layer < 0 || layer >= w.L || !training) return -1; p = w.h(stage, layer); break;
    case 2: p = w.logits(stage); break;
    case 3: if (!training) return -1; p = w.q(stage); break;
    default: return -1;
  }
  return p - static_cast<const float*>(nullptr);
}

// one stage's L layers as a chain launch; planes = L+1 contiguous activation planes, hplanes = L planes or NULL
int do_stage_fwd_tc(const Layout& lay, const float* packed, int s, float* planes, float* hplanes, const int* lens, int B, int T,
                    const mstcn_dropout* drop, int* flags, cudaStream_t st, const int* flags_in = nullptr, int publish_last = 0) {
  const int L = lay.L;
  TcChain ch;
  ch.nsteps = L; ch.lyr0 = 0; ch.dir = 1; ch.nx = L + 1; ch.plane = (int64_t)B * T * 64;
  ch.wimg_stride = Layout::kTcLayerImage; ch.bias_stride = L > 1 ? lay.p_bd(s, 1) - lay.p_bd(s, 0) : 0;
  ch.flags = flags; ch.flags_in = flags_in; ch.publish_last = publish_last;
  return launch_tc_layer<0>(planes, nullptr, planes + ch.plane, hplanes, lens, B, T, 1, packed + lay.p_tc(s, 0),
                            packed + lay.p_bd(s, 0), packed + lay.p_b1(s, 0), drop, s * L, st, 0, nullptr, nullptr, nullptr, 0, ch);
}

int mstcn_forward(const mstcn_dims* d, const float* packed, const float* x, const int32_t* lens,
                  const int32_t* lens_host, int32_t groups, int32_t B, int32_t T, const mstcn_dropout* drop,
                  int32_t training, float* workspace, float* out, uint8_t* winner, void* stream) {
  if (check_dims(d)) return 1;
  if (!packed || !x || !lens || !workspace) return fail("forward: NULL pointer");
  if ((!out || !winner) && !(use_tc(d) && training)) return fail("forward: out / winner may only be NULL on the tensor-core training path (mstcn_loss_head takes the max)");
  if (B < 1 || T < 1) return fail("forward: B and T must be >= 1");
  if ((int64_t)B * T >= (1LL << 31) / 64) return fail("forward: B*T too large for 32-bit tile indexing");
  Layout lay = make_layout(d);
  Ws w = carve(d, B, T, training != 0, workspace);
  const int L = lay.L, K = lay.K;
  if (use_tc(d)) {
    // Tensor-core path, all on the caller's stream: projection, then per stage ONE chain launch over its L layers
    // (tile-level dataflow between layers, no kernel boundary) and the stage tail.  Chain launches spin on tiles
    // of their own grid, so two of them must never share the GPU: no video groups here.
    (void)lens_host; (void)groups;
    cudaStream_t st = S(stream);
    if (lane_enter(st)) return 1;
    if (cudaMemsetAsync(w.flags(0, 0), 0, sizeof(int) * lay.S * (L + 2) * w.num_tiles, st) != cudaSuccess)
      return fail("forward: clearing the tile flags failed");
    // Training: every kernel writes planes of its own, so consecutive kernels are chained by the per-tile flags alone
    // (no grid dependency): the tail starts on tiles the chain's last layer has published, the next stage's chain on
    // tiles the tail has published.  Inference shares the layer planes between stages and keeps the grid dependency.
    static const int df_fwd = getenv("MSTCN_DF_FWD") ? atoi(getenv("MSTCN_DF_FWD")) : 1;   // diagnosis: 0 = forward links off
    const bool df = training != 0 && pdl_enabled() && df_enabled() && df_fwd != 0;
    if (do_proj_fwd_tc(x, w.N, lay.dim, packed + lay.p_tp(), packed + lay.p_bin(0), lens, T, w.act(0, 0), st)) return 1;
    for (int s = 0; s < lay.S; ++s) {
      int* const fl = w.flags(0, s);                       // rows 0..L-1: the chain's steps, row L: the tail
      const int* const fl_prev_tail = (df && s > 0) ? w.flags(0, s - 1) + (int64_t)L * w.num_tiles : nullptr;
      if (do_stage_fwd_tc(lay, packed, s, w.act(s, 0), w.h(s, 0), lens, B, T, drop, fl, st, fl_prev_tail, df ? 1 : 0)) return 1;
      const bool last = s == lay.S - 1;
      if (do_tail_fwd_tc(w.act(s, L), lens, B, T, K, packed + lay.p_tt(s), packed + lay.p_bout(s),
                         last ? nullptr : packed + lay.p_bin(s + 1), w.logits(s), (w.q(s) && !last) ? w.q(s) : nullptr,
                         last ? nullptr : w.act(s + 1, 0), st, df ? fl + (int64_t)(L - 1) * w.num_tiles : nullptr,
                         df ? fl + (int64_t)L * w.num_tiles : nullptr))
        return 1;
    }
    if (!out || !winner) return lane_exit(st);          // the caller takes the max inside mstcn_loss_head
    const int64_t n = w.N * K;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8 * 148) blocks = 8 * 148;
    // max over stages + winner from the per-stage logits (torch.cat / permute / torch.max, :312-319)
    tc::stage_max_kernel<<<blocks, 256, 0, st>>>(w.logits(0), training ? w.act_stage : w.lg(), lay.S, n, out, winner);
    if (check_launch("stage_max_kernel")) return 1;
    return lane_exit(st);
  }
  int gb[kMaxGroups + 1];
  const int G = plan_groups(lens_host, B, groups, gb);
  Fork fk;
  if (fk.begin(S(stream), G)) return 1;
  for (int g = 0; g < G; ++g) {                 // one independent kernel chain per video group
    cudaStream_t st = fk.st[g];
    const int b0 = gb[g], Bg = gb[g + 1] - gb[g];
    const size_t f0 = (size_t)b0 * T;           // first frame of the group
    const int* gl = lens + b0;
    if (do_proj_fwd(x + f0 * lay.dim, (int64_t)Bg * T, lay.dim, packed + lay.p_win_t(0), packed + lay.p_bin(0),
                    w.act(0, 0) + f0 * 64, st))
      return 1;
    for (int s = 0; s < lay.S; ++s) {
      for (int l = 0; l < L; ++l) {
        const float* xin = w.act(s, l) + f0 * 64;
        float* yout = w.act(s, l + 1) + f0 * 64;
        float* hout = w.h(s, l) ? w.h(s, l) + f0 * 64 : nullptr;
        if (do_layer_fwd(xin, yout, hout, gl, Bg, T, 1 << l, packed + lay.p_wd_t(s, l), packed + lay.p_bd(s, l),
                         packed + lay.p_w1_t(s, l), packed + lay.p_b1(s, l), drop, s * L + l, st, (uint32_t)f0))
          return 1;
      }
      const bool last = s == lay.S - 1;
      float* next_x0 = last ? nullptr : w.act(s + 1, 0) + f0 * 64;
      if (do_tail_fwd(w.act(s, L) + f0 * 64, gl, Bg, T, K, s, packed + lay.p_wout_t(s), packed + lay.p_bout(s),
                             w.logits(s) + f0 * K, out + f0 * K, winner + f0 * K,
                             last ? nullptr : packed + lay.p_win_t(s + 1), last ? nullptr : packed + lay.p_bin(s + 1),
                      next_x0, st))
        return 1;
    }
  }
  return fk.join();
}

int mstcn_backward_stage(const mstcn_dims* d, const float* packed, const float* x, const int32_t* lens,
                         const int32_t* lens_host, int32_t groups, int32_t B, int32_t T, const mstcn_dropout* drop,
                         float* workspace, const uint8_t* winner, const float* gout, const float* gscale, float* grads,
                         int32_t accumulate, int32_t stage, void* stream) {
  if (check_dims(d)) return 1;
  if (!packed || !x || !lens || !workspace || !grads) return fail("backward: NULL pointer");   // winner may be NULL
  if (!gout && !use_tc_bwd(d)) return fail("backward: gout may only be NULL (gradient planes written by mstcn_loss_head) on the tensor-core path");
  Layout lay = make_layout(d);
  if (stage < 0 || stage >= lay.S) return fail("backward_stage: stage out of range");
  Ws w = carve(d, B, T, true, workspace);
  cudaStream_t main = S(stream);
  const int L = lay.L, K = lay.K;
  const bool tcb = use_tc_bwd(d);
  if (lay.S > 16) return fail("backward: more than 16 stages");
  (void)lens_host; (void)groups;                  // the backward runs one chain + one weight-gradient stream
  const int s = stage;
  const bool last = s == lay.S - 1;
  const int p = s & 1;                            // plane set of this stage; stage s+1 used the other one
  // scratch: [tail partials | per-layer wgrad partials | proj partials]
  float* sc_tail = w.scratch();
  float* sc_layer = sc_tail + scratch_tail_region();
  float* sc_proj = sc_layer + scratch_layer_region(d);
  const int64_t plane = w.N * 64;
  const float* gin = last ? nullptr : w.gl(1 - p, 0);
  if (tcb && pool().init()) return 1;
  cudaStream_t wst = tcb ? pool().side[kMaxGroups] : main;
  if (tcb && lane_enter(main)) return 1;

  // Kernel-to-kernel dataflow: consecutive kernels of the chain are linked by per-tile flags instead of grid
  // dependencies (rows of this stage's flag block: tail | top-layer gu | chain steps | layer-0 gx).  Plane-set reuse
  // across stages stays protected by the ev_stage events below.
  const bool df = tcb && pdl_enabled() && df_enabled();
  const int64_t nt = w.num_tiles;
  int* const r_tail = w.flags(1, s);
  int* const r_gu = r_tail + nt;
  int* const r_chain = r_gu + nt;
  int* const r_m1 = r_tail + (int64_t)(L + 1) * nt;
  // diagnosis: MSTCN_DF_OFF = bit mask of consumer-side links that fall back to the grid dependency (griddepcontrol.wait):
  // 1 = tail <- previous stage's layer-0 gx, 2 = top-layer gu <- tail, 4 = chain <- gu, 8 = layer-0 gx <- chain
  static const int df_off = getenv("MSTCN_DF_OFF") ? atoi(getenv("MSTCN_DF_OFF")) : 0;
  // the top layer's gu rides in the tail launch (MSTCN_FUSE_GU=0: tc_bwd_gu_kernel as a launch of its own, as in round 1)
  static const bool fuse_gu_on = !(getenv("MSTCN_FUSE_GU") != nullptr && getenv("MSTCN_FUSE_GU")[0] == '0');
  const bool fuse_gu = tcb && fuse_gu_on;
  int* const r_top = fuse_gu ? r_tail : r_gu;     // the flag row that says "gu(L-1) and Gl[L] of this tile are stored"
  // ---- the critical-path chain on the caller's stream ----
  int tail_p = 0;
  if (tcb) {
    // dL/dout routed to the winning stage of every (frame, class), once per backward: S zero-padded (N, 64) planes
    if (last) {
      if (cudaMemsetAsync(w.flags(1, 0), 0, sizeof(int) * lay.S * (L + 2) * w.num_tiles, main) != cudaSuccess)
        return fail("backward: clearing the tile flags failed");
      if (gout != nullptr) {
        int blocks = (int)((w.N * 16 + 255) / 256);
        if (blocks > 8 * 148) blocks = 8 * 148;
        tc::route_grad_kernel<<<blocks, 256, 0, main>>>(gout, gscale, winner, lay.S, K, w.N, w.gr(0), plane);
        if (check_launch("route_grad_kernel")) return 1;
      }
    }
    if (do_tail_bwd_tc(gin, w.q(s), w.gr(s), w.gz(p), w.gl(p, L), lens, B, T, K, packed + lay.p_ttb(s), main,
                       (df && !last && !(df_off & 1)) ? w.flags(1, s + 1) + (int64_t)(L + 1) * nt : nullptr, df ? r_tail : nullptr,
                       fuse_gu ? w.gu(p, L - 1) : nullptr, w.h(s, L - 1), packed + lay.p_tcb(s, L - 1), drop, s * L + L - 1))
      return 1;
  } else if (do_tail_bwd(w.act(s, L), w.logits(s), winner ? gout : gout + (int64_t)s * w.N * K, gscale, winner, gin, lens, B, T, K, s,
                         packed + lay.p_wout_b(s),
                         last ? nullptr : packed + lay.p_win_b(s + 1), w.gl(p, L), grads + lay.wout(s), grads + lay.bout(s),
                         last ? nullptr : grads + lay.win_w(s + 1), last ? nullptr : grads + lay.win_b(s + 1), sc_tail,
                         accumulate, main, &tail_p)) {
    return 1;
  }
  if (tcb) {
    // top layer: its pre-activation gradient comes from the tail's ga; every other gu(l-1) is produced by the
    // fused kernel of layer l together with gx(l)
    if (!fuse_gu && do_bwd_gu_tc(w.gl(p, L), w.h(s, L - 1), w.gu(p, L - 1), lens, B, T, packed + lay.p_tcb(s, L - 1), drop,
                                 s * L + L - 1, main, 0, (df && !(df_off & 2)) ? r_tail : nullptr, df ? r_gu : nullptr))
      return 1;
    if (L > 1) {
      // layers L-1 .. 1 as ONE chain launch: step j = layer L-1-j reads gu(l) (tm_x plane l), gy = Gl[l+1], h(l-1) and
      // writes gx = Gl[l] and gu(l-1); tiles of step j start as soon as step j-1's tiles under their taps are done
      TcChain ch;
      ch.nsteps = L - 1; ch.lyr0 = L - 1; ch.dir = -1; ch.nx = L; ch.ng = L + 1; ch.nhp = L;
      ch.cg_off = 1; ch.chp_off = -1; ch.plane = plane; ch.wimg_stride = Layout::kTcLayerImage;
      ch.flags = r_chain; ch.flags_in = (df && !(df_off & 4)) ? r_top : nullptr; ch.publish_last = df ? 1 : 0;
      StageTimer tm(main, wst);
      if (launch_tc_layer<2>(w.gu(p, 0), w.gl(p, 0), w.gu(p, 0) - plane, w.gl(p, 0), lens, B, T, 1, packed + lay.p_tcb(s, 0),
                             nullptr, nullptr, drop, s * L - 1, main, 0, w.h(s, 0), packed + lay.p_tcb(s, 0) - Layout::kTcLayerImage,
                             nullptr, 0, ch))
        return 1;
      tm.stop(&g_bwd_times[2 * s]);
    }
    if (do_layer_bwd_gx_tc(w.gu(p, 0), w.gl(p, 1), w.gl(p, 0), lens, B, T, 1, packed + lay.p_tcb(s, 0), main,
                           (df && !(df_off & 8)) ? (L > 1 ? r_chain + (int64_t)(L - 2) * nt : r_top) : nullptr, (df && s > 0) ? r_m1 : nullptr))
      return 1;
  } else {
    for (int l = L - 1; l >= 0; --l)
      if (do_layer_bwd(w.act(s, l), w.h(s, l), w.gl(p, l + 1), w.gl(p, l), w.gu(p, 0), lens, B, T, 1 << l,
                       packed + lay.p_wd_b(s, l), packed + lay.p_w1_n(s, l), drop, s * L + l, grads + lay.wd(s, l),
                       grads + lay.bd(s, l), grads + lay.w1(s, l), grads + lay.b1(s, l), sc_layer, accumulate, main))
        return 1;
  }
  // FFMA tail: partials -> conv_out(s) and the next stage's input projection
  if (!tcb) {
    ReduceArgs ra; ra.accumulate = accumulate; ra.nseg = 2;
    ra.seg[0] = seg(sc_tail, grads + lay.wout(s), kTailBwdPart, tail_p, K, 64, 64);
    ra.seg[1] = seg(sc_tail + 4096, grads + lay.bout(s), kTailBwdPart, tail_p, 1, 64, K);
    if (!last) {
      ra.seg[2] = seg(sc_tail + 4160, grads + lay.win_w(s + 1), kTailBwdPart, tail_p, 64, 64, K);
      ra.seg[3] = seg(sc_tail + 4160 + 4096, grads + lay.win_b(s + 1), kTailBwdPart, tail_p, 1, 64, 64);
      ra.nseg = 4;
    }
    if (launch_reduce(ra, main)) return 1;
  }
  // stage-1 input projection: on the tensor-core path its weight gradient rides the weight-gradient stream below
  // (MSTCN_PROJ_WGRAD_FFMA=1: the fp32 FFMA kernel on the caller's stream, as in round 1)
  static const bool proj_ffma = getenv("MSTCN_PROJ_WGRAD_FFMA") != nullptr && getenv("MSTCN_PROJ_WGRAD_FFMA")[0] == '1';
  if (s == 0 && (!tcb || proj_ffma) &&
      do_proj_bwd(x, w.gl(p, 0), w.N, lay.dim, grads + lay.win_w(0), grads + lay.win_b(0), sc_proj, accumulate, main))
    return 1;

  if (tcb) {
    // ---- all weight gradients of the stage: ONE kernel on the side stream, overlapping the next stage's chain.
    //      Grid = L x R CTAs; each layer's R CTAs keep that layer's four accumulators in TMEM over ~tiles/R tiles. ----
    int R = sm_count() / L;
    if (R < 1) R = 1;
    cudaEvent_t e = pool().event();
    if (cudaEventRecord(e, main) != cudaSuccess || cudaStreamWaitEvent(wst, e, 0) != cudaSuccess)
      return fail("event record / wait failed");
    // ... plus kTailWgradCtas CTAs for the 1x1 convolutions around the stage: dWout(s) = gz^T a(s,L) (tap 0) and, for
    // s > 0, this stage's input projection dWn(s) = g0^T q(s-1), dbn(s) = sum over ALL frames of g0 (tap 3; unmasked conv)
    const int Rt = kTailWgradCtas;
    R = (sm_count() - Rt) / L;
    if (R < 1) R = 1;
    float* sc_tail_w = sc_layer + (int64_t)L * R * tc::kWgPartFloats;
    StageTimer tmw(wst, main);
    if (do_wgrad_tc_multi(w.gu(p, 0), plane, w.gl(p, 1), plane, w.act(s, 0), plane, w.h(s, 0), plane, lens, B, T, 1, L, R, drop,
                          s * L, sc_layer, wst, 0, Rt, s > 0 ? 0x9 : 0x1, s > 0 ? w.q(s - 1) : nullptr))
      return 1;
    tmw.stop(&g_bwd_times[2 * s + 1]);
    // stage 0 ends the backward: dW / db of the stage-1 input projection from the features and Gl[0] (tc_wgrad_kernel in
    // projection mode).  It needs the whole SM like the launch above, so it queues directly behind it (its prologue
    // overlaps that launch's drain); the small reductions of both follow.
    const bool proj_tc = s == 0 && !proj_ffma;
    if (proj_tc && do_proj_wgrad_tc(x, w.gl(p, 0), lens, B, T, lay.dim, grads + lay.win_w(0), grads + lay.win_b(0), sc_proj,
                                    accumulate, wst, false))
      return 1;
    {
      ReduceArgs ra; ra.accumulate = accumulate; ra.nseg = 2;
      ra.seg[0] = seg(sc_tail_w, grads + lay.wout(s), tc::kWgPartFloats, Rt, K, 64, 64);
      ra.seg[1] = seg(sc_tail_w + 4 * 4096, grads + lay.bout(s), tc::kWgPartFloats, Rt, 1, 64, K);
      if (s > 0) {
        ra.seg[2] = seg(sc_tail_w + 3 * 4096, grads + lay.win_w(s), tc::kWgPartFloats, Rt, 64, 64, K);
        ra.seg[3] = seg(sc_tail_w + 4 * 4096 + 192, grads + lay.win_b(s), tc::kWgPartFloats, Rt, 1, 64, 64);
        ra.nseg = 4;
      }
      if (launch_reduce(ra, wst)) return 1;
    }
    ReduceLayersArgs ra;
    ra.src0 = sc_layer; ra.dst0 = grads + lay.wd(s, 0);
    ra.layer_src_stride = (int64_t)R * tc::kWgPartFloats; ra.layer_dst_stride = Layout::kLayerParams;
    ra.part_stride = tc::kWgPartFloats; ra.P = R; ra.accumulate = accumulate;
    if (launch_pdl_grid("reduce_layers_kernel", reduce_layers_kernel, dim3((12288 + 4096 + 128 + 255) / 256, L), dim3(256), wst, ra))
      return 1;
    if (proj_tc && do_proj_wgrad_reduce(lay.dim, B * ((T + tc::TW - 1) / tc::TW), grads + lay.win_w(0), grads + lay.win_b(0), sc_proj,
                                        accumulate, wst))
      return 1;
    if (cudaEventRecord(pool().ev_stage[s], wst) != cudaSuccess) return fail("cudaEventRecord failed");
    // stage s+1's weight gradients ran under this stage's chain: absorb them now, so that on return (in stream
    // order) every gradient of stages > s is final; after the last stage absorb this one as well
    if (!last && cudaStreamWaitEvent(main, pool().ev_stage[s + 1], 0) != cudaSuccess) return fail("cudaStreamWaitEvent failed");
    if (s == 0 && cudaStreamWaitEvent(main, pool().ev_stage[0], 0) != cudaSuccess) return fail("cudaStreamWaitEvent failed");
    return lane_exit(main);
  }
  return 0;
}

int mstcn_backward(const mstcn_dims* d, const float* packed, const float* x, const int32_t* lens,
                   const int32_t* lens_host, int32_t groups, int32_t B, int32_t T, const mstcn_dropout* drop,
                   float* workspace, const uint8_t* winner, const float* gout, const float* gscale, float* grads,
                   int32_t accumulate, void* stream) {
  if (check_dims(d)) return 1;
  for (int s = d->num_stages - 1; s >= 0; --s)
    if (mstcn_backward_stage(d, packed, x, lens, lens_host, groups, B, T, drop, workspace, winner, gout, gscale, grads,
                             accumulate, s, stream))
      return 1;
  return 0;
}

int mstcn_proj_fwd(const float* x, int64_t n_frames, int32_t dim, const float* w_t, const float* bias, float* y,
                   void* stream) {
  if (!x || !w_t || !bias || !y) return fail("proj_fwd: NULL pointer");
  if (dim < 4 || dim % 4) return fail("proj_fwd: dim must be a positive multiple of 4");
  return do_proj_fwd(x, n_frames, dim, w_t, bias, y, S(stream));
}

int mstcn_proj_fwd_tc(const float* x, int64_t n_frames, int32_t dim, const float* wimg, const float* bias, const int32_t* lens,
                      int32_t T, float* y, void* stream) {
  if (!x || !wimg || !bias || !y) return fail("proj_fwd_tc: NULL pointer");
  if (dim < 4 || dim % 4) return fail("proj_fwd_tc: dim must be a positive multiple of 4");
  if (n_frames < 0 || n_frames >= (1LL << 31) / 64) return fail("proj_fwd_tc: n_frames out of range");
  return do_proj_fwd_tc(x, n_frames, dim, wimg, bias, lens, lens ? T : 0, y, S(stream));
}

int64_t mstcn_proj_bwd_scratch_floats(int32_t dim) { return proj_bwd_scratch(dim); }

int64_t mstcn_proj_wgrad_tc_scratch_floats(int32_t dim) { return proj_wgrad_tc_scratch(dim); }

int mstcn_proj_wgrad_tc(const float* x, const float* gy, const int32_t* lens, int32_t B, int32_t T, int32_t dim, float* gw,
                        float* gb, float* scratch, int32_t accumulate, void* stream) {
  if (!x || !gy || !lens || !gw || !gb || !scratch) return fail("proj_wgrad_tc: NULL pointer");
  if (dim < 4 || dim % 4) return fail("proj_wgrad_tc: dim must be a positive multiple of 4");
  if (B < 1 || T < 1 || (int64_t)B * T >= (1LL << 31) / 64) return fail("proj_wgrad_tc: B, T out of range");
  return do_proj_wgrad_tc(x, gy, lens, B, T, dim, gw, gb, scratch, accumulate, S(stream));
}

int mstcn_proj_bwd(const float* x, const float* gy, int64_t n_frames, int32_t dim, float* gw, float* gb, float* scratch,
                   int32_t accumulate, void* stream) {
  if (!x || !gy || !gw || !gb || !scratch) return fail("proj_bwd: NULL pointer");
  if (dim < 4 || dim % 4) return fail("proj_bwd: dim must be a positive multiple of 4");
  return do_proj_bwd(x, gy, n_frames, dim, gw, gb, scratch, accumulate, S(stream));
}

int mstcn_layer_fwd(const float* x, float* y, float* h_out, const int32_t* lens, int32_t B, int32_t T, int32_t dilation,
                    const float* wd_t, const float* bd, const float* w1_t, const float* b1, const mstcn_dropout* drop,
                    int32_t layer_id, void* stream) {
  if (!x || !y || !lens || !wd_t || !bd || !w1_t || !b1) return fail("layer_fwd: NULL pointer");
  if (B < 1 || T < 1 || dilation < 1) return fail("layer_fwd: bad B/T/dilation");
  return do_layer_fwd(x, y, h_out, lens, B, T, dilation, wd_t, bd, w1_t, b1, drop, layer_id, S(stream));
}

int mstcn_layer_fwd_tc(const float* x, float* y, float* h_out, const int32_t* lens, int32_t B, int32_t T,
                       int32_t dilation, const float* wimg, const float* bd, const float* b1, const mstcn_dropout* drop,
                       int32_t layer_id, void* stream) {
  if (!x || !y || !lens || !wimg || !bd || !b1) return fail("layer_fwd_tc: NULL pointer");
  if (B < 1 || T < 1 || dilation < 1) return fail("layer_fwd_tc: bad B/T/dilation");
  return do_layer_fwd_tc(x, y, h_out, lens, B, T, dilation, wimg, bd, b1, drop, layer_id, S(stream));
}

int mstcn_stage_fwd_tc(const mstcn_dims* d, const float* packed, int32_t stage, float* planes, float* h_planes,
                       const int32_t* lens, int32_t B, int32_t T, const mstcn_dropout* drop, int32_t* flags, void* stream) {
  if (check_dims(d)) return 1;
  if (!use_tc(d)) return fail("stage_fwd_tc: dims.flags lacks MSTCN_FLAG_TENSOR_CORES");
  if (!packed || !planes || !lens || !flags) return fail("stage_fwd_tc: NULL pointer");
  if (B < 1 || T < 1 || stage < 0 || stage >= d->num_stages) return fail("stage_fwd_tc: bad B/T/stage");
  Layout lay = make_layout(d);
  const int64_t nt = (int64_t)B * ((T + tc::TM - 1) / tc::TM);
  if (lane_enter(S(stream))) return 1;
  if (cudaMemsetAsync(flags, 0, sizeof(int) * lay.L * nt, S(stream)) != cudaSuccess) return fail("stage_fwd_tc: clearing the tile flags failed");
  if (do_stage_fwd_tc(lay, packed, stage, planes, h_planes, lens, B, T, drop, flags, S(stream))) return 1;
  return lane_exit(S(stream));
}

int mstcn_layer_bwd_gx_tc(const float* gu, const float* gy, float* gx, const int32_t* lens, int32_t B, int32_t T,
                          int32_t dilation, const float* wimg_b, void* stream) {
  if (!gu || !gy || !gx || !lens || !wimg_b) return fail("layer_bwd_gx_tc: NULL pointer");
  if (B < 1 || T < 1 || dilation < 1) return fail("layer_bwd_gx_tc: bad B/T/dilation");
  return do_layer_bwd_gx_tc(gu, gy, gx, lens, B, T, dilation, wimg_b, S(stream));
}

int mstcn_debug_tc_timing(int64_t* device_buf) {
  g_tc_dbg = reinterpret_cast<long long*>(device_buf);
  return 0;
}

int64_t mstcn_dp_flag_words(void) { return dp::kFlagWords; }

int mstcn_dp_allreduce(float* const* peer_bufs, uint32_t* const* peer_flags, float* mc_buf, int64_t offset, int64_t n,
                       int32_t rank, int32_t world, int32_t channel, void* stream) {
  if (!peer_bufs || !peer_flags) return fail("dp_allreduce: NULL peer arrays");
  if (world < 1 || world > dp::kMaxRanks || rank < 0 || rank >= world) return fail("dp_allreduce: bad rank / world (at most 16 ranks)");
  if (channel < 0 || channel >= dp::kChannels) return fail("dp_allreduce: channel must be in [0, 8)");
  if (offset < 0 || n < 0) return fail("dp_allreduce: bad range");
  if (n == 0 || world == 1) return 0;
  dp::DpArgs a;
  a.bufs = peer_bufs; a.flags = peer_flags; a.mc = mc_buf; a.offset = offset; a.n = n;
  a.rank = rank; a.world = world; a.channel = channel;
  // one CTA per SM at most (it shares the SM with a chain CTA); every rank derives the same grid from n alone
  const int64_t units = ((offset | n) & 3) == 0 ? n / 4 : n;
  int64_t grid = (units + dp::kThreads - 1) / dp::kThreads;
  const int cap = dp::kMaxCtas < 148 ? dp::kMaxCtas : 148;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  dp::dp_allreduce_kernel<<<(int)grid, dp::kThreads, 0, S(stream)>>>(a);
  return check_launch("dp_allreduce_kernel");
}

int mstcn_debug_trap_report(int64_t* host_words) {
  // host_words: >= 8 int64 of PINNED, device-mapped host memory (cudaHostAlloc / torch pin_memory), zeroed; NULL detaches
  long long* dev = nullptr;
  if (host_words != nullptr && cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), host_words, 0) != cudaSuccess)
    return fail("debug_trap_report: the buffer is not pinned host memory");
  if (cudaMemcpyToSymbol(tc::g_trap_report, &dev, sizeof dev) != cudaSuccess) return fail("debug_trap_report: cudaMemcpyToSymbol failed");
  return 0;
}

int mstcn_debug_backward_timing(int32_t enable) {
  g_bwd_timing = enable != 0;
  return 0;
}

int mstcn_debug_backward_times(float* out, int32_t n) {
  if (!out || n < 0) return fail("debug_backward_times: bad arguments");
  for (int i = 0; i < n && i < 32; ++i) out[i] = g_bwd_times[i];
  return 0;
}

int mstcn_debug_chain_trace(int64_t* device_buf) {
  g_tc_trace = reinterpret_cast<long long*>(device_buf);
  return 0;
}

int64_t mstcn_layer_bwd_scratch_floats(void) { return layer_bwd_scratch(); }

int mstcn_layer_bwd(const float* x, const float* h, const float* gy, float* gx, float* gu, const int32_t* lens, int32_t B,
                    int32_t T, int32_t dilation, const float* wd_b, const float* w1, const mstcn_dropout* drop,
                    int32_t layer_id, float* gwd, float* gbd, float* gw1, float* gb1, float* scratch, int32_t accumulate,
                    void* stream) {
  if (!x || !h || !gy || !gx || !gu || !lens || !wd_b || !w1 || !gwd || !gbd || !gw1 || !gb1 || !scratch)
    return fail("layer_bwd: NULL pointer");
  if (B < 1 || T < 1 || dilation < 1) return fail("layer_bwd: bad B/T/dilation");
  return do_layer_bwd(x, h, gy, gx, gu, lens, B, T, dilation, wd_b, w1, drop, layer_id, gwd, gbd, gw1, gb1, scratch,
                      accumulate, S(stream));
}

int mstcn_tail_fwd(const float* a, const int32_t* lens, int32_t B, int32_t T, int32_t n_class, int32_t stage,
                   const float* wout_t, const float* bout, float* logits, float* out, uint8_t* winner, const float* wn_t,
                   const float* bn, float* next_x0, void* stream) {
  if (!a || !lens || !wout_t || !bout || !logits || !out || !winner) return fail("tail_fwd: NULL pointer");
  if (n_class < 1 || n_class > MSTCN_KMAX) return fail("tail_fwd: n_class must be in [1, 64]");
  if (next_x0 && (!wn_t || !bn)) return fail("tail_fwd: next-stage weights missing");
  return do_tail_fwd(a, lens, B, T, n_class, stage, wout_t, bout, logits, out, winner, wn_t, bn, next_x0, S(stream));
}

int64_t mstcn_tail_bwd_scratch_floats(void) { return tail_bwd_scratch(); }

int mstcn_tail_bwd(const float* a, const float* logits, const float* gout, const float* gscale, const uint8_t* winner,
                   const float* gin, const int32_t* lens, int32_t B, int32_t T, int32_t n_class, int32_t stage,
                   const float* wout_b, const float* wn_b, float* ga, float* gwout, float* gbout, float* gwn, float* gbn,
                   float* scratch, int32_t accumulate, void* stream) {
  if (!a || !logits || !gout || !winner || !lens || !wout_b || !ga || !gwout || !gbout || !scratch)
    return fail("tail_bwd: NULL pointer");
  if (gin && (!wn_b || !gwn || !gbn)) return fail("tail_bwd: next-stage pointers missing");
  if (n_class < 1 || n_class > MSTCN_KMAX) return fail("tail_bwd: n_class must be in [1, 64]");
  return do_tail_bwd(a, logits, gout, gscale, winner, gin, lens, B, T, n_class, stage, wout_b, wn_b, ga, gwout, gbout,
                     gwn, gbn, scratch, accumulate, S(stream));
}

int64_t mstcn_ce_scratch_floats(int64_t n_rows) {
  int64_t blocks = (n_rows + 7) / 8;
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  return 2 * blocks;
}

int mstcn_ce_loss(const float* logits, const int64_t* labels, int64_t n_rows, int32_t n_class, int64_t n_valid_override,
                  float* gout, float* result, float* scratch, void* stream) {
  if (!logits || !labels || !gout || !result || !scratch) return fail("ce_loss: NULL pointer");
  if (n_class < 1 || n_class > MSTCN_KMAX) return fail("ce_loss: n_class must be in [1, 64]");
  const int blocks = (int)(mstcn_ce_scratch_floats(n_rows) / 2);
  ce_loss_kernel<<<blocks, 256, 0, S(stream)>>>(logits, labels, n_rows, n_class, gout, scratch);
  if (check_launch("ce_loss_kernel")) return 1;
  ce_finalize_kernel<<<1, 256, 0, S(stream)>>>(scratch, blocks, n_valid_override, result);
  return check_launch("ce_finalize_kernel");
}

int mstcn_loss_head(const mstcn_dims* d, float* workspace, int32_t B, int32_t T, const int64_t* labels, int64_t n_valid,
                    float* out, uint8_t* winner, float* result, float* scratch, void* stream) {
  if (check_dims(d)) return 1;
  if (!use_tc_bwd(d)) return fail("loss_head: needs the tensor-core path (it writes the routed gradient planes of its backward)");
  if (!workspace || !labels || !result || !scratch) return fail("loss_head: NULL pointer");
  if ((out == nullptr) != (winner == nullptr)) return fail("loss_head: out and winner go together");
  if (B < 1 || T < 1 || n_valid < 1) return fail("loss_head: B, T and n_valid must be >= 1");
  Ws w = carve(d, B, T, true, workspace);
  const int blocks = (int)(mstcn_ce_scratch_floats(w.N) / 2);
  tc::loss_head_kernel<<<blocks, 256, 0, S(stream)>>>(w.logits(0), w.act_stage, w.S, w.K, w.N, labels,
                                                       (float)(1.0 / (double)n_valid), w.gr(0), w.N * 64, out, winner, scratch);
  if (check_launch("loss_head_kernel")) return 1;
  ce_finalize_kernel<<<1, 256, 0, S(stream)>>>(scratch, blocks, n_valid, result);
  return check_launch("ce_finalize_kernel");
}

int64_t mstcn_paper_loss_scratch_floats(int32_t S, int64_t n_rows) {
  int64_t blocks = ((int64_t)S * n_rows + 7) / 8;
  if (blocks > 2048) blocks = 2048;
  if (blocks < 1) blocks = 1;
  return 2 * blocks;
}

int mstcn_paper_loss(const float* stage_logits, const int64_t* labels, const int32_t* lens, int32_t S, int32_t B, int32_t T,
                     int32_t n_class, int64_t n_valid, float lam, float tau, float* gstage, float* result, float* scratch,
                     void* stream) {
  if (!stage_logits || !labels || !lens || !gstage || !result || !scratch) return fail("paper_loss: NULL pointer");
  if (n_class < 1 || n_class > MSTCN_KMAX) return fail("paper_loss: n_class must be in [1, 64]");
  if (S < 1 || B < 1 || T < 1) return fail("paper_loss: S, B and T must be >= 1");
  if (n_valid < 1) return fail("paper_loss: n_valid must be >= 1");
  PaperLossArgs a;
  a.z = stage_logits; a.labels = labels; a.lens = lens; a.g = gstage; a.part = scratch;
  a.S = S; a.B = B; a.T = T; a.K = n_class;
  a.inv_nvalid = (float)(1.0 / (double)n_valid);
  a.tmse_scale = T > 1 ? (float)((double)lam / ((double)B * (T - 1) * n_class)) : 0.f;   // torch's mean over (B, K, T-1)
  a.tau2 = tau * tau;
  const int blocks = (int)(mstcn_paper_loss_scratch_floats(S, (int64_t)B * T) / 2);
  paper_loss_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  if (check_launch("paper_loss_kernel")) return 1;
  paper_loss_finalize_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(scratch, blocks, a.inv_nvalid, a.tmse_scale, result);
  return check_launch("paper_loss_finalize_kernel");
}

int mstcn_pad_batch(const float* feats, const int64_t* labels, const int64_t* offsets, const int32_t* video_idx,
                    int32_t B, int32_t T, int32_t dim, float* x, int64_t* y, int32_t* lens_out, void* stream) {
  if (!feats || !offsets || !video_idx || !x) return fail("pad_batch: NULL pointer");
  if (B < 1 || T < 1) return fail("pad_batch: B and T must be >= 1");
  if (dim < 4 || dim % 4 != 0) return fail("pad_batch: dim must be a positive multiple of 4");
  if ((reinterpret_cast<uintptr_t>(feats) | reinterpret_cast<uintptr_t>(x)) & 15) return fail("pad_batch: buffers must be 16-byte aligned");
  const int64_t rows = (int64_t)B * T;
  pad_batch_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, S(stream)>>>(feats, labels, offsets, video_idx, B, T, dim, x, y, lens_out);
  return check_launch("pad_batch_kernel");
}

int mstcn_frame_argmax(const float* logits, int64_t n_rows, int32_t n_class, int64_t* idx, float* val, void* stream) {
  if (!logits || !idx) return fail("frame_argmax: NULL pointer");
  if (n_class < 1) return fail("frame_argmax: n_class must be >= 1");
  if (n_rows == 0) return 0;
  frame_argmax_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, S(stream)>>>(logits, n_rows, n_class, idx, val);
  return check_launch("frame_argmax_kernel");
}

int mstcn_segment_vote(const int64_t* pred, const int32_t* bounds, int32_t n_seg, int32_t n_class,
                       int32_t inference_fallback, int32_t* labels_out, void* stream) {
  if (!pred || !bounds || !labels_out) return fail("segment_vote: NULL pointer");
  if (n_class < 1 || n_class > MSTCN_KMAX) return fail("segment_vote: n_class must be in [1, 64]");
  if (n_seg <= 0) return 0;
  segment_vote_kernel<<<n_seg, 128, 0, S(stream)>>>(pred, bounds, n_class, inference_fallback, labels_out);
  return check_launch("segment_vote_kernel");
}

int mstcn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, int32_t step, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return fail("adam_step: NULL pointer");
  if (step < 1) return fail("adam_step: step counts from 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * 148) blocks = 4 * 148;
  if (blocks < 1) return 0;
  adam_kernel<<<blocks, 256, 0, S(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, step_size, 1.f - beta1, beta2,
                                               1.f - beta2, bc2_sqrt, eps);
  return check_launch("adam_kernel");
}

int mstcn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev,
                        float beta1, float beta2, float eps, int64_t* step_dev, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev) return fail("adam_step_dev: NULL pointer");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * 148) blocks = 4 * 148;
  if (blocks < 1) return 0;
  adam_dev_kernel<<<blocks, 256, 0, S(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, lr_dev, beta1, beta2, eps,
                                                   reinterpret_cast<const long long*>(step_dev));
  if (check_launch("adam_dev_kernel")) return 1;
  adam_bump_kernel<<<1, 1, 0, S(stream)>>>(reinterpret_cast<long long*>(step_dev));
  return check_launch("adam_bump_kernel");
}

int mstcn_dropout_scale(const mstcn_dropout* drop, int32_t layer_id, int64_t n_frames, float* out, void* stream) {
  if (!drop || !out) return fail("dropout_scale: NULL pointer");
  const int64_t n = n_frames * 16;
  dropout_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(drop->seed, drop->offset, reinterpret_cast<const unsigned long long*>(drop->offset_dev), (uint32_t)layer_id,
                                                                         n_frames, out);
  return check_launch("dropout_scale_kernel");
}

}  // extern "C"
