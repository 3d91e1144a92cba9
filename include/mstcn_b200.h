/*
 * mstcn_b200.h -- C ABI of the B200-native MS-TCN hot path (libmstcn_b200.so).
 *
 * The reference (mrqorib/pytorch-video-action) has NO native / FFI interface: its hot path
 * is the Python class networks.MultiStageModel (networks.py:298-347) driven by
 * train.py:298-332 and inference.py:113-179, and every GPU kernel it runs is a PyTorch
 * library call.  This header is therefore the boundary a maintainer binds with ctypes
 * (see INTEGRATION.md); each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the
 *     name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - int return: 0 = ok, non-zero = error; text via mstcn_last_error() (thread-local).
 *   - no device allocation inside; global state is limited to thread-local tensor-map caches, one lazily created
 *     pool of internal side streams / events (the backward's weight-gradient stream) and the optional profiling
 *     hooks; re-entrant (autograd's engine thread calls the backward entries).
 *   - tensor-core path: the layers of a stage run as one persistent "chain" launch whose CTAs wait on each other's
 *     tiles, and consecutive kernels are linked by per-tile flags in the workspace.  Two such launch sequences must
 *     not share one GPU concurrently, so the library serialises them per device: an entry issued on another stream
 *     than the previous one first makes its stream wait for that sequence's end (an internal event; a no-op on the
 *     same stream, and left to the caller inside a stream capture, where a captured step is one stream-ordered unit).
 *     Waits are bounded: a launch that still cannot make progress traps after ~4 s instead of hanging, and
 *     mstcn_debug_trap_report tells which wait it was.
 *   - activations are channels-last fp32: frame n = b*T + t owns 64 contiguous floats;
 *     logits are (B*T, n_class) row-major, exactly what MultiStageModel.forward returns
 *     (networks.py:317-320).
 *   - num_f_maps must be 64 and n_class <= 64 (the reference's only configuration is
 *     64 / 48; others are rejected loudly, never routed to a fallback).
 */
#ifndef MSTCN_B200_H_
#define MSTCN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSTCN_ABI_VERSION 2
#define MSTCN_C 64          /* num_f_maps the kernels are specialised for */
#define MSTCN_KMAX 64       /* largest n_class */

/* MultiStageModel.__init__(dim, num_stages, num_layers, num_f_maps, n_class), networks.py:299 */
typedef struct mstcn_dims {
  int32_t dim;
  int32_t num_stages;
  int32_t num_layers;
  int32_t num_f_maps;
  int32_t n_class;
  int32_t flags;            /* MSTCN_FLAG_* */
} mstcn_dims;

/* run the dilated residual layers on the tcgen05 tensor cores with error-compensated 3xTF32
 * (fp32-equivalent) instead of the fp32 FFMA kernels */
#define MSTCN_FLAG_TENSOR_CORES 1
/* with MSTCN_FLAG_TENSOR_CORES: keep the layer BACKWARD on the fp32 FFMA kernels (diagnostics) */
#define MSTCN_FLAG_FFMA_BACKWARD 2
/* with MSTCN_FLAG_TENSOR_CORES (and without MSTCN_FLAG_FFMA_BACKWARD): mstcn_pack_params refreshes only what the
 * tensor-core path reads -- the biases and the operand images -- and leaves the transposed fp32 operands of the
 * FFMA kernels (which = 0,2,3,5,7,8,9,11) stale.  Saves ~15 us per optimizer step. */
#define MSTCN_FLAG_PACK_TC_ONLY 4

/* dropout stream: Philox4x32-10, key=(seed), counter=(frame, global_layer, offset) */
typedef struct mstcn_dropout {
  int32_t  enabled;         /* 1 = train mode (nn.Dropout p=0.5 active, networks.py:341,346) */
  int32_t  _pad;
  uint64_t seed;
  uint64_t offset;          /* bump once per forward call */
  const uint64_t* offset_dev; /* optional DEVICE counter added to `offset` when the kernels run: lets a captured
                               * CUDA graph draw a fresh mask on every replay (bump it inside the graph); NULL = unused */
} mstcn_dropout;

int         mstcn_abi_version(void);
const char* mstcn_last_error(void);
/* number of SMs of the current device (grid sizing), <0 on error */
int         mstcn_sm_count(void);

/* ---- parameter layout ------------------------------------------------------------------
 * `params` is ONE flat fp32 buffer holding the 176 (at 4x10) state_dict tensors in
 * state_dict order and native nn.Conv1d (out, in, tap) layout.  `packed` is
 * the kernel-side copy (transposed GEMM operands, zero-padded class dim) rebuilt by
 * mstcn_pack_params after every optimizer step. */
int64_t mstcn_param_count(const mstcn_dims* d);
int64_t mstcn_packed_count(const mstcn_dims* d);
/* offset (in floats) of tensor `index` (state_dict order) inside the flat buffer; -1 if out of range */
int64_t mstcn_param_offset(const mstcn_dims* d, int32_t index);
int32_t mstcn_param_tensors(const mstcn_dims* d);
/* float offset of one packed operand: which = 0 win_t (din,64) | 1 bin | 2 win_b (64,64pad) |
 * 3 wd_t (3,in,out) | 4 bd | 5 w1_t (in,out) | 6 b1 | 7 wd_b (3,out,in) | 8 w1_n (out,in) |
 * 9 wout_t (64,64pad) | 10 bout (64pad) | 11 wout_b (64pad,64) | 12 / 13 tensor-core forward / backward
 * operand image of the layer | 14 tensor-core image of the stage-1 input projection;
 * `layer` is ignored for stage-level operands */
int64_t mstcn_packed_offset(const mstcn_dims* d, int32_t stage, int32_t layer, int32_t which);
int     mstcn_pack_params(const mstcn_dims* d, const float* params, float* packed, void* stream);

/* ---- workspace ------------------------------------------------------------------------ */
/* floats needed by mstcn_forward (+ mstcn_backward when training != 0) for a (B, T) batch */
int64_t mstcn_workspace_floats(const mstcn_dims* d, int32_t B, int32_t T, int32_t training);

/* float offset inside the workspace of: what = 0 input plane of `layer` (layer == num_layers: the stage's last
 * output), 1 relu output of `layer`, 2 the stage's masked logits (B*T, n_class), 3 the stage's softmax*mask
 * (B*T, 64); -1 if that plane does not exist in this mode.  Diagnostics / tests. */
int64_t mstcn_workspace_offset(const mstcn_dims* d, int32_t B, int32_t T, int32_t training, int32_t what,
                               int32_t stage, int32_t layer);

/* ---- whole-model entries ---------------------------------------------------------------
 * mstcn_forward  = MultiStageModel.forward(x, x_len), networks.py:305-320.
 *   x (B,T,dim) batch-first fp32; lens (B) int32 on device, max(lens)==T is the caller's
 *   contract; out (B*T, n_class) = max over stages; winner (B*T, n_class) uint8 = index of
 *   the winning stage (first on ties, like torch.max).  training!=0 keeps what backward needs; on the tensor-core
 *   training path out and winner may be NULL (mstcn_loss_head takes the max).
 * mstcn_backward = what loss.backward() (train.py:328) replays: gout (B*T, n_class) is
 *   dLoss/dout, optionally scaled by the device scalar *gscale (NULL = 1); writes the flat
 *   gradient buffer `grads` (same layout as `params`); accumulate!=0 adds instead of overwriting.
 *   gout == NULL (tensor-core path): the gradient planes were already written by mstcn_loss_head.
 *   winner == NULL selects the per-stage mode: gout is then (num_stages, B*T, n_class) = dLoss/dz_s for every
 *   stage's masked logits (what a loss on the per-stage outputs produces, e.g. mstcn_paper_loss), no max routing.
 * lens_host (optional HOST copy of lens) + groups (1..4): every op is per-video, so the batch is cut into
 *   `groups` contiguous video ranges with ~equal tile counts whose kernel chains run on concurrent
 *   internal streams forked from / joined to `stream` (FFMA path only: the tensor-core path runs every stage's
 *   layers as one chain launch on `stream`).  lens_host == NULL or groups <= 1: one chain on `stream`. */
int mstcn_forward(const mstcn_dims* d, const float* packed, const float* x, const int32_t* lens,
                  const int32_t* lens_host, int32_t groups,
                  int32_t B, int32_t T, const mstcn_dropout* drop, int32_t training,
                  float* workspace, float* out, uint8_t* winner, void* stream);
int mstcn_backward(const mstcn_dims* d, const float* packed, const float* x, const int32_t* lens,
                   const int32_t* lens_host, int32_t groups,
                   int32_t B, int32_t T, const mstcn_dropout* drop,
                   float* workspace, const uint8_t* winner, const float* gout, const float* gscale,
                   float* grads, int32_t accumulate, void* stream);
/* one stage of the above (call with stage = num_stages-1 ... 0).  A stage's dilated-layer weight gradients
 * are computed by one kernel on an internal side stream, under the next stage's chain.  In stream order
 * after the call for stage s, every gradient of stages > s is final, i.e. the contiguous ranges
 * [layers(k,0) .. layers(k+1,0)) for k > s (stage k's layers and class head plus stage k+1's input
 * projection) -- the buckets a data-parallel caller can all-reduce while stage s-1 is still running
 * (SURVEY.md 8e); after the call for stage 0 the whole buffer is final. */
int mstcn_backward_stage(const mstcn_dims* d, const float* packed, const float* x, const int32_t* lens,
                         const int32_t* lens_host, int32_t groups,
                         int32_t B, int32_t T, const mstcn_dropout* drop,
                         float* workspace, const uint8_t* winner, const float* gout, const float* gscale,
                         float* grads, int32_t accumulate, int32_t stage, void* stream);
/* float offset of stage s's first dilated layer inside the flat parameter/gradient buffer
 * (s == num_stages returns the total) -- the bucket boundaries for the above. */
int64_t mstcn_bucket_boundary(const mstcn_dims* d, int32_t stage);

/* ---- fused units (one kernel each; used by the whole-model entries and by the tests) --- */
/* SingleStageModel.conv_1x1 of stage 1, NOT masked (networks.py:325,330): y = x W^T + b.
 * w_t is (dim,64) (packed form). */
int mstcn_proj_fwd(const float* x, int64_t n_frames, int32_t dim, const float* w_t, const float* bias,
                   float* y, void* stream);
/* same contract on the tensor cores (3xTF32): x is read in place through a 2-D tensor map (16-byte aligned,
 * dim % 4 == 0); wimg = the projection's operand image inside `packed` (mstcn_packed_offset(.., which = 14)).
 * lens/T (optional): tiles lying wholly in one video's zero padding are answered with the bias directly. */
int mstcn_proj_fwd_tc(const float* x, int64_t n_frames, int32_t dim, const float* wimg, const float* bias,
                      const int32_t* lens, int32_t T, float* y, void* stream);
/* its weight/bias gradient (input features need no grad): gw native (64,dim), gb (64).
 * scratch: >= mstcn_proj_bwd_scratch_floats(dim) floats. */
int64_t mstcn_proj_bwd_scratch_floats(int32_t dim);
int mstcn_proj_bwd(const float* x, const float* gy, int64_t n_frames, int32_t dim, float* gw, float* gb,
                   float* scratch, int32_t accumulate, void* stream);
/* the same gradient on the tensor cores (exact 4-term tf32, the weight-gradient kernel in projection mode): x (B,T,dim)
 * read in place by TMA, gy (B*T,64) the gradient of the projection output, lens (B) device int32 (the conv is unmasked:
 * every frame counts, networks.py:330).  scratch: >= mstcn_proj_wgrad_tc_scratch_floats(dim) floats. */
int64_t mstcn_proj_wgrad_tc_scratch_floats(int32_t dim);
int mstcn_proj_wgrad_tc(const float* x, const float* gy, const int32_t* lens, int32_t B, int32_t T, int32_t dim,
                        float* gw, float* gb, float* scratch, int32_t accumulate, void* stream);

/* DilatedResidualLayer.forward (networks.py:343-347):
 *   y = (x + drop(W1 relu(Wd (*)_d x + bd) + b1)) * mask.   h_out (may be NULL) keeps relu(.)
 * wd_t (3,64in,64out), w1_t (64in,64out) are the packed forms. */
int mstcn_layer_fwd(const float* x, float* y, float* h_out, const int32_t* lens, int32_t B, int32_t T,
                    int32_t dilation, const float* wd_t, const float* bd, const float* w1_t, const float* b1,
                    const mstcn_dropout* drop, int32_t layer_id, void* stream);
/* same contract on the tcgen05 tensor cores (3xTF32, TMA-fed, accumulators in TMEM).
 * wimg = the layer's operand image inside `packed` (offset mstcn_packed_offset(.., which = 12)). */
int mstcn_layer_fwd_tc(const float* x, float* y, float* h_out, const int32_t* lens, int32_t B, int32_t T,
                       int32_t dilation, const float* wimg, const float* bd, const float* b1,
                       const mstcn_dropout* drop, int32_t layer_id, void* stream);
/* All num_layers DilatedResidualLayers of one stage (SingleStageModel.forward's loop, networks.py:331-332) as ONE
 * persistent "chain" launch: tasks (layer, 128-frame tile) are dealt round-robin to the CTAs and a task starts as soon
 * as the previous layer's tiles under its three taps are flagged complete (tile-level dataflow, no kernel boundary
 * between layers).  planes = num_layers+1 contiguous (B*T,64) planes: plane 0 holds the stage input, plane l+1
 * receives layer l's output; h_planes (num_layers planes, may be NULL) receive relu(.) for the backward.
 * flags = num_layers * B * ceil(T/128) int32 of scratch (cleared here).  Bit-identical to num_layers calls of
 * mstcn_layer_fwd_tc.  Two chain launches must never run concurrently on one GPU (they spin on their own tiles). */
int mstcn_stage_fwd_tc(const mstcn_dims* d, const float* packed, int32_t stage, float* planes, float* h_planes,
                       const int32_t* lens, int32_t B, int32_t T, const mstcn_dropout* drop, int32_t* flags,
                       void* stream);
/* input-gradient half of the layer backward on the tensor cores:
 *   gx[t] = gy[t]*mask + sum_k Wd[:,:,k]^T gu[t-(k-1)d]   (gu = dL/d(pre-ReLU), from mstcn_layer_bwd's pass A).
 * wimg_b = the layer's backward operand image (mstcn_packed_offset(.., which = 13)). */
int mstcn_layer_bwd_gx_tc(const float* gu, const float* gy, float* gx, const int32_t* lens, int32_t B, int32_t T,
                          int32_t dilation, const float* wimg_b, void* stream);
/* its backward. gy = dL/dy; writes gx = dL/dx and native-layout weight grads
 * gwd (64,64,3), gbd (64), gw1 (64,64), gb1 (64).  wd_b (3,64out,64in) packed, w1 native (64out,64in).
 * gu: scratch (B*T,64); scratch: >= mstcn_layer_bwd_scratch_floats() floats. */
int64_t mstcn_layer_bwd_scratch_floats(void);
int mstcn_layer_bwd(const float* x, const float* h, const float* gy, float* gx, float* gu,
                    const int32_t* lens, int32_t B, int32_t T, int32_t dilation,
                    const float* wd_b, const float* w1, const mstcn_dropout* drop, int32_t layer_id,
                    float* gwd, float* gbd, float* gw1, float* gb1,
                    float* scratch, int32_t accumulate, void* stream);

/* Stage tail (networks.py:333 conv_out*mask, :312-319 running max over stages, :314
 * softmax*mask, :330 next stage's unmasked conv_1x1):
 *   z = (Wout a + bout) * mask -> logits (kept for backward); out/winner updated with stage s;
 *   if next_* != NULL: next_x0 = Wn (softmax(z) * mask) + bn.
 * wout_t (64, 64pad) , wn_t (64pad, 64) packed forms. */
int mstcn_tail_fwd(const float* a, const int32_t* lens, int32_t B, int32_t T, int32_t n_class, int32_t stage,
                   const float* wout_t, const float* bout, float* logits, float* out, uint8_t* winner,
                   const float* wn_t, const float* bn, float* next_x0, void* stream);
/* its backward: gz = [winner==stage] gout*gscale + softmax-backward(Wn^T gin); then
 * ga = Wout^T gz, gwout (K,64), gbout (K), and for the next stage's projection gwn (64,K), gbn (64).
 * wout_b (64pad,64) and wn_b (64,64pad) are packed forms.  gin==NULL for the last stage. */
int64_t mstcn_tail_bwd_scratch_floats(void);
int mstcn_tail_bwd(const float* a, const float* logits, const float* gout, const float* gscale,
                   const uint8_t* winner, const float* gin, const int32_t* lens,
                   int32_t B, int32_t T, int32_t n_class, int32_t stage,
                   const float* wout_b, const float* wn_b,
                   float* ga, float* gwout, float* gbout, float* gwn, float* gbn,
                   float* scratch, int32_t accumulate, void* stream);

/* nn.CrossEntropyLoss(ignore_index=-1) forward+backward in one pass (train.py:266-267,326).
 * labels int64 (N); gout (N,K) receives (softmax - onehot) on valid rows (UNNORMALISED);
 * result[0] = mean loss, result[1] = 1/n_valid (feed as gscale), result[2] = n_valid.
 * n_valid_override > 0 replaces the divisor (data-parallel shards divide by the global count).
 * scratch: >= mstcn_ce_scratch_floats(N) floats. */
int64_t mstcn_ce_scratch_floats(int64_t n_rows);
int mstcn_ce_loss(const float* logits, const int64_t* labels, int64_t n_rows, int32_t n_class,
                  int64_t n_valid_override, float* gout, float* result, float* scratch, void* stream);

/* Fused loss head of the reference's training step on the tensor-core path: max over stages (networks.py:319) +
 * CrossEntropyLoss(ignore_index=-1, mean over n_valid) (train.py:267,326) + the backward of both, in one pass over the
 * per-stage logits of a training workspace filled by mstcn_forward (which may then be called with out = winner = NULL).
 * dLoss/dz_s goes straight into the workspace's routed-gradient planes: follow with mstcn_backward(gout = NULL).
 * out / winner (optional, both or neither) receive what mstcn_forward would have returned.  result as mstcn_ce_loss;
 * scratch >= mstcn_ce_scratch_floats(B*T).  Bit-identical to mstcn_forward + mstcn_ce_loss + mstcn_backward(gout). */
int mstcn_loss_head(const mstcn_dims* d, float* workspace, int32_t B, int32_t T, const int64_t* labels, int64_t n_valid,
                    float* out, uint8_t* winner, float* result, float* scratch, void* stream);

/* Canonical MS-TCN loss (Farha & Gall, CVPR 2019) -- NOT in the reference, parity unpinned:
 *   sum_s [ CE(z_s, y; ignore -1, mean over n_valid) + lam * mean_{b,c,t>=1}( clamp((logp_s[t] - logp_s[t-1].detach())^2, 0, tau^2) * m[b,t] ) ]
 * forward + backward in one pass.  stage_logits (S, B*T, n_class) = the workspace's per-stage logits
 * (mstcn_workspace_offset what = 2, stacked); gstage (S, B*T, n_class) receives dLoss/dz_s -- feed it to
 * mstcn_backward with winner = NULL.  result[0] = loss, [1] = CE part, [2] = T-MSE part.
 * scratch: >= mstcn_paper_loss_scratch_floats(S, B*T) floats. */
int64_t mstcn_paper_loss_scratch_floats(int32_t S, int64_t n_rows);
int mstcn_paper_loss(const float* stage_logits, const int64_t* labels, const int32_t* lens, int32_t S, int32_t B, int32_t T,
                     int32_t n_class, int64_t n_valid, float lam, float tau, float* gstage, float* result, float* scratch,
                     void* stream);

/* pad_batch on the device (train.py:183-205): the dataset's frames are resident in HBM, concatenated --
 * feats (sum T_i, dim) fp32, labels (sum T_i,) int64 (may be NULL), offsets (V+1,) int64 frame offsets.
 * Gathers videos video_idx[0..B) (device int32) into x (B, T, dim): frames beyond a video's length are zero;
 * y (B*T,) int64 (may be NULL) gets the labels and -1 (_TARGET_PAD) beyond the length; lens_out (B,) int32
 * (may be NULL) gets min(len, T) -- the device copy of x_len the kernels take.  T must be >= the longest
 * selected video for reference semantics (max_length); dim % 4 == 0. */
int mstcn_pad_batch(const float* feats, const int64_t* labels, const int64_t* offsets, const int32_t* video_idx,
                    int32_t B, int32_t T, int32_t dim, float* x, int64_t* y, int32_t* lens_out, void* stream);

/* per-frame argmax (train.py:157, inference.py:123): first index on ties */
int mstcn_frame_argmax(const float* logits, int64_t n_rows, int32_t n_class,
                       int64_t* idx, float* val, void* stream);
/* segment majority vote (train.py:161-170; inference.py:129-151 when inference_fallback!=0).
 * bounds (n_seg+1) int32 frame boundaries into pred; labels_out (n_seg) int32. */
int mstcn_segment_vote(const int64_t* pred, const int32_t* bounds, int32_t n_seg, int32_t n_class,
                       int32_t inference_fallback, int32_t* labels_out, void* stream);

/* torch.optim.Adam step over the flat buffers (train.py:273,329), eps outside the sqrt.
 * step counts from 1. */
int mstcn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                    float lr, float beta1, float beta2, float eps, int32_t step, void* stream);

/* The same step with its state in device memory, so that it can sit inside a captured CUDA graph: *step_dev (int64) holds
 * the number of steps taken so far and is advanced by the call; *lr_dev (float) is read at execution time (a scheduler
 * rewrites it in stream order). */
int mstcn_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const float* lr_dev, float beta1, float beta2, float eps, int64_t* step_dev, void* stream);

/* Data-parallel gradient sum over peer memory (the reference is single-GPU, train.py:181): sums floats
 * [offset, offset + n) of every rank's flat gradient buffer IN PLACE, every rank ending with identical bits (one-shot:
 * each rank adds the world's buckets in rank order straight out of the peers' HBM over NVLink; with a multicast mapping
 * `mc_buf` the NVSwitch adds them -- multimem.ld_reduce / multimem.st).  Every rank calls it with the same (offset, n,
 * channel) in the same order, on any stream behind the kernels that wrote the bucket.
 *   peer_bufs  : DEVICE array [world] of float*: the ranks' gradient buffers, peer-mapped (index = rank), same layout
 *   peer_flags : DEVICE array [world] of uint32_t*: the ranks' flag areas, mstcn_dp_flag_words() words each, zeroed once
 *   mc_buf     : multicast address of the same buffers, or NULL
 *   channel    : 0 .. 7; calls that may be in flight at the same time (one bucket per stage) use different channels
 * The kernel's CTAs (128 threads, no shared memory) fit beside a resident chain CTA, so a bucket's sum runs under the
 * next stage's backward.  A rank that never arrives traps after ~4 s instead of hanging the GPU. */
int mstcn_dp_allreduce(float* const* peer_bufs, uint32_t* const* peer_flags, float* mc_buf, int64_t offset, int64_t n,
                       int32_t rank, int32_t world, int32_t channel, void* stream);
int64_t mstcn_dp_flag_words(void);

/* profiling hook: when device_buf (>= 64 int64) is non-NULL, CTA 0 of every tensor-core layer kernel
 * records SM-clock timestamps of its first tile's pipeline phases in slots 0..31, and CTA 0 of every weight-gradient
 * launch its per-role wait / run clock totals in slots 32..43 (tools/wgrad_profile.py); NULL switches it off */
int mstcn_debug_tc_timing(int64_t* device_buf);

/* profiling hook: when device_buf (>= 8 * num_layers * B * ceil(T/128) int64) is non-NULL, every forward chain launch
 * records per task eight %globaltimer stamps (dependency poll start, dependencies satisfied, tap GEMM
 * complete, tile published, first TMA issued, centre tap landed, x_lo parked, 1x1 GEMM complete); NULL switches it off */
int mstcn_debug_chain_trace(int64_t* device_buf);

/* post-mortem hook: every bounded device-side wait (mbarrier waits, tile-flag polls) that runs out writes
 * [code (1 = mbarrier, 2 = tile flags), a, b, blockIdx.x, threadIdx.x] (a, b = barrier shared address and parity, or task
 * and CTA) to host_words before it traps.  host_words: >= 8 int64 of pinned (device-mapped) host memory, zeroed; it
 * survives the context the trap destroys.  NULL detaches. */
int mstcn_debug_trap_report(int64_t* host_words);

/* measurement hook (bench.py's per-kernel roofline entries): with enable != 0 every tensor-core mstcn_backward_stage
 * call drains its streams around the stage's backward chain launch (tc_layer_kernel<2>) and around its weight-gradient
 * launch (tc_wgrad_kernel), times each alone with CUDA events on the stream it is launched on, and keeps the two
 * durations of the latest call per stage.  mstcn_debug_backward_times copies [chain_ms, wgrad_ms] x num_stages (up to
 * n floats) to host memory `out`.  The step is serialised while this is on: never leave it enabled for a timed run. */
int mstcn_debug_backward_timing(int32_t enable);
int mstcn_debug_backward_times(float* out, int32_t n);

/* test hook: the {0,2} multiplier the kernels apply for (layer_id, frame n, channel c) -> (N,64) */
int mstcn_dropout_scale(const mstcn_dropout* drop, int32_t layer_id, int64_t n_frames, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* MSTCN_B200_H_ */
