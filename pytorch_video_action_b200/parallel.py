"""Data-parallel MS-TCN training: shard by video, one process per GPU, bucketed gradient
all-reduce overlapped with backward.  The reference has no distributed code
(single cuda:0, train.py:181); this is the north-star addition.

Every op of the model is per-video, so the only exchange per step is the gradient sum.  The
loss divisor is the GLOBAL valid-frame count (known on the host from the length list), so
ranks SUM gradients -- the result equals the single-process mean of train.py:267,326.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_videos(lens, world_size):
    """Deal videos round-robin by descending length so every rank gets ~equal frames
    (BucketBatchSampler sorts by length too, data_utils.py:24).  Returns world_size index lists."""
    order = sorted(range(len(lens)), key=lambda i: (-int(lens[i]), i))
    shards = [[] for _ in range(world_size)]
    for pos, i in enumerate(order):
        lap, r = divmod(pos, world_size)
        shards[r if lap % 2 == 0 else world_size - 1 - r].append(i)     # snake order balances sums
    return shards


def local_pad_length(local_lens, global_T):
    """Padded length a shard must use so results match the un-sharded batch.

    The stage-input 1x1 conv is not masked (networks.py:330), so a video shorter than the global
    padded length sees the bias on its first padded frame through layer 0's +1 tap (SURVEY.md
    fact 0.5).  One padded frame reproduces that exactly; a video as long as the global batch's
    longest needs none."""
    m = max(int(v) for v in local_lens)
    return m if m >= global_T else m + 1


class GradBucketReducer:
    """Sum-all-reduce of the flat gradient buffer in buckets, issued as each bucket becomes final.

    boundaries = model.bucket_boundaries() = [0, b1, ..., total]; bucket i is [boundaries[i],
    boundaries[i+1]) = stage i-1's dilated layers and class head plus stage i's input projection.
    A stage's weight gradients are computed by one kernel that runs under the NEXT stage's chain, so
    once the backward call for stage s has been issued the gradients of every stage > s are final:
    the call for stage s releases bucket s+2, the last call (stage 0) also buckets 1 and 0.  With the
    NCCL backend each all_reduce runs on the process group's own stream, so it overlaps the backward
    kernels of the earlier stages still being launched."""

    def __init__(self, flat_grads, boundaries, group=None, overlap=False):
        """overlap=False (default): one all-reduce of the whole buffer after the last stage instead of a bucket per
        stage.  The NCCL kernels need SMs of their own, and a chain launch whose CTAs cannot all become resident stalls
        until they are free, so at small step times the single late all-reduce is the faster choice.
        flat_grads: the flat gradient buffer, or a zero-argument callable returning it (re-read on every step, so a
        model whose buffers were re-created by .to() is followed)."""
        self._flat_src = flat_grads if callable(flat_grads) else (lambda: flat_grads)
        self.bounds = list(boundaries) if overlap else [boundaries[0], boundaries[-1]]
        self.n_stage_buckets = len(boundaries) - 1
        self.overlap = overlap
        self.group = group
        self.pending = []
        self.issued = []

    @property
    def flat(self):
        return self._flat_src()

    def bucket(self, i):
        return self.flat[self.bounds[i]: self.bounds[i + 1]]

    def reduce_bucket(self, i):
        self.issued.append(i)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            self.pending.append(dist.all_reduce(self.bucket(i), op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def on_stage_done(self, s):
        """Hook for MultiStageModel backward: called right after stage s's kernels are enqueued."""
        if not self.overlap:
            if s == 0:
                self.reduce_bucket(0)
            return
        n_buckets = len(self.bounds) - 1
        if s + 2 < n_buckets:
            self.reduce_bucket(s + 2)
        if s == 0:
            if n_buckets > 1:
                self.reduce_bucket(1)
            self.reduce_bucket(0)

    def finish(self):
        for w in self.pending:
            w.wait()
        n_buckets = len(self.bounds) - 1
        ok = sorted(self.issued) == list(range(n_buckets))
        self.pending, self.issued = [], []
        if not ok:
            raise RuntimeError("gradient buckets were not all reduced exactly once")


class PeerGradReducer:
    """The same bucket protocol as GradBucketReducer, summed by `mstcn_dp_allreduce` over NVLink / NVSwitch peer memory
    instead of NCCL: the model's flat gradient buffer is re-homed into a symmetric allocation (every rank maps every
    peer's buffer, plus an NVLS multicast mapping where the fabric has one), and each bucket is summed in place by one
    small kernel whose CTAs fit beside the resident backward-chain CTAs -- the sum of stage s+1's gradients runs UNDER
    stage s's backward instead of after it, and takes no SMs away from the chain (the reason the NCCL overlap lost).
    `torch.distributed` / NCCL stay the plumbing (rendezvous, the parameter broadcast, barriers)."""

    def __init__(self, net, boundaries, group=None, overlap=False, nvls=None):
        """overlap=False (default, measured on 8 B200s at config 2: 1.301 ms/step against 1.322 with per-stage buckets and
        1.345 with NCCL): ONE launch over the whole buffer behind the backward; True: a bucket per stage on a side
        stream under the remaining backward (the extra cross-rank rendezvous cost more than the hidden microseconds).
        nvls: None = through the NVSwitch multicast mapping when the fabric offers one, False = one-shot peer reads."""
        import torch.distributed._symmetric_memory as symm
        from . import _cabi
        self._cabi = _cabi
        self.net, self.group, self.overlap = net, group, overlap
        pg = group if group is not None else dist.group.WORLD
        _, gflat = net.flat_parameters()
        dev = gflat.device
        self.buf = symm.empty(gflat.numel(), dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, pg)
        words = int(_cabi.lib().mstcn_dp_flag_words())
        self.flagbuf = symm.empty(words, dtype=torch.int32, device=dev)
        self.flagbuf.zero_()
        self.fhdl = symm.rendezvous(self.flagbuf, pg)
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        self.mc = int(self.hdl.multicast_ptr) if (nvls is not False and self.hdl.multicast_ptr) else 0
        if nvls is True and not self.mc:
            raise RuntimeError("PeerGradReducer(nvls=True): no multicast mapping on this fabric")
        self.bounds = list(boundaries)
        self.n_stage_buckets = len(self.bounds) - 1
        if self.n_stage_buckets > 8:
            raise NotImplementedError("PeerGradReducer: at most 8 gradient buckets (7 stages)")
        net.rebind_grad_buffer(self.buf)              # every .grad now aliases the peer-visible buffer
        self.comm = torch.cuda.Stream(device=dev)
        self.issued, self._forked = [], False
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                     # every rank's flag area is zeroed before anyone signals

    @property
    def flat(self):
        g = self.net.flat_parameters()[1]
        if g.data_ptr() != self.buf.data_ptr():
            raise RuntimeError("the model's gradient buffer was re-created (.to() after wrapping): rebuild the data-parallel wrapper")
        return g

    def _reduce_range(self, lo_bucket, hi_bucket):
        """sum buckets [lo_bucket, hi_bucket) as one launch on the communication stream, behind everything enqueued on the
        caller's stream so far"""
        _ = self.flat
        lo, hi = self.bounds[lo_bucket], self.bounds[hi_bucket]
        self.issued += list(range(lo_bucket, hi_bucket))
        if os.environ.get("MSTCN_DP_SKIP") == "1":       # measurement only: the step without its gradient sum
            return
        c = self._cabi
        args = (int(self.hdl.buffer_ptrs_dev), int(self.fhdl.buffer_ptrs_dev), self.mc or None, lo, hi - lo, self.rank,
                self.world, lo_bucket)
        if not self.overlap:                             # behind the backward, on the caller's stream
            c.check(c.lib().mstcn_dp_allreduce(*args, c.stream_ptr()))
            return
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        self._forked = True
        with torch.cuda.stream(self.comm):
            c.check(c.lib().mstcn_dp_allreduce(*args, c.stream_ptr()))

    def on_stage_done(self, s):
        n = self.n_stage_buckets
        if not self.overlap:
            if s == 0:
                self._reduce_range(0, n)
            return
        if s + 2 < n:
            self._reduce_range(s + 2, s + 3)
        if s == 0:
            self._reduce_range(0, min(2, n))          # the last two buckets become final together: one launch

    def finish(self):
        if self._forked:
            torch.cuda.current_stream().wait_stream(self.comm)
            self._forked = False
        ok = sorted(self.issued) == list(range(self.n_stage_buckets))
        self.issued = []
        if not ok:
            raise RuntimeError("gradient buckets were not all reduced exactly once")


class DataParallelMSTCN:
    """Thin trainer-side wrapper: net(x, x_len) -> FrameCrossEntropy(n_valid=global) -> backward with
    the bucket hook -> finish.  `net` is a pytorch_video_action_b200.MultiStageModel on this rank's GPU."""

    def __init__(self, net, criterion, group=None, overlap=None, allreduce="auto", nvls=None):
        """allreduce: "peer" = mstcn_dp_allreduce over peer-mapped gradient buffers (PeerGradReducer), "nccl" =
        ncclAllReduce (GradBucketReducer), "auto" = peer when every rank can set it up, else nccl.  nvls: None = NVSwitch
        multicast (multimem) when available, False = one-shot peer reads, True = required.  overlap: None / False = one
        sum behind the backward (fastest measured for both paths), True = per-stage buckets under the backward."""
        self.net, self.criterion, self.group = net, criterion, group
        flat, _ = net.flat_parameters()
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        self.reducer, self.allreduce, self.fallback_reason = None, "nccl", None
        if multi and allreduce in ("peer", "auto"):
            ok = 1
            try:
                self.reducer = PeerGradReducer(net, net.bucket_boundaries(), group, overlap=bool(overlap), nvls=nvls)
            except Exception as e:                 # noqa: BLE001 -- no peer mapping on this box / torch build
                ok, self.fallback_reason = 0, f"{type(e).__name__}: {e}"
                if allreduce == "peer":
                    raise
            t = torch.tensor([ok], device=flat.device, dtype=torch.int32)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            if int(t) == 1:
                self.allreduce = "peer-nvls" if self.reducer.mc else "peer"
            else:
                self.reducer = None
                net._ensure_flat_private_grads()
        if self.reducer is None:
            # the reducer always sums the model's CURRENT flat gradient buffer (net.flat_parameters() re-flattens after .to())
            self.reducer = GradBucketReducer(lambda: net.flat_parameters()[1], net.bucket_boundaries(), group,
                                             overlap=bool(overlap))
        if multi:
            dist.broadcast(flat, src=0, group=group)      # identical replicas

    def forward_backward(self, x, x_len, labels, n_valid_global):
        """One local micro-step.  x may be padded to local_pad_length(...) > max(x_len).

        The all-reduce sums the flat gradient buffer, so this step's LOCAL gradients must be the only thing in it:
        gradients already present (accumulation over micro-steps, a missing zero_grad, foreign .grad tensors) are set
        aside first and added back after the reduction -- reducing them a second time would count them world_size
        times."""
        params = self.net._params_in_order()
        saved = None
        if any(p.grad is not None for p in params):
            saved = [None if p.grad is None else p.grad.detach().clone() for p in params]
            for p in params:
                p.grad = None
        self.net._stage_hook = self.reducer.on_stage_done
        try:
            out = self.net._forward_impl(x, x_len, strict_len=False)
            loss = self.criterion(out, labels, n_valid=n_valid_global)
            loss.backward()
        finally:
            self.net._stage_hook = None
        gflat = self.net.flat_parameters()[1]
        if any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(params, self.net._gviews)):
            raise RuntimeError("data-parallel step: the backward did not write the model's flat gradient buffer")
        assert self.reducer.flat.data_ptr() == gflat.data_ptr()
        self.reducer.finish()
        if saved is not None:
            for p, g in zip(params, saved):
                if g is not None:
                    p.grad.add_(g)
        return loss.detach()      # local contribution: sum over ranks = global mean loss
