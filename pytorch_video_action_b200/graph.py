"""CUDA-graph replay of the MS-TCN train step (forward -> CrossEntropy -> backward).

One step is ~300 kernel launches on several internal streams; issued from Python/C++ they cost about as
much host time as the GPU needs to run them.  For a fixed batch shape the whole step is captured once and
replayed with a single launch.  Dropout still draws a fresh mask on every replay: the Philox offset lives
in a device counter (mstcn_dropout.offset_dev) that the captured step increments.
"""
from __future__ import annotations

import torch

from . import _cabi
from .loss import FrameCrossEntropy, ce_forward_backward


# Replays of captured steps hold chain launches (CTAs spinning on other CTAs' tile flags), so two of them must not
# share a GPU: replays issued on different streams of one device are ordered through an event (same idea as the chain
# lane inside libmstcn_b200.so, which cannot see into a graph launch).
_last_replay = {}          # device index -> (event, stream id)


def _ordered_replay(graph, device):
    cur = torch.cuda.current_stream(device)
    prev = _last_replay.get(device.index)
    if prev is not None and prev[1] != cur.cuda_stream:
        cur.wait_event(prev[0])
    graph.replay()
    ev = prev[0] if prev is not None else torch.cuda.Event()
    ev.record(cur)
    _last_replay[device.index] = (ev, cur.cuda_stream)


class GraphedTrainStep:
    """step = GraphedTrainStep(net, criterion, x_len, example_x, example_y); loss = step(x, y)

    x: (B, T, dim) float32 and y: (B*T,) int64 CUDA tensors of the captured shape; `x_len` is fixed.
    After the call every parameter's .grad holds this batch's gradients (as after loss.backward()); the
    returned loss is a 0-dim CUDA tensor that the next call overwrites.  `n_valid` is forwarded to the
    criterion (data-parallel shards pass the global valid-frame count).  Call optimizer.step() yourself.
    """

    def __init__(self, net, criterion, x_len, example_x, example_y, n_valid=None, warmup=3, dp=None, inputs=None,
                 optimizer=None):
        """inputs (optional): list of (x, y) CUDA tensor pairs the caller keeps refilling in place (e.g. the two halves
        of an H2D double buffer).  One graph is captured per pair, reading the pair directly; `step.replay(i)` then runs
        a step on pair i without the device-to-device copy into the static buffers that `step(x, y)` needs.
        optimizer (optional): a FusedAdam; its step then runs INSIDE the captured graph (step count and learning rate
        on the device: FusedAdam.step_capturable), i.e. one replay = zero_grad -> forward -> loss -> backward ->
        (gradient all-reduce) -> optimizer.step() (train.py:305-329).  lr_scheduler changes are picked up at the next
        replay.  Parameters, moments and the step count are restored after the warm-up steps."""
        self.net, self.criterion, self.x_len, self.n_valid, self.dp = net, criterion, list(x_len), n_valid, dp
        self.optimizer = optimizer
        if optimizer is not None and not hasattr(optimizer, "step_capturable"):
            raise TypeError("GraphedTrainStep(optimizer=...) needs a pytorch_video_action_b200.FusedAdam")
        self.static_x = example_x.clone()
        self.static_y = example_y.clone()
        net._ensure_flat()
        net._drop_counter = torch.zeros(1, dtype=torch.int64, device=example_x.device)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        self._slots = []
        with torch.cuda.stream(side):
            self._x, self._y = self.static_x, self.static_y
            snap = optimizer.snapshot() if optimizer is not None else None
            for _ in range(warmup):
                self._step()
            torch.cuda.synchronize()
            if snap is not None:
                optimizer.restore(snap)
                torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.static_loss = self._step()
            for x, y in (inputs or []):
                if x.shape != self.static_x.shape or y.shape != self.static_y.shape:
                    raise ValueError("every input pair must have the captured batch shape")
                self._x, self._y = x, y
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    loss = self._step()
                self._slots.append((gr, loss))
            self._x, self._y = self.static_x, self.static_y
        torch.cuda.current_stream().wait_stream(side)

    def replay(self, i):
        """One step on inputs[i] as they are now (no copy).  Returns the 0-dim device loss of that slot."""
        gr, loss = self._slots[i]
        if self.optimizer is not None:
            self.optimizer.push_lr()
        _ordered_replay(gr, self.static_x.device)
        return loss

    def _step(self):
        net = self.net
        for p in net.parameters():
            p.grad = None
        if self.dp is not None and not isinstance(self.criterion, FrameCrossEntropy):
            loss = self.dp.forward_backward(self._x, self.x_len, self._y, self.n_valid)
        elif isinstance(self.criterion, FrameCrossEntropy):
            # the fused criterion: forward -> CE -> backward called directly (no autograd nodes, and the 1/n_valid
            # scale goes to the backward kernels as a device scalar instead of an elementwise pass over the gradient)
            loss = self._step_direct()
        else:
            loss = self.criterion(net._forward_impl(self._x, self.x_len, strict_len=False), self._y,
                                  n_valid=self.n_valid)
            loss.backward()
        net._drop_counter.add_(1)
        if self.optimizer is not None:
            self.optimizer.step_capturable()
        return loss.detach()

    def _step_direct(self):
        import ctypes as C
        net, x = self.net, self._x
        B, T = net._check_input(x, self.x_len, False)
        net._ensure_flat()
        lens_dev = net._lens_device(self.x_len, x.device)
        net._lens_host = (C.c_int32 * B)(*[int(v) for v in self.x_len])
        drop = net._next_dropout()
        hook = self.dp.reducer.on_stage_done if self.dp is not None else None     # per-stage gradient all-reduce
        fused_head = net.tensor_cores and self.n_valid is not None and not (net._dims.flags & _cabi.FLAG_FFMA_BACKWARD)
        if fused_head:
            # max over stages + CE + their backward in one kernel, straight into the backward's gradient planes
            lib = _cabi.lib()
            _, _, ws = net._launch_forward(x, lens_dev, B, T, drop, training=True, want_out=False)
            result = torch.empty(3, dtype=torch.float32, device=x.device)
            scratch = torch.empty(lib.mstcn_ce_scratch_floats(B * T), dtype=torch.float32, device=x.device)
            _cabi.check(lib.mstcn_loss_head(C.byref(net._dims), _cabi.ptr(ws), B, T, _cabi.ptr(self._y), int(self.n_valid), None,
                                            None, _cabi.ptr(result), _cabi.ptr(scratch), _cabi.stream_ptr()))
            net._launch_backward(x, lens_dev, B, T, drop, ws, None, None, stage_hook=hook)
        else:
            out, winner, ws = net._launch_forward(x, lens_dev, B, T, drop, training=True)
            result, gout = ce_forward_backward(out, self._y, self.n_valid)
            net._launch_backward(x, lens_dev, B, T, drop, ws, winner, gout, gscale=result[1:2], stage_hook=hook)
        if self.dp is not None:
            self.dp.reducer.finish()
        net._release_workspace(ws)
        return result[0]          # data parallel: this rank's contribution (sum over ranks = the global mean loss)

    def __call__(self, x, y):
        if x.shape != self.static_x.shape or y.shape != self.static_y.shape:
            raise ValueError("GraphedTrainStep was captured for a different batch shape")
        self.static_x.copy_(x, non_blocking=True)
        self.static_y.copy_(y, non_blocking=True)
        if self.optimizer is not None:
            self.optimizer.push_lr()
        _ordered_replay(self.graph, self.static_x.device)
        return self.static_loss
