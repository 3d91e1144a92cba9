"""Fused frame-wise cross-entropy: nn.CrossEntropyLoss(ignore_index=-1) forward + backward in one
kernel pass (train.py:12,266-267,326), over the (B*T, n_class) output of MultiStageModel."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _cabi
from ._cabi import check, ptr, stream_ptr


def ce_forward_backward(outputs, labels, n_valid=None):
    """One launch pair: returns (result, gout) with result = [mean loss, 1/n_valid, n_valid] on the device and
    gout = softmax - onehot (UNSCALED; multiply by result[1], or hand result[1:2] to the backward as its gscale)."""
    lib = _cabi.lib()
    if not outputs.is_cuda or outputs.dtype != torch.float32:
        raise RuntimeError("fused cross-entropy needs CUDA float32 logits (no CPU path)")
    if labels.dtype != torch.int64 or labels.device != outputs.device:
        raise RuntimeError("labels must be int64 on the logits' device")
    outputs = outputs.contiguous()
    labels = labels.contiguous()
    n, k = outputs.shape
    if labels.numel() != n:
        raise ValueError("labels must have one entry per logits row")
    gout = torch.empty_like(outputs)
    result = torch.empty(3, dtype=torch.float32, device=outputs.device)
    scratch = torch.empty(lib.mstcn_ce_scratch_floats(n), dtype=torch.float32, device=outputs.device)
    check(lib.mstcn_ce_loss(ptr(outputs), ptr(labels), n, k, int(n_valid or 0), ptr(gout), ptr(result),
                            ptr(scratch), stream_ptr()))
    return result, gout


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, labels, n_valid):
        result, gout = ce_forward_backward(outputs, labels, n_valid)
        ctx.gout, ctx.result = gout, result
        return result[0]

    @staticmethod
    def backward(ctx, gl):
        return ctx.gout * (ctx.result[1] * gl), None, None          # out of place: retain_graph backward stays correct


class FrameCrossEntropy(nn.Module):
    """Drop-in for nn.CrossEntropyLoss(ignore_index=-1) on MultiStageModel outputs.

    n_valid (optional) replaces the divisor: a data-parallel shard passes the GLOBAL number of
    valid frames so that summed gradients equal the single-process mean.
    Labels other than -1 outside [0, n_class) make the loss NaN (torch asserts on the device); a batch
    whose rows are all ignored gives NaN like torch's empty mean."""

    def __init__(self, ignore_index=-1):
        super().__init__()
        if ignore_index != -1:
            raise NotImplementedError("only ignore_index = -1 (train.py:12 _TARGET_PAD) is supported")
        self.ignore_index = ignore_index

    def forward(self, outputs, labels, n_valid=None):
        return _FusedCE.apply(outputs, labels, n_valid)


class _FusedPaperLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stage_logits, labels, lens_dev, B, T, n_valid, lam, tau):
        lib = _cabi.lib()
        if not stage_logits.is_cuda or stage_logits.dtype != torch.float32 or stage_logits.dim() != 3:
            raise RuntimeError("the fused MS-TCN loss needs CUDA float32 (S, B*T, n_class) stage logits (no CPU path)")
        if labels.dtype != torch.int64 or labels.device != stage_logits.device:
            raise RuntimeError("labels must be int64 on the logits' device")
        z = stage_logits.contiguous()
        S, n, k = z.shape
        if n != B * T or labels.numel() != n:
            raise ValueError("stage logits / labels do not match (B, T)")
        g = torch.empty_like(z)
        result = torch.empty(3, dtype=torch.float32, device=z.device)
        scratch = torch.empty(lib.mstcn_paper_loss_scratch_floats(S, n), dtype=torch.float32, device=z.device)
        check(lib.mstcn_paper_loss(ptr(z), ptr(labels.contiguous()), ptr(lens_dev), S, B, T, k, int(n_valid), float(lam),
                                   float(tau), ptr(g), ptr(result), ptr(scratch), stream_ptr()))
        ctx.g = g
        ctx.parts = result
        return result[0]

    @staticmethod
    def backward(ctx, gl):
        return ctx.g * gl, None, None, None, None, None, None, None


class MsTcnLoss(nn.Module):
    """Canonical MS-TCN training loss (Farha & Gall, CVPR 2019) on MultiStageModel.forward_stages(x, x_len):
    sum over stages of CrossEntropy(ignore_index=-1) + lam * truncated-MSE smoothing of the frame-to-frame
    log-probabilities (tau = 4, i.e. clamp at 16), one fused forward + backward kernel.

    The reference trains with plain CE on the max over stages instead (train.py:266-267,326), so
    this loss has no reference result to pin it: its oracle is oracle.ms_tcn_paper_loss and a torch autograd
    restatement in the tests ("parity unpinned")."""

    def __init__(self, lam=0.15, tau=4.0):
        super().__init__()
        self.lam, self.tau = float(lam), float(tau)
        self.last_parts = None            # device tensor [loss, CE part, T-MSE part] of the latest call

    def forward(self, stage_logits, labels, x_len, n_valid=None):
        B = len(x_len)
        T = stage_logits.shape[1] // B
        lens_dev = torch.tensor([int(v) for v in x_len], dtype=torch.int32, device=stage_logits.device)
        if n_valid is None:
            n_valid = int(sum(min(int(v), T) for v in x_len))      # labels are -1 exactly beyond x_len (pad_batch)
        loss = _FusedPaperLoss.apply(stage_logits, labels, lens_dev, B, T, n_valid, self.lam, self.tau)
        return loss
