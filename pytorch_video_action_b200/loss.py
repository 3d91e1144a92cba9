"""Fused frame-wise cross-entropy: nn.CrossEntropyLoss(ignore_index=-1) forward + backward in one
kernel pass (train.py:12,266-267,326), over the (B*T, n_class) output of MultiStageModel."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _cabi
from ._cabi import check, ptr, stream_ptr


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, labels, n_valid):
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
        ctx.gout, ctx.result = gout, result
        return result[0]

    @staticmethod
    def backward(ctx, gl):
        return ctx.gout.mul_(ctx.result[1] * gl), None, None


class FrameCrossEntropy(nn.Module):
    """Drop-in for nn.CrossEntropyLoss(ignore_index=-1) on MultiStageModel outputs.

    n_valid (optional) replaces the divisor: a data-parallel shard passes the GLOBAL number of
    valid frames so that summed gradients equal the single-process mean."""

    def __init__(self, ignore_index=-1):
        super().__init__()
        if ignore_index >= 0:
            raise NotImplementedError("only negative ignore_index (the reference uses -1) is supported")
        self.ignore_index = ignore_index

    def forward(self, outputs, labels, n_valid=None):
        return _FusedCE.apply(outputs, labels, n_valid)
