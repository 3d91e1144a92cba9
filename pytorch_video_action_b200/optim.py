"""Fused Adam over the model's flat parameter / gradient buffers: one kernel instead of the
foreach launches of torch.optim.Adam over 176 tensors (train.py:273,305,329)."""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps) semantics (no weight decay / amsgrad), eps outside the sqrt."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self.param_groups = [{"lr": self.lr}]      # what StepLR (train.py:274) touches

    def zero_grad(self, set_to_none=True):
        if set_to_none:
            for p in self.model.parameters():
                p.grad = None
        else:
            _, g = self.model.flat_parameters()
            g.zero_()

    def step(self):
        flat, gflat = self.model.flat_parameters()
        params = self.model._params_in_order()
        if any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(params, self.model._gviews)):
            raise RuntimeError("FusedAdam needs the gradients produced by MultiStageModel's backward")
        if self.exp_avg is None or self.exp_avg.device != flat.device:
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count += 1
        check(_cabi.lib().mstcn_adam_step(ptr(flat), ptr(gflat), ptr(self.exp_avg), ptr(self.exp_avg_sq), flat.numel(),
                                          float(self.param_groups[0]["lr"]), self.betas[0], self.betas[1], self.eps,
                                          self.step_count, stream_ptr()))
