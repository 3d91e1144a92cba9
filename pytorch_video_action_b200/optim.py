"""Fused Adam over the model's flat parameter / gradient buffers: one kernel instead of the
foreach launches of torch.optim.Adam over 176 tensors (train.py:273,305,329)."""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    """Drop-in for `optim.Adam(net.parameters(), lr, betas=(0.9, 0.999), eps=1e-8)` (train.py:273): a real
    torch.optim.Optimizer (one param group over the model's parameters), so `lr_scheduler.StepLR(optimizer, ...)`
    (train.py:274,334-335), `zero_grad()`, `state_dict()` / `load_state_dict()` work as with torch's Adam.
    torch.optim.Adam semantics without weight decay / amsgrad (the reference uses neither), eps outside the sqrt.

    The moments live in two flat buffers parallel to the model's flat parameter buffer; `state[p]` exposes them per
    parameter as views (`exp_avg`, `exp_avg_sq`) plus the shared `step`, which is also the checkpoint layout of
    torch.optim.Adam -- a state_dict saved by either optimizer loads into the other."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if isinstance(model, torch.nn.Module) and hasattr(model, "flat_parameters"):
            self.model = model
        else:
            raise TypeError("FusedAdam(model, ...): pass the pytorch_video_action_b200.MultiStageModel itself "
                            "(the optimizer steps its flat parameter buffer), not model.parameters()")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        self._flat_ptr = None
        self._exp_avg = None
        self._exp_avg_sq = None
        self._dev_step, self._dev_lr, self._dev_lr_host, self._dev_step_seen = None, None, None, 0
        # the full key set of torch.optim.Adam's param group, so a checkpoint moves between the two optimizers unchanged
        defaults = dict(torch.optim.Adam([torch.zeros(1)]).defaults)
        defaults.update(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps), weight_decay=0,
                        amsgrad=False, maximize=False)
        super().__init__(list(model.parameters()), defaults)

    # -- the flat moment buffers follow the model's flat parameter buffer (re-created by .to() / dtype rebinding) --
    def _ensure_state(self):
        flat, _ = self.model.flat_parameters()
        if self._flat_ptr == flat.data_ptr() and self._exp_avg is not None:
            return flat
        old = [(self.state[p].get("exp_avg"), self.state[p].get("exp_avg_sq"), self.state[p].get("step"))
               for p in self.model._params_in_order()]
        self._exp_avg = torch.zeros_like(flat)
        self._exp_avg_sq = torch.zeros_like(flat)
        step = 0
        for p, off, (m, v, st) in zip(self.model._params_in_order(), self.model.grad_offsets(), old):
            mv = self._exp_avg[off: off + p.numel()].view(p.shape)
            vv = self._exp_avg_sq[off: off + p.numel()].view(p.shape)
            if m is not None:
                mv.copy_(m)
                vv.copy_(v)
                step = max(step, int(st))
            self.state[p] = {"step": torch.tensor(float(step)), "exp_avg": mv, "exp_avg_sq": vv}
        for p in self.model._params_in_order():
            self.state[p]["step"] = torch.tensor(float(step))
        self._flat_ptr = flat.data_ptr()
        return flat

    @property
    def step_count(self):
        self._pull_device_steps()
        params = self.model._params_in_order()
        st = self.state.get(params[0], {}).get("step") if params else None
        return 0 if st is None else int(st)

    # -- graph mode (GraphedTrainStep(optimizer=...)): step count and learning rate live on the device -----------------
    def _pull_device_steps(self):
        """Fold the steps taken by captured replays (device counter) back into state[p]["step"]."""
        if self._dev_step is None:
            return
        done = int(self._dev_step.item())
        if done != self._dev_step_seen:
            self._dev_step_seen = done
            st = torch.tensor(float(done))
            for p in self.model._params_in_order():
                if p in self.state:
                    self.state[p]["step"] = st

    def snapshot(self):
        """(params, moments, step count) copies -- GraphedTrainStep restores them after its warm-up steps"""
        flat = self._ensure_state()
        return flat.clone(), self._exp_avg.clone(), self._exp_avg_sq.clone(), self.step_count

    def restore(self, snap):
        flat = self._ensure_state()
        flat.copy_(snap[0]); self._exp_avg.copy_(snap[1]); self._exp_avg_sq.copy_(snap[2])
        st = torch.tensor(float(snap[3]))
        for p in self.model._params_in_order():
            self.state[p]["step"] = st
        if self._dev_step is not None:
            self._dev_step.fill_(snap[3])
            self._dev_step_seen = snap[3]

    def push_lr(self):
        """Copy param_groups[0]["lr"] to the device scalar the captured step reads (stream-ordered; only when it changed)."""
        lr = float(self.param_groups[0]["lr"])
        if self._dev_lr is not None and lr != self._dev_lr_host:
            self._dev_lr.fill_(lr)
            self._dev_lr_host = lr

    @torch.no_grad()
    def step_capturable(self):
        """optimizer.step() with device-side state: safe to capture in a CUDA graph and to replay.  Between replays call
        push_lr() (GraphedTrainStep does) so that lr_scheduler changes reach the device."""
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdam keeps one param group (the model's flat buffer)")
        group = self.param_groups[0]
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise NotImplementedError("FusedAdam implements plain Adam (no weight decay / amsgrad / maximize), as train.py:273 uses it")
        flat = self._ensure_state()
        _, gflat = self.model.flat_parameters()
        if self._dev_step is None:
            self._dev_step = torch.zeros(1, dtype=torch.int64, device=flat.device)
            self._dev_lr = torch.zeros(1, dtype=torch.float32, device=flat.device)
            self._dev_lr_host = None
        if not torch.cuda.is_current_stream_capturing():
            # (re)seed the device counter from the host-side state outside a capture
            params = self.model._params_in_order()
            st = self.state.get(params[0], {}).get("step")
            host_steps = 0 if st is None else int(st)
            if host_steps != self._dev_step_seen:
                self._dev_step.fill_(host_steps)
                self._dev_step_seen = host_steps
            self.push_lr()
        check(_cabi.lib().mstcn_adam_step_dev(ptr(flat), ptr(gflat), ptr(self._exp_avg), ptr(self._exp_avg_sq), flat.numel(),
                                              ptr(self._dev_lr), float(group["betas"][0]), float(group["betas"][1]),
                                              float(group["eps"]), ptr(self._dev_step), stream_ptr()))

    def zero_grad(self, set_to_none=True):
        if set_to_none:
            for p in self.model.parameters():
                p.grad = None
        else:
            _, g = self.model.flat_parameters()
            g.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdam keeps one param group (the model's flat buffer)")
        flat = self._ensure_state()
        _, gflat = self.model.flat_parameters()
        params = self.model._params_in_order()
        if any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(params, self.model._gviews)):
            raise RuntimeError("FusedAdam needs the gradients produced by MultiStageModel's backward")
        group = self.param_groups[0]
        if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
            raise NotImplementedError("FusedAdam implements plain Adam (no weight decay / amsgrad / maximize), as train.py:273 uses it")
        step = self.step_count + 1            # (includes the steps captured replays have taken)
        check(_cabi.lib().mstcn_adam_step(ptr(flat), ptr(gflat), ptr(self._exp_avg), ptr(self._exp_avg_sq), flat.numel(),
                                          float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                                          float(group["eps"]), step, stream_ptr()))
        st = torch.tensor(float(step))
        for p in params:
            self.state[p]["step"] = st
        if self._dev_step is not None:
            self._dev_step.fill_(step)
            self._dev_step_seen = step
        return loss

    def state_dict(self):
        self._pull_device_steps()
        # every parameter gets its OWN `step` tensor in the checkpoint: torch.optim.Adam's foreach path increments the
        # step tensors in place, so one tensor shared by all parameters would advance 176 steps per step after a load
        sd = super().state_dict()
        sd["state"] = {k: {**v, "step": v["step"].clone()} if "step" in v else dict(v) for k, v in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the loaded per-parameter moments are fresh tensors: fold them back into the flat buffers
        self._flat_ptr = None
        self._ensure_state()
        if self._dev_step is not None:
            steps = int(self.state[self.model._params_in_order()[0]]["step"])
            self._dev_step.fill_(steps)
            self._dev_step_seen = steps
            self._dev_lr_host = None
            self.push_lr()
