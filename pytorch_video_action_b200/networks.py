"""Drop-in MS-TCN model class: same constructor, forward(x, x_len) signature, state_dict keys
and same-seed initial values as the reference's networks.MultiStageModel
(/root/reference/networks.py:298-347), with all arithmetic done by the hand-written sm_100a
kernels behind the C ABI in include/mstcn_b200.h.

The nn.Conv1d / nn.Dropout sub-modules below are parameter holders only (they reproduce the
176 state_dict keys, shapes and default init of the reference); their own
forward() is never called.  There is no CPU path and no PyTorch fallback: a non-CUDA or
non-fp32 input raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _cabi
from ._cabi import MstcnDims, MstcnDropout, check, ptr, stream_ptr


class DilatedResidualLayer(nn.Module):
    """Parameter holder mirroring networks.py:336-341."""

    def __init__(self, dilation, in_channels, out_channels):
        super().__init__()
        self.conv_dilated = nn.Conv1d(in_channels, out_channels, 3, padding=dilation, dilation=dilation)
        self.conv_1x1 = nn.Conv1d(out_channels, out_channels, 1)
        self.dropout = nn.Dropout()


class SingleStageModel(nn.Module):
    """Parameter holder mirroring networks.py:322-327 (same construction order -> same init)."""

    def __init__(self, num_layers, num_f_maps, dim, n_class):
        super().__init__()
        self.conv_1x1 = nn.Conv1d(dim, num_f_maps, 1)
        self.layers = nn.ModuleList([DilatedResidualLayer(2 ** i, num_f_maps, num_f_maps) for i in range(num_layers)])
        self.conv_out = nn.Conv1d(num_f_maps, n_class, 1)


class _MstcnFunction(torch.autograd.Function):
    """One autograd node for the whole model: forward = mstcn_forward, backward = mstcn_backward."""

    @staticmethod
    def forward(ctx, x, anchor, model, lens_dev, drop, per_stage=False):
        B, T, _ = x.shape
        out, winner, ws = model._launch_forward(x, lens_dev, B, T, drop, training=True)
        ctx.model, ctx.x, ctx.lens_dev, ctx.drop = model, x, lens_dev, drop
        ctx.lens_host = model._lens_host
        ctx.ws, ctx.winner, ctx.BT = ws, (None if per_stage else winner), (B, T)
        if per_stage:                       # (S, B*T, K) per-stage masked logits; backward takes dLoss/dz_s, no max routing
            return model.stage_logits()
        return out

    @staticmethod
    def backward(ctx, gout):
        model = ctx.model
        B, T = ctx.BT
        if ctx.ws is None:      # the saved planes went back to the workspace pool with the first backward
            raise RuntimeError("MultiStageModel: backward through the same forward a second time "
                               "(the activation workspace is released after the first backward; run the forward again)")
        model._lens_host = ctx.lens_host
        model._launch_backward(ctx.x, ctx.lens_dev, B, T, ctx.drop, ctx.ws, ctx.winner, gout,
                               stage_hook=model._stage_hook)
        model._release_workspace(ctx.ws)
        ctx.ws = None
        return None, None, None, None, None, None


class MultiStageModel(nn.Module):
    """networks.MultiStageModel (networks.py:298-320), B200-native.

    forward(x, x_len): x float32 (B, T, dim) on a CUDA device, x_len a host list[int] with
    len == B and max == T.  Returns float32 (B*T, n_class) = max over stages of the per-stage
    logits, row b*T+t, differentiable w.r.t. the parameters.
    """

    def __init__(self, dim=400, num_stages=4, num_layers=20, num_f_maps=64, n_class=2):
        super().__init__()
        if num_f_maps != 64:
            raise NotImplementedError("the sm_100a kernels are specialised for num_f_maps == 64")
        if not 1 <= n_class <= 64:
            raise NotImplementedError("the sm_100a kernels support 1 <= n_class <= 64")
        if dim < 4 or dim % 4:
            raise NotImplementedError("dim must be a positive multiple of 4 (float4 feature loads)")
        self.stage1 = SingleStageModel(num_layers, num_f_maps, dim, n_class)
        self.stages = nn.ModuleList(
            [SingleStageModel(num_layers, num_f_maps, n_class, n_class) for _ in range(num_stages - 1)])
        self.n_class = n_class
        self._dims = MstcnDims(dim, num_stages, num_layers, num_f_maps, n_class,
                               _cabi.FLAG_TENSOR_CORES | _cabi.FLAG_PACK_TC_ONLY)
        self._flat = None            # flat parameter buffer the nn.Parameters alias
        self._gflat = None           # flat gradient buffer the .grad tensors alias
        self._packed = None
        self._views = []
        self._gviews = []
        self._plist = []
        self._anchor = None
        self._ws_pool = {}
        self._lens_cache = {}
        self._drop_seed = None
        self._drop_offset = 0
        self._drop_counter = None    # optional device int64 step counter (graph.GraphedTrainStep)
        self.last_workspace = None   # (tensor, B, T) of the latest training forward (for stage_logits)
        self._stage_hook = None      # set by parallel.DataParallelMSTCN: called after each backward stage
        self._lens_host = None       # ctypes int32 array of the current batch's lengths (video-group planning)
        # fp32 FFMA path only (tensor_cores = False): the forward cuts the batch into video groups whose kernel chains run
        # on concurrent streams.  The tensor-core path runs every stage's layers as one chain launch on the caller's stream
        # (two chain launches must never share the GPU) and ignores both settings.
        self.stream_groups = 4
        self.backward_stream_groups = 1

    @property
    def tensor_cores(self):
        """True: the dilated residual layers run on the tcgen05 tensor cores with error-compensated
        3xTF32 (fp32-equivalent); False: the exact fp32 FFMA kernels."""
        return bool(self._dims.flags & _cabi.FLAG_TENSOR_CORES)

    @tensor_cores.setter
    def tensor_cores(self, on):
        self._dims.flags = (self._dims.flags | _cabi.FLAG_TENSOR_CORES) if on else (self._dims.flags & ~_cabi.FLAG_TENSOR_CORES)

    @property
    def pack_ffma_operands(self):
        """False (default): with tensor cores on, a forward refreshes only the biases and the tensor-core operand
        images; True: also the transposed fp32 operands of the FFMA kernels (tests / tools that call those kernels
        through the C ABI with this model's packed buffer)."""
        return not (self._dims.flags & _cabi.FLAG_PACK_TC_ONLY)

    @pack_ffma_operands.setter
    def pack_ffma_operands(self, on):
        self._dims.flags = (self._dims.flags & ~_cabi.FLAG_PACK_TC_ONLY) if on else (self._dims.flags | _cabi.FLAG_PACK_TC_ONLY)

    # ------------------------------------------------------------------ parameters
    def _params_in_order(self):
        return self._plist if self._plist else [p for _, p in self.named_parameters()]

    def _ensure_flat(self):
        if self._flat is not None:
            pl, vs = self._plist, self._views
            mid = len(pl) // 2
            # .to()/.cuda()/.float() rebind every param.data at once, so three probes suffice
            if (pl[0].data_ptr() == vs[0].data_ptr() and pl[mid].data_ptr() == vs[mid].data_ptr()
                    and pl[-1].data_ptr() == vs[-1].data_ptr()):
                return
        params = [p for _, p in self.named_parameters()]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("MultiStageModel (B200-native) has no CPU path: call .to('cuda') first")
        lib = _cabi.lib()
        n = lib.mstcn_param_count(C.byref(self._dims))
        if n < 0:
            raise _cabi.MstcnError(lib.mstcn_last_error().decode())
        if lib.mstcn_param_tensors(C.byref(self._dims)) != len(params):
            raise RuntimeError("parameter list does not match the C layout")
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        gflat = torch.zeros(n, dtype=torch.float32, device=dev)
        views, gviews = [], []
        for i, p in enumerate(params):
            if p.dtype != torch.float32:
                raise RuntimeError("MultiStageModel (B200-native) is fp32 only")
            off = lib.mstcn_param_offset(C.byref(self._dims), i)
            v = flat[off: off + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            views.append(v)
            gviews.append(gflat[off: off + p.numel()].view(p.shape))
        assert off + params[-1].numel() == n
        self._flat, self._gflat, self._views, self._gviews, self._plist = flat, gflat, views, gviews, params
        self._packed = torch.empty(lib.mstcn_packed_count(C.byref(self._dims)), dtype=torch.float32, device=dev)
        self._anchor = torch.zeros(1, device=dev, requires_grad=True)
        self._ws_pool.clear()
        self._lens_cache.clear()

    def rebind_grad_buffer(self, buf):
        """Re-home the flat gradient buffer into `buf` (same length, float32, this device) -- e.g. a peer-mapped
        symmetric allocation the data-parallel all-reduce sums in place.  Existing .grad tensors are dropped."""
        self._ensure_flat()
        if buf.numel() != self._gflat.numel() or buf.dtype != torch.float32 or buf.device != self._gflat.device or not buf.is_contiguous():
            raise ValueError("gradient buffer must be a contiguous float32 tensor of the flat parameter count on the model's device")
        buf.zero_()
        offs = self.grad_offsets()
        self._gviews = [buf[o: o + v.numel()].view(v.shape) for o, v in zip(offs, self._gviews)]
        self._gflat = buf
        for p in self._plist:
            p.grad = None

    def grad_offsets(self):
        """float offset of every parameter (state_dict order) inside the flat parameter / gradient buffers"""
        self._ensure_flat()
        base = self._gflat.data_ptr()
        return [(v.data_ptr() - base) // 4 for v in self._gviews]

    def _ensure_flat_private_grads(self):
        """Back to a private gradient buffer (after a failed peer-memory setup)."""
        self._ensure_flat()
        self.rebind_grad_buffer(torch.zeros_like(self._flat))

    def flat_parameters(self):
        """(params, grads): the two flat fp32 buffers every parameter / .grad aliases."""
        self._ensure_flat()
        return self._flat, self._gflat

    def bucket_boundaries(self):
        """Float offsets [0, layers(0,0), layers(1,0), ..., total]: gradient bucket s+1 =
        [b[s+1], b[s+2]) is final once backward stage s has run."""
        lib = _cabi.lib()
        S = self._dims.num_stages
        return [0] + [int(lib.mstcn_bucket_boundary(C.byref(self._dims), s)) for s in range(S + 1)]

    # ------------------------------------------------------------------ dropout stream
    def set_dropout_state(self, seed, offset=0):
        """Fix the Philox stream (tests inject the same mask into the oracle)."""
        self._drop_seed, self._drop_offset = int(seed) & (2 ** 64 - 1), int(offset)

    def _next_dropout(self):
        if self._drop_seed is None:
            self._drop_seed = int(torch.empty((), dtype=torch.int64).random_().item())
        d = MstcnDropout(1 if self.training else 0, 0, self._drop_seed, self._drop_offset,
                         None if self._drop_counter is None else self._drop_counter.data_ptr())
        if self.training and self._drop_counter is None:
            self._drop_offset += 1
        return d

    # ------------------------------------------------------------------ launches
    def _lens_device(self, x_len, dev):
        key = (tuple(int(v) for v in x_len), dev)
        t = self._lens_cache.get(key)
        if t is None:
            if len(self._lens_cache) > 256:
                self._lens_cache.clear()
            t = torch.tensor(key[0], dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
            self._lens_cache[key] = t
        return t

    def _acquire_workspace(self, B, T, training, dev):
        lib = _cabi.lib()
        n = lib.mstcn_workspace_floats(C.byref(self._dims), B, T, 1 if training else 0)
        if n < 0:
            raise _cabi.MstcnError(lib.mstcn_last_error().decode())
        pool = self._ws_pool.setdefault((n, training), [])
        if pool:
            return pool.pop()
        ws = torch.empty(n, dtype=torch.float32, device=dev)
        poison = os.environ.get("MSTCN_POISON_WS")
        if poison:                      # debugging aid: a kernel that reads a workspace row nobody wrote shows up as NaN / junk
            ws.fill_(float(poison))
        return ws

    def _release_workspace(self, ws):
        for (n, _), pool in self._ws_pool.items():
            if n == ws.numel() and len(pool) < 2:
                pool.append(ws)
                return

    def _launch_forward(self, x, lens_dev, B, T, drop, training, want_out=True):
        """want_out=False (tensor-core training path only): skip the max over stages; mstcn_loss_head takes it."""
        lib = _cabi.lib()
        st = stream_ptr()
        check(lib.mstcn_pack_params(C.byref(self._dims), ptr(self._flat), ptr(self._packed), st))
        ws = self._acquire_workspace(B, T, training, x.device)
        out = torch.empty(B * T, self.n_class, dtype=torch.float32, device=x.device) if want_out else None
        winner = torch.empty(B * T, self.n_class, dtype=torch.uint8, device=x.device) if want_out else None
        check(lib.mstcn_forward(C.byref(self._dims), ptr(self._packed), ptr(x), ptr(lens_dev), self._lens_host,
                                int(self.stream_groups), B, T, C.byref(drop), 1 if training else 0, ptr(ws), ptr(out),
                                ptr(winner), st))
        if training:
            self.last_workspace = (ws, B, T)
        return out, winner, ws

    def _launch_backward(self, x, lens_dev, B, T, drop, ws, winner, gout, gscale=None, stage_hook=None):
        """Runs the backward kernels and leaves the result in every parameter's .grad."""
        lib = _cabi.lib()
        st = stream_ptr()
        if gout is not None:                # None: the gradient planes were written by mstcn_loss_head
            if gout.dtype != torch.float32 or not gout.is_cuda:
                raise RuntimeError("upstream gradient must be a CUDA float32 tensor")
            gout = gout.contiguous()
        params = self._params_in_order()
        grads = [p.grad for p in params]
        if all(g is None for g in grads):
            target, accumulate, foreign = self._gflat, 0, False
        elif all(g is not None and g.data_ptr() == v.data_ptr() for g, v in zip(grads, self._gviews)):
            target, accumulate, foreign = self._gflat, 1, False      # zeroed-in-place or accumulating
        else:
            target, accumulate, foreign = torch.empty_like(self._gflat), 0, True
        args = (C.byref(self._dims), ptr(self._packed), ptr(x), ptr(lens_dev), self._lens_host, int(self.backward_stream_groups),
                B, T, C.byref(drop), ptr(ws), ptr(winner), ptr(gout), ptr(gscale), ptr(target), accumulate)
        if stage_hook is None:
            check(lib.mstcn_backward(*args, st))
        else:
            for s in range(self._dims.num_stages - 1, -1, -1):
                check(lib.mstcn_backward_stage(*args, s, st))
                stage_hook(s)
        if foreign:
            for p, g, off_view in zip(params, grads, self._gviews):
                o = (off_view.data_ptr() - self._gflat.data_ptr()) // 4
                tv = target[o: o + p.numel()].view(p.shape)
                if g is None:
                    p.grad = tv.clone()
                else:
                    g.add_(tv)
        elif accumulate == 0:
            for p, v in zip(params, self._gviews):
                p.grad = v

    # ------------------------------------------------------------------ public API
    def _check_input(self, x, x_len, strict_len=True):
        if not isinstance(x, torch.Tensor) or x.dim() != 3:
            raise ValueError("x must be a (B, T, dim) tensor")
        if not x.is_cuda:
            raise RuntimeError("MultiStageModel (B200-native) has no CPU path: x must be a CUDA tensor")
        if x.dtype != torch.float32:
            raise RuntimeError("x must be float32")
        B, T, D = x.shape
        if D != self._dims.dim:
            raise ValueError(f"feature dim {D} != model dim {self._dims.dim}")
        if len(x_len) != B:
            raise IndexError(f"x_len has {len(x_len)} entries for a batch of {B}")   # networks.py:309 raises IndexError
        if min(x_len) < 0 or max(x_len) > T or (strict_len and max(x_len) != T):
            raise RuntimeError(f"max(x_len)={max(x_len)} must equal T={T}")          # networks.py:333 broadcast error
        if x.requires_grad:
            raise NotImplementedError("gradients w.r.t. the input features are not produced (train.py never needs them)")
        return B, T

    def forward(self, x, x_len):
        return self._forward_impl(x, x_len, strict_len=True)

    def _forward_impl(self, x, x_len, strict_len=True):
        """strict_len=False lets a data-parallel shard pad beyond its own longest video
        (parallel.local_pad_length) -- the reference contract max(x_len) == T is otherwise enforced."""
        B, T = self._check_input(x, x_len, strict_len)
        self._ensure_flat()
        if x.device != self._flat.device:
            raise RuntimeError("x and the model are on different devices")
        x = x.contiguous()
        lens_dev = self._lens_device(x_len, x.device)
        self._lens_host = (C.c_int32 * B)(*[int(v) for v in x_len])
        drop = self._next_dropout()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            return _MstcnFunction.apply(x, self._anchor, self, lens_dev, drop)
        out, _, ws = self._launch_forward(x, lens_dev, B, T, drop, training=False)
        self._release_workspace(ws)
        return out

    def forward_stages(self, x, x_len):
        """Per-stage outputs (S, B*T, n_class) -- the list canonical MS-TCN returns -- differentiable w.r.t. the
        parameters (the backward takes the gradient of every stage's logits; no max over stages is involved).
        Feed it to loss.MsTcnLoss.  Not part of the reference API."""
        B, T = self._check_input(x, x_len, True)
        self._ensure_flat()
        if x.device != self._flat.device:
            raise RuntimeError("x and the model are on different devices")
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("forward_stages needs a grad-enabled call (it keeps the per-stage logits of the training workspace)")
        x = x.contiguous()
        lens_dev = self._lens_device(x_len, x.device)
        self._lens_host = (C.c_int32 * B)(*[int(v) for v in x_len])
        return _MstcnFunction.apply(x, self._anchor, self, lens_dev, self._next_dropout(), True)

    def forward_mask(self, x, mask):
        """The (x, mask) spelling of canonical MS-TCN: mask (B, *, T) of prefix ones.
        Lengths are read back from the mask (one D2H sync)."""
        m = mask[:, 0, :] if mask.dim() == 3 else mask
        return self.forward(x, [int(v) for v in m.sum(dim=1).round().long().tolist()])

    def _ws_view(self, what, stage, layer, cols):
        ws, B, T = self.last_workspace
        off = _cabi.lib().mstcn_workspace_offset(C.byref(self._dims), B, T, 1, what, stage, layer)
        if off < 0:
            raise RuntimeError("workspace plane not available")
        return ws[off: off + B * T * cols]

    def saved_relu_outputs(self):
        """[[h (B, T, 64) per layer] per stage]: relu outputs kept by the latest grad-enabled forward
        (views into its workspace).  Test / diagnostics hook."""
        if self.last_workspace is None:
            raise RuntimeError("no grad-enabled forward has run yet")
        _, B, T = self.last_workspace
        return [[self._ws_view(1, s, l, 64).view(B, T, 64) for l in range(self._dims.num_layers)]
                for s in range(self._dims.num_stages)]

    def stage_logits(self):
        """(S, B*T, n_class) per-stage masked logits of the latest grad-enabled forward (views into
        its workspace; valid until that workspace is reused).  Not part of the reference API."""
        if self.last_workspace is None:
            raise RuntimeError("no grad-enabled forward has run yet")
        _, B, T = self.last_workspace
        return torch.stack([self._ws_view(2, s, 0, self.n_class).view(B * T, self.n_class)
                            for s in range(self._dims.num_stages)])
