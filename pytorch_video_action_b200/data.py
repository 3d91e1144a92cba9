"""Device-resident input feed: the reference's `pad_batch` collate (train.py:183-205, inference.py:32-44)
as one gather kernel over a dataset that lives in HBM.

The reference pads on the host and ships every batch over PCIe with pageable copies (train.py:301-302);
at this path's speed that feature stream (1 600 B/frame) is the first thing to stall.  An I3D feature set
of a few thousand videos is a few GB -- it fits in a B200's 180 GB many times over -- so it is uploaded
once and a batch costs one HBM-bound kernel.
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr

TARGET_PAD = -1            # train.py:12 _TARGET_PAD


class DeviceFeatureStore:
    """All videos' per-frame features (and labels) concatenated on one GPU.

    features: sequence of (T_i, dim) float32 arrays/tensors; labels: matching sequence of (T_i,) integer
    arrays/tensors or None.  `pad_batch(indices)` returns exactly what the reference's collate returns for
    those videos -- `(padded_seqs (B, max_len, dim) float32, x_len list[int], target (B*max_len,) int64)` --
    with the tensors already on the device."""

    def __init__(self, features, labels=None, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceFeatureStore keeps the dataset in GPU memory (no CPU path)")
        feats = [torch.as_tensor(f, dtype=torch.float32) for f in features]
        if not feats:
            raise ValueError("empty dataset")
        self.dim = int(feats[0].shape[1])
        if self.dim % 4 != 0 or any(f.dim() != 2 or f.shape[1] != self.dim for f in feats):
            raise ValueError("features must be (T_i, dim) with one dim, a multiple of 4")
        self.lengths = [int(f.shape[0]) for f in feats]
        offs = [0]
        for n in self.lengths:
            offs.append(offs[-1] + n)
        self.feats = torch.cat(feats).to(dev).contiguous()
        self.offsets = torch.tensor(offs, dtype=torch.int64, device=dev)
        self.labels = None
        if labels is not None:
            labs = [torch.as_tensor(l, dtype=torch.int64).reshape(-1) for l in labels]
            if [int(l.numel()) for l in labs] != self.lengths:
                raise ValueError("labels must have one entry per frame")
            self.labels = torch.cat(labs).to(dev).contiguous()
        self.device = dev

    def __len__(self):
        return len(self.lengths)

    def pad_batch(self, indices, pad_to=None, with_lens_tensor=False, out=None):
        """indices: the videos of the batch, in batch order.  pad_to: padded length (default: the longest selected
        video, like the reference; a data-parallel shard passes parallel.local_pad_length(...)).
        out (optional): (x, y) or (x, y, lens_dev) preallocated CUDA tensors of the batch's shape to gather into (the
        fixed input buffers a captured CUDA graph reads) instead of fresh allocations."""
        idx = [int(i) for i in indices]
        if not idx:
            raise ValueError("empty batch")
        x_len = [self.lengths[i] for i in idx]
        T = int(pad_to) if pad_to is not None else max(x_len)
        if T < max(x_len):
            raise ValueError("pad_to is shorter than the longest selected video")
        B = len(idx)
        # the only bytes that cross PCIe per batch: B video indices
        vid = torch.tensor(idx, dtype=torch.int32).pin_memory().to(self.device, non_blocking=True)
        if out is not None:
            x, y = out[0], out[1]
            if (tuple(x.shape) != (B, T, self.dim) or x.dtype != torch.float32 or y.numel() != B * T or y.dtype != torch.int64
                    or x.device != self.device or y.device != self.device or not x.is_contiguous() or not y.is_contiguous()):
                raise ValueError("out buffers do not match the batch (B, T, dim) / (B*T,) float32 / int64 on the store's device")
            lens_dev = out[2] if len(out) > 2 else torch.empty(B, dtype=torch.int32, device=self.device)
        else:
            x = torch.empty(B, T, self.dim, dtype=torch.float32, device=self.device)
            y = torch.empty(B * T, dtype=torch.int64, device=self.device)
            lens_dev = torch.empty(B, dtype=torch.int32, device=self.device)
        check(_cabi.lib().mstcn_pad_batch(ptr(self.feats), ptr(self.labels) if self.labels is not None else None,
                                          ptr(self.offsets), ptr(vid), B, T, self.dim, ptr(x), ptr(y), ptr(lens_dev),
                                          stream_ptr()))
        if with_lens_tensor:
            return x, x_len, y, lens_dev
        return x, x_len, y
