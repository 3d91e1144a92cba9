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
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
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
        if min(idx) < 0 or max(idx) >= len(self.lengths):      # (the gather kernel indexes the offset table with these)
            raise IndexError(f"video index out of range [0, {len(self.lengths)})")
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


class RaggedBatchUploader:
    """Host -> device feed of one batch WITHOUT its padding (train.py:183-205 and :301-302, reordered: ship first, pad second).

    The reference pads on the host and copies the padded (B, max_len, dim) tensor; the zeros beyond each `x_len` carry
    nothing.  Here the collate's valid frames travel as one ragged pinned block -- (sum(x_len), dim) float32 features and
    (sum(x_len),) int64 labels, videos in batch order -- and `pad_batch_kernel` builds exactly the reference's padded batch
    (zeros / -1 beyond `x_len`) on the device.  At BASELINE config 2 that is 33.8 MB instead of 51.2 MB per step over PCIe.

        up = RaggedBatchUploader(x_len, dim, device="cuda")            # once per batch shape
        x, y = up.upload(ragged_x_pinned, ragged_y_pinned)             # on the current stream (use a copy stream to overlap)
        x, y = up.upload(rx, ry, out=(x_buf, y_buf))                   # ... into fixed buffers (CUDA-graph inputs)
    """

    def __init__(self, x_len, dim, pad_to=None, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("RaggedBatchUploader targets GPU memory (no CPU path)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.x_len = [int(v) for v in x_len]
        if not self.x_len or min(self.x_len) < 1:
            raise ValueError("x_len must list at least one video of at least one frame")
        if dim % 4 != 0:
            raise ValueError("dim must be a multiple of 4")
        self.dim, self.B = int(dim), len(self.x_len)
        self.T = int(pad_to) if pad_to is not None else max(self.x_len)
        if self.T < max(self.x_len):
            raise ValueError("pad_to is shorter than the longest video")
        self.total = sum(self.x_len)
        offs = [0]
        for n in self.x_len:
            offs.append(offs[-1] + n)
        self.device = dev
        self._feats = torch.empty(self.total, self.dim, dtype=torch.float32, device=dev)
        self._labels = torch.empty(self.total, dtype=torch.int64, device=dev)
        self._offsets = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._vid = torch.arange(self.B, dtype=torch.int32, device=dev)
        self._lens_dev = torch.empty(self.B, dtype=torch.int32, device=dev)

    @property
    def h2d_bytes(self):
        return self.total * self.dim * 4 + self.total * 8

    def upload(self, ragged_x, ragged_y=None, out=None):
        """ragged_x: (sum(x_len), dim) float32 host tensor (pinned for an asynchronous copy); ragged_y: (sum(x_len),) int64
        or None.  Returns (x (B, T, dim), y (B*T,) int64 or None) on the device; out=(x, y) gathers into given buffers."""
        if tuple(ragged_x.shape) != (self.total, self.dim) or ragged_x.dtype != torch.float32:
            raise ValueError(f"ragged_x must be ({self.total}, {self.dim}) float32")
        self._feats.copy_(ragged_x, non_blocking=True)
        if ragged_y is not None:
            if ragged_y.numel() != self.total or ragged_y.dtype != torch.int64:
                raise ValueError(f"ragged_y must hold {self.total} int64 labels")
            self._labels.copy_(ragged_y.reshape(-1), non_blocking=True)
        if out is not None:
            x, y = out[0], out[1]
            if (tuple(x.shape) != (self.B, self.T, self.dim) or x.dtype != torch.float32 or x.device != self.device or not x.is_contiguous()
                    or (y is not None and (y.numel() != self.B * self.T or y.dtype != torch.int64 or y.device != self.device))):
                raise ValueError("out buffers do not match (B, T, dim) float32 / (B*T,) int64 on the uploader's device")
        else:
            x = torch.empty(self.B, self.T, self.dim, dtype=torch.float32, device=self.device)
            y = torch.empty(self.B * self.T, dtype=torch.int64, device=self.device) if ragged_y is not None else None
        if ragged_y is not None and y is None:
            raise ValueError("labels were given but out has no label buffer")
        check(_cabi.lib().mstcn_pad_batch(ptr(self._feats), ptr(self._labels) if ragged_y is not None else None, ptr(self._offsets),
                                          ptr(self._vid), self.B, self.T, self.dim, ptr(x), ptr(y) if ragged_y is not None else None,
                                          ptr(self._lens_dev), stream_ptr()))
        return x, (y if ragged_y is not None else None)
