"""Per-frame argmax, segment majority vote and checkpoint-ensemble vote
(train.py:143-176 evaluate; inference.py:113-192), on the GPU where the reference loops in Python
with one .item() sync per segment."""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


def frame_argmax(outputs):
    """torch.max(outputs.data, 1) (train.py:157, inference.py:123): (values, int64 indices),
    first index on ties."""
    lib = _cabi.lib()
    if not outputs.is_cuda or outputs.dtype != torch.float32 or outputs.dim() != 2:
        raise RuntimeError("frame_argmax needs a CUDA float32 (N, n_class) tensor (no CPU path)")
    outputs = outputs.detach().contiguous()
    n, k = outputs.shape
    idx = torch.empty(n, dtype=torch.int64, device=outputs.device)
    val = torch.empty(n, dtype=torch.float32, device=outputs.device)
    check(lib.mstcn_frame_argmax(ptr(outputs), n, k, ptr(idx), ptr(val), stream_ptr()))
    return val, idx


def segment_vote(predicted, bounds, n_class, inference_fallback=False):
    """argmax(bincount(predicted[s:e])) per segment, lowest class on ties (train.py:161-170).
    inference_fallback=True adds the class-0 rule of inference.py:147-151 (second entry of the
    ascending stable argsort of the bincount).  bounds: n_seg+1 frame boundaries (list or int32
    tensor).  Returns an int32 CUDA tensor of n_seg labels (one D2H copy for the whole video,
    not one sync per segment)."""
    lib = _cabi.lib()
    if not predicted.is_cuda or predicted.dtype != torch.int64:
        raise RuntimeError("segment_vote needs CUDA int64 predictions (no CPU path)")
    if not isinstance(bounds, torch.Tensor):
        bounds = torch.tensor([int(b) for b in bounds], dtype=torch.int32)
    bounds = bounds.to(device=predicted.device, dtype=torch.int32).contiguous()
    n_seg = bounds.numel() - 1
    labels = torch.empty(max(n_seg, 0), dtype=torch.int32, device=predicted.device)
    check(lib.mstcn_segment_vote(ptr(predicted.contiguous()), ptr(bounds), n_seg, n_class,
                                 1 if inference_fallback else 0, ptr(labels), stream_ptr()))
    return labels


def label_runs(labels):
    """get_label_length_seq (train.py:70-83): run labels and their boundaries, from a label tensor."""
    labels = labels.flatten()
    n = labels.numel()
    change = torch.nonzero(labels[1:] != labels[:-1]).flatten() + 1
    bounds = torch.cat([torch.zeros(1, dtype=change.dtype, device=change.device), change,
                        torch.full((1,), n, dtype=change.dtype, device=change.device)])
    return labels[bounds[:-1]], bounds.to(torch.int32)


def ensemble_vote(per_model_labels):
    """statistics.mode over the checkpoints' votes in CLI order with zero votes dropped
    (inference.py:151,159-179): first-seen value wins ties; no votes -> 0."""
    n_seg = len(per_model_labels[0])
    out = []
    for j in range(n_seg):
        votes = [int(ml[j]) for ml in per_model_labels if int(ml[j]) != 0]
        best, best_n, seen = 0, 0, {}
        for v in votes:
            seen[v] = seen.get(v, 0) + 1
        for v in votes:
            if seen[v] > best_n:
                best, best_n = v, seen[v]
        out.append(best)
    return out


def evaluate_video(outputs, labels, n_class):
    """One iteration of evaluate() (train.py:153-172) for a batch-1 video: returns
    (correct_frames, total_frames, correct_segments, total_segments) as Python ints."""
    _, predicted = frame_argmax(outputs)
    run_labels, bounds = label_runs(labels)
    votes = segment_vote(predicted, bounds, n_class, inference_fallback=False)
    stats = torch.stack([(predicted == labels).sum(), (votes.to(run_labels.dtype) == run_labels).sum()]).tolist()
    return int(stats[0]), labels.numel(), int(stats[1]), run_labels.numel()


def ensemble_predict(nets, videos, segments, n_class):
    """The inference.py:113-179 loop for a whole test set with ONE device-to-host transfer: every video goes through every
    checkpoint at batch 1 with T = its own length, exactly as inference.py:78 feeds them (a padded batch would NOT be the
    same function: the stage-1 conv is unmasked, so frames behind a video's end carry its bias into the dilated layers,
    SURVEY fact 0.5), per-frame argmax and the inference vote rule run on the GPU, all votes are gathered on the device
    and read back once; the checkpoint mode (statistics.mode, zero votes dropped) is taken on the host as in the reference.

    nets: MultiStageModels in eval mode, in CLI order; videos: list of (T, dim) or (1, T, dim) CUDA float32 tensors;
    segments: per video the n_seg+1 frame boundaries.  Returns a list (one per video) of lists of n_seg labels."""
    per_model, counts = [], []
    with torch.no_grad():
        for x, seg in zip(videos, segments):
            x = x if x.dim() == 3 else x.unsqueeze(0)
            if x.shape[0] != 1:
                raise ValueError("ensemble_predict takes single videos (the reference runs inference at batch 1)")
            counts.append(len(seg) - 1)
            for net in nets:
                _, pred = frame_argmax(net(x, [x.shape[1]]))
                per_model.append(segment_vote(pred, seg, n_class, inference_fallback=True))
    if not per_model:
        return []
    flat = torch.cat(per_model).cpu().tolist()          # the only synchronisation
    out, pos, m = [], 0, len(nets)
    for n in counts:
        votes = [flat[pos + j * n: pos + (j + 1) * n] for j in range(m)]
        pos += m * n
        out.append(ensemble_vote(votes))
    return out
