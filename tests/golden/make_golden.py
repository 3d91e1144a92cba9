"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (or $MSTCN_REF)

Imports /root/reference/networks.py unchanged, drives MultiStageModel exactly as
train.py:298-332 / inference.py:113-179 do, and writes small .npz fixtures next to this
file.  The GPU box has no /root/reference; tests only read the committed .npz files.

Dropout: the reference's nn.Dropout draws from ATen's generator, which no fused kernel can
reproduce.  For the train-mode fixture each DilatedResidualLayer.dropout is replaced by a
module that multiplies by an explicit {0,2} mask (same semantics as nn.Dropout(p=0.5) in
train mode); the mask comes from oracle.dropout_scale (Philox4x32-10), the same stream the
CUDA kernels regenerate.
"""
import os
import sys
import statistics

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MSTCN_REF", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from networks import MultiStageModel  # noqa: E402  (the reference, unmodified)
from oracle import mstcn_oracle as O  # noqa: E402


class _FixedMask(nn.Module):
    def __init__(self, scale_bct):
        super().__init__()
        self.scale = scale_bct

    def forward(self, x):
        return x * self.scale


def _synth(B, T, D, lens, K, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, D, generator=g)
    y = torch.full((B, T), -1, dtype=torch.long)
    for b, l in enumerate(lens):
        x[b, l:] = 0                       # pad_batch zero-fills (train.py:188,194)
        t = 0
        while t < l:                       # piecewise-constant label runs
            run = int(torch.randint(3, 40, (1,), generator=g))
            y[b, t:min(l, t + run)] = int(torch.randint(1, K, (1,), generator=g)) if K > 1 else 0
            t += run
    return x, y.flatten()


def _run_train_step(net, x, lens, y, dropout_seed=None, dropout_offset=0):
    B, T, _ = x.shape
    layers = [l for st in [net.stage1, *net.stages] for l in st.layers]
    if dropout_seed is not None:
        net.train()
        for gi, layer in enumerate(layers):
            sc = O.dropout_scale(dropout_seed, dropout_offset, gi, B * T).reshape(B, T, 64)
            layer.dropout = _FixedMask(torch.from_numpy(sc).permute(0, 2, 1).contiguous())
    else:
        net.eval()                          # dropout = identity, grads still enabled
    net.zero_grad()
    out = net(x, lens)                                            # train.py:308
    loss = nn.CrossEntropyLoss(ignore_index=-1)(out, y)           # train.py:266-267,326
    loss.backward()                                               # train.py:328
    grads = {k: p.grad.detach().numpy().copy() for k, p in net.named_parameters()}
    return out.detach().numpy().copy(), float(loss), grads


def make_case(name, dim, S, L, K, B, T, lens, wseed, xseed, dropout_seed=None, dropout_offset=0):
    torch.manual_seed(wseed)
    net = MultiStageModel(dim, S, L, 64, K)
    x, y = _synth(B, T, dim, lens, K, xseed)
    sd = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    out, loss, grads = _run_train_step(net, x, lens, y, dropout_seed, dropout_offset)
    val, idx = torch.max(torch.from_numpy(out), 1)                # train.py:157
    payload = {"x": x.numpy(), "lens": np.array(lens), "y": y.numpy(), "out": out,
               "loss": np.float32(loss), "argmax": idx.numpy(),
               "cfg": np.array([dim, S, L, 64, K]),
               "dropout": np.array([-1 if dropout_seed is None else dropout_seed, dropout_offset])}
    payload.update({"p/" + k: v for k, v in sd.items()})
    payload.update({"g/" + k: v for k, v in grads.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print(f"{name}: loss={loss:.6f} out{out.shape} max|out|={np.abs(out).max():.4f}")


def _vote_reference(predicted, segments, fallback, stable=True):
    """The reference's own torch snippet: train.py:161-170 / inference.py:129-151.

    inference.py:148 calls torch.argsort with the default stable=False, whose order among
    EQUAL counts is implementation-defined (CPU introsort vs CUDA radix/bitonic give
    different classes).  The pinned behaviour is the stable order (lowest class first);
    the unstable CPU answer is stored too and must agree wherever position [1] is tie-free."""
    labels, tie_free = [], []
    for index in range(len(segments) - 1):
        pl = predicted[int(segments[index]): int(segments[index + 1])]
        cnt = torch.bincount(pl)
        mp = int(torch.argmax(cnt).item())
        free = True
        if fallback and mp == 0 and cnt.shape[0] > 1:
            order = torch.argsort(cnt, stable=stable)
            mp = int(order[1].item())
            srt = torch.sort(cnt).values
            free = bool(srt[1] != srt[0]) and (cnt.shape[0] < 3 or bool(srt[1] != srt[2]))
        labels.append(mp)
        tie_free.append(free)
    return labels, tie_free


def make_inference_case(name="inference_ensemble"):
    """inference.py:113-179 on synthetic videos whose segment boundaries are the first
    lines of segment.txt (re-based to 0 as data_utils.py:189 does), 2 checkpoints."""
    dim, S, L, K = 32, 2, 5, 48
    seg_lines = []
    with open(os.path.join(REF, "segment.txt")) as f:
        for line in f:
            v = [int(t) for t in line.split()]
            if v and v[-1] - v[0] <= 800:
                seg_lines.append([t - v[0] for t in v])
            if len(seg_lines) == 6:
                break
    nets = []
    for ws in (0, 1):
        torch.manual_seed(ws)
        nets.append(MultiStageModel(dim, S, L, 64, K).eval())
    payload = {"cfg": np.array([dim, S, L, 64, K]), "n_videos": np.array(len(seg_lines))}
    for mi, net in enumerate(nets):
        payload.update({f"p{mi}/" + k: v.detach().numpy().copy() for k, v in net.state_dict().items()})
    for vi, seg in enumerate(seg_lines):
        T = seg[-1]
        x, _ = _synth(1, T, dim, [T], K, 100 + vi)
        x = x * 3.0
        per_model_dev, per_model_inf = [], []
        for mi, net in enumerate(nets):
            with torch.no_grad():
                out = net(x, [T])                                  # inference.py:122
            _, predicted = torch.max(out.data, 1)                  # inference.py:123
            if vi % 3 == 1:
                # force class-0 heavy segments so the fallback branch (inference.py:147-151) runs
                predicted = torch.where(torch.arange(T) % 3 != 0, torch.zeros_like(predicted), predicted)
            elif vi % 3 == 2:
                # few bins with distinct counts: the fallback answer is tie-free
                t = torch.arange(T)
                predicted = torch.where(t % 7 < 4, torch.zeros_like(predicted),
                                        torch.where(t % 7 < 6, torch.full_like(predicted, 2 + mi),
                                                    torch.ones_like(predicted)))
            payload[f"v{vi}/argmax{mi}"] = predicted.numpy()
            payload[f"v{vi}/out{mi}"] = out.numpy()
            per_model_dev.append(_vote_reference(predicted, seg, False)[0])
            lab, free = _vote_reference(predicted, seg, True, stable=True)
            lab_u, _ = _vote_reference(predicted, seg, True, stable=False)
            per_model_inf.append(lab)
            payload[f"v{vi}/vote_inf_unstable{mi}"] = np.array(lab_u)
            payload[f"v{vi}/vote_inf_tiefree{mi}"] = np.array(free)
        final = []
        for j in range(len(seg) - 1):                               # inference.py:159-179
            votes = [pm[j] for pm in per_model_inf if pm[j] != 0]
            try:
                final.append(statistics.mode(votes))
            except Exception:
                final.append(0)
        payload[f"v{vi}/x"] = x.numpy()
        payload[f"v{vi}/segments"] = np.array(seg)
        payload[f"v{vi}/vote_dev"] = np.array(per_model_dev)
        payload[f"v{vi}/vote_inf"] = np.array(per_model_inf)
        payload[f"v{vi}/final"] = np.array(final)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print(name, "videos", len(seg_lines))


if __name__ == "__main__":
    torch.set_num_threads(4)
    # eval-mode-with-grad (dropout off): ragged lens incl. len=1 and len=T
    make_case("small_eval", dim=24, S=3, L=4, K=48, B=3, T=97, lens=[97, 60, 1], wseed=0, xseed=11)
    # train mode with the Philox mask injected on the reference side
    make_case("small_train", dim=24, S=3, L=4, K=48, B=3, T=97, lens=[97, 60, 33], wseed=0, xseed=12,
              dropout_seed=0x1234ABCD5678, dropout_offset=7)
    # dilation >= T (fact 0.4): L=8 -> d up to 128 on T=50; tiny class count like the ctor default
    make_case("deep_d_ge_T", dim=16, S=2, L=8, K=5, B=2, T=50, lens=[50, 17], wseed=3, xseed=13)
    make_inference_case()
