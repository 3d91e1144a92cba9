"""CPU oracle for the MS-TCN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (pytorch_video_action_b200/) never does.

This is a plain-numpy restatement of the algorithm the reference executes through
PyTorch (the arithmetic lives in torch, a third-party dependency the reference does
not pin: README.md:9-14 asks for "PyTorch >= 1.1.0").  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Parity pinning: the reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the reference itself, run in the build
container by tests/golden/make_golden.py (which imports /root/reference/networks.py
unchanged) and committed as tests/golden/*.npz.  tests/test_oracle.py replays them.
The dropout bit-stream (Philox4x32-10) is additionally pinned to the Random123
known-answer vectors.  The truncated-MSE term (north-star addition, not in the
reference) has no reference pin: "parity unpinned" for `ms_tcn_paper_loss` only.

Layouts are channels-last: activations (B, T, C); logits (B*T, K) exactly as
MultiStageModel.forward returns them (networks.py:317-320).
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants).  This is the
# counter-based generator the CUDA kernels use to regenerate the dropout keep-bits in
# backward instead of storing a mask.  nn.Dropout() in networks.py:341 has p = 0.5, so
# one random bit per element suffices.
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds. All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _PHILOX_M0
            p1 = c2.astype(np.uint64) * _PHILOX_M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PHILOX_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PHILOX_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_keep_bits(seed: int, offset: int, layer: int, n_frames: int):
    """Keep-bits for one dilated residual layer: uint32 (n_frames, 2).

    Frame n (= b*T_pad + t), channel c is KEPT iff bit (c & 31) of word (c >> 5) is 1.
    Counter = (n, layer, offset_lo, offset_hi), key = (seed_lo, seed_hi).
    Mirrors csrc/philox.cuh::dropout_bits (same integer arithmetic, bit-exact).
    """
    n = np.arange(n_frames, dtype=np.uint32)
    r0, r1, _, _ = philox4x32_10(
        n, np.uint32(layer), np.uint32(offset & 0xFFFFFFFF), np.uint32((offset >> 32) & 0xFFFFFFFF),
        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack([r0, r1], axis=1)


def dropout_scale(seed: int, offset: int, layer: int, n_frames: int, channels: int = 64):
    """(n_frames, channels) float32 multiplier in {0, 2}: nn.Dropout(p=0.5) in train mode
    scales kept values by 1/(1-p) = 2 (networks.py:341,346)."""
    assert channels == 64
    bits = dropout_keep_bits(seed, offset, layer, n_frames)
    c = np.arange(channels)
    keep = (bits[:, c >> 5] >> (c & 31).astype(np.uint32)) & np.uint32(1)
    return keep.astype(np.float32) * np.float32(2.0)


# --------------------------------------------------------------------------------------
# Parameters
# --------------------------------------------------------------------------------------

def stage_prefixes(num_stages: int):
    """state_dict prefixes in execution order (networks.py:301-302)."""
    return ["stage1."] + [f"stages.{s}." for s in range(num_stages - 1)]


def infer_config(params: dict):
    """(dim, num_stages, num_layers, num_f_maps, n_class) from state_dict shapes."""
    w = params["stage1.conv_1x1.weight"]
    num_f_maps, dim = w.shape[0], w.shape[1]
    n_class = params["stage1.conv_out.weight"].shape[0]
    num_layers = 0
    while f"stage1.layers.{num_layers}.conv_dilated.weight" in params:
        num_layers += 1
    num_stages = 1
    while f"stages.{num_stages - 1}.conv_1x1.weight" in params:
        num_stages += 1
    return dim, num_stages, num_layers, num_f_maps, n_class


def _mask_from_lens(lens, B, T, dtype):
    """networks.py:307-309: mask[i, :, :x_len[i]] = 1 (only row 0 is ever read)."""
    m = np.zeros((B, T, 1), dtype=dtype)
    for i in range(B):
        m[i, : lens[i], 0] = 1
    return m


def _shift(x, off):
    """x[:, t+off, :] with zero fill outside [0, T): the conv's zero padding
    (nn.Conv1d(..., padding=d, dilation=d), networks.py:339)."""
    B, T, C = x.shape
    out = np.zeros_like(x)
    if off == 0:
        out[:] = x
    elif abs(off) < T:
        if off > 0:
            out[:, : T - off] = x[:, off:]
        else:
            out[:, -off:] = x[:, : T + off]
    return out


def _softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


# --------------------------------------------------------------------------------------
# Forward (networks.py:305-347)
# --------------------------------------------------------------------------------------

def layer_forward(a, Wd, bd, W1, b1, d, m, dm=None):
    """DilatedResidualLayer.forward (networks.py:343-347) on channels-last (B, T, C) activations.
    Wd (out, in, 3), W1 (out, in); m (B, T, 1) the 0/1 mask; dm None (eval) or the {0, 2} dropout multiplier.
    Returns (y, h, u): layer output, relu output, pre-activation."""
    u = bd + _shift(a, -d) @ Wd[:, :, 0].T + a @ Wd[:, :, 1].T + _shift(a, d) @ Wd[:, :, 2].T
    h = np.maximum(u, 0)                               # F.relu, :344
    o = h @ W1.T + b1                                  # conv_1x1, :345
    y = (a + (o * dm if dm is not None else o)) * m    # dropout :346, residual + mask :347
    return y, h, u


def forward(params: dict, x, lens, train_dropout=None, dtype=np.float32, keep_cache=True):
    """MultiStageModel.forward(x, x_len) (networks.py:305-320).

    x: (B, T, D) batch-first; lens: list[int] with max(lens) == T.
    train_dropout: None (eval: dropout is identity) or a callable
        f(global_layer_index, n_frames) -> (n_frames, C) multiplier in {0, 2}.
    Returns (out (B*T, K), cache) where cache feeds `backward`.
    """
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    dim, S, L, C, K = infer_config(P)
    x = np.asarray(x, dtype=dtype)
    B, T, D = x.shape
    if D != dim:
        raise ValueError("feature dim mismatch")
    if len(lens) != B or max(lens) != T:
        raise ValueError("x_len must have one entry per video and max(x_len) == T")
    m = _mask_from_lens(lens, B, T, dtype)
    cache = {"lens": list(lens), "m": m, "stages": [], "cfg": (dim, S, L, C, K), "P": P}

    inp = x
    stage_logits = []
    for si, pre in enumerate(stage_prefixes(S)):
        sc = {"inp": inp, "layers": []}
        # SingleStageModel.forward: conv_1x1, NOT masked (networks.py:330)
        a = inp @ P[pre + "conv_1x1.weight"][:, :, 0].T + P[pre + "conv_1x1.bias"]
        for li in range(L):
            d = 2 ** li  # networks.py:326
            Wd = P[f"{pre}layers.{li}.conv_dilated.weight"]   # (out, in, tap)
            bd = P[f"{pre}layers.{li}.conv_dilated.bias"]
            W1 = P[f"{pre}layers.{li}.conv_1x1.weight"][:, :, 0]
            b1 = P[f"{pre}layers.{li}.conv_1x1.bias"]
            if train_dropout is not None:              # dropout, :346
                dm = np.asarray(train_dropout(si * L + li, B * T), dtype=dtype).reshape(B, T, C)
            else:
                dm = None
            y, h, u = layer_forward(a, Wd, bd, W1, b1, d, m, dm)   # DilatedResidualLayer.forward (networks.py:343-347)
            if keep_cache:
                sc["layers"].append({"x": a, "h": h, "u": u, "dm": dm, "d": d})
            a = y
        z = (a @ P[pre + "conv_out.weight"][:, :, 0].T + P[pre + "conv_out.bias"]) * m  # :333
        sc["a_last"] = a
        sc["z"] = z
        stage_logits.append(z)
        if si < S - 1:
            p = _softmax(z)                            # F.softmax(out, dim=1), :314
            sc["p"] = p
            inp = p * m                                # * mask[:, 0:1, :], :314
        cache["stages"].append(sc)

    stack = np.stack(stage_logits, axis=0)             # torch.cat, :312,315
    winner = np.argmax(stack, axis=0)                  # first index on ties, like torch.max
    out = np.take_along_axis(stack, winner[None], axis=0)[0]   # torch.max(outputs, 0)[0], :319
    cache["winner"] = winner
    cache["stage_logits"] = stack
    return out.reshape(B * T, K), cache


# --------------------------------------------------------------------------------------
# Loss (train.py:12,266-267,326)
# --------------------------------------------------------------------------------------

def cross_entropy(out, labels, ignore_index=-1, n_valid=None):
    """nn.CrossEntropyLoss(ignore_index=-1): mean over valid rows of -log_softmax[y].
    Returns (loss, dloss/dout).  n_valid overrides the divisor (data-parallel shards
    divide by the GLOBAL valid-frame count)."""
    out = np.asarray(out)
    labels = np.asarray(labels).astype(np.int64)
    valid = labels != ignore_index
    nv = int(valid.sum()) if n_valid is None else int(n_valid)
    zmax = out.max(axis=1, keepdims=True)
    lse = zmax[:, 0] + np.log(np.exp(out - zmax).sum(axis=1))
    safe = np.where(valid, labels, 0)
    nll = lse - out[np.arange(out.shape[0]), safe]
    loss = (nll * valid).sum(dtype=np.float64) / max(nv, 1)
    g = np.exp(out - lse[:, None])
    g[np.arange(out.shape[0]), safe] -= 1
    g = g * valid[:, None] / max(nv, 1)
    return out.dtype.type(loss), g.astype(out.dtype)


def ms_tcn_paper_loss(stage_logits, labels, lens, lam=0.15, tau=4.0, ignore_index=-1):
    """Canonical MS-TCN loss (Farha & Gall, CVPR'19): sum over stages of CE +
    lam * mean(clamp((logp[t] - logp[t-1].detach())^2, 0, tau^2) * mask[t]).
    NOT in the reference -> PARITY UNPINNED.  stage_logits: (S, B, T, K).
    Returns the scalar loss only (used to check the optional fused loss mode)."""
    S, B, T, K = stage_logits.shape
    m = _mask_from_lens(lens, B, T, stage_logits.dtype)
    total = 0.0
    for s in range(S):
        z = stage_logits[s]
        ce, _ = cross_entropy(z.reshape(B * T, K), labels, ignore_index)
        zmax = z.max(axis=-1, keepdims=True)
        logp = z - zmax - np.log(np.exp(z - zmax).sum(axis=-1, keepdims=True))
        diff = np.clip((logp[:, 1:] - logp[:, :-1]) ** 2, 0, tau * tau)
        total += float(ce) + lam * float((diff * m[:, 1:]).mean())
    return total


# --------------------------------------------------------------------------------------
# Backward: what autograd replays for loss.backward() (train.py:328)
# --------------------------------------------------------------------------------------

def backward(cache, gout):
    """Gradients of sum(out * gout) w.r.t. every state_dict tensor.
    gout: (B*T, K) upstream gradient of MultiStageModel.forward's return value.
    cache["winner"] and each layer's optional "relu_mask" select the sub-gradient at the two kinds of
    kink in the model (max over stages, ReLU)."""
    P = cache["P"]
    dim, S, L, C, K = cache["cfg"]
    m = cache["m"]
    B, T, _ = m.shape
    dtype = m.dtype
    gout = np.asarray(gout, dtype=dtype).reshape(B, T, K)
    winner = cache["winner"]
    grads = {}
    prefixes = stage_prefixes(S)
    g_inp_next = None   # gradient w.r.t. the NEXT stage's input (p * m)
    for si in range(S - 1, -1, -1):
        pre = prefixes[si]
        sc = cache["stages"][si]
        # torch.max over stages routes each (frame, class) gradient to the winning stage
        gz = gout * (winner == si)
        if g_inp_next is not None:
            # inp = softmax(z) * m  (networks.py:314); no detach -> softmax backward
            gp = g_inp_next * m
            p = sc["p"]
            gz = gz + p * (gp - (gp * p).sum(axis=-1, keepdims=True))
        gz = gz * m                                    # conv_out(...) * mask, :333
        a = sc["a_last"]
        grads[pre + "conv_out.weight"] = np.einsum("btk,btc->kc", gz, a)[:, :, None]
        grads[pre + "conv_out.bias"] = gz.sum(axis=(0, 1))
        ga = gz @ P[pre + "conv_out.weight"][:, :, 0]
        for li in range(L - 1, -1, -1):
            lc = sc["layers"][li]
            d, xin, h, dm = lc["d"], lc["x"], lc["h"], lc["dm"]
            Wd = P[f"{pre}layers.{li}.conv_dilated.weight"]
            W1 = P[f"{pre}layers.{li}.conv_1x1.weight"][:, :, 0]
            g = ga * m                                 # (x + out) * mask, :347
            go = g * dm if dm is not None else g       # dropout backward
            grads[f"{pre}layers.{li}.conv_1x1.weight"] = np.einsum("bto,btc->oc", go, h)[:, :, None]
            grads[f"{pre}layers.{li}.conv_1x1.bias"] = go.sum(axis=(0, 1))
            # relu backward; "relu_mask" lets a test impose the sub-gradient choice made at a kink (u == 0
            # up to rounding) by the implementation under test -- see tests/parity.py
            gu = (go @ W1) * lc.get("relu_mask", h > 0)
            grads[f"{pre}layers.{li}.conv_dilated.bias"] = gu.sum(axis=(0, 1))
            gWd = np.zeros_like(Wd)
            gx = g.copy()                              # residual branch
            for k in range(3):
                off = (k - 1) * d
                gWd[:, :, k] = np.einsum("bto,btc->oc", gu, _shift(xin, off))
                gx += _shift(gu, -off) @ Wd[:, :, k]
            grads[f"{pre}layers.{li}.conv_dilated.weight"] = gWd
            ga = gx
        inp = sc["inp"]
        # stage-input 1x1 is unmasked: padded frames still feed the bias gradient (0.5)
        grads[pre + "conv_1x1.weight"] = np.einsum("bto,btc->oc", ga, inp)[:, :, None]
        grads[pre + "conv_1x1.bias"] = ga.sum(axis=(0, 1))
        g_inp_next = ga @ P[pre + "conv_1x1.weight"][:, :, 0] if si > 0 else None
    return grads


# --------------------------------------------------------------------------------------
# Post-processing: per-frame argmax, segment vote, ensemble vote
# --------------------------------------------------------------------------------------

def frame_argmax(out):
    """torch.max(outputs.data, 1) (train.py:157, inference.py:123): first index on ties."""
    out = np.asarray(out)
    idx = np.argmax(out, axis=1).astype(np.int64)
    return out[np.arange(out.shape[0]), idx], idx


def label_runs(labels):
    """get_label_length_seq (train.py:70-83, inference.py:49-62): run labels + boundaries."""
    labels = [int(v) for v in labels]
    label_seq, bounds, start = [], [0], 0
    for i in range(len(labels)):
        if labels[i] != labels[start]:
            label_seq.append(labels[start])
            bounds.append(i)
            start = i
    label_seq.append(labels[start])
    bounds.append(len(labels))
    return label_seq, bounds


def segment_vote(pred, bounds, inference_fallback=False):
    """Per segment argmax(bincount(pred[s:e])) -- lowest class wins ties (train.py:161-170).

    inference_fallback=True adds inference.py:147-151: if the vote is class 0 and the
    bincount has more than one bin, take argsort(bincount)[1] (ASCENDING, stable: the
    class with the second-smallest count among 0..max(pred)); a segment whose vote is
    still 0 is reported as 0 here (the caller drops that model's vote, inference.py:151).
    """
    pred = np.asarray(pred).astype(np.int64)
    labels = []
    for i in range(len(bounds) - 1):
        s, e = int(bounds[i]), int(bounds[i + 1])
        cnt = np.bincount(pred[s:e])
        lab = int(np.argmax(cnt))
        if inference_fallback and lab == 0 and cnt.shape[0] > 1:
            lab = int(np.argsort(cnt, kind="stable")[1])
        labels.append(lab)
    return labels


def ensemble_vote(per_model_labels):
    """statistics.mode over the per-checkpoint labels in CLI order, zero votes dropped
    (inference.py:151,159-179).  On Python >= 3.8 mode() returns the first-seen mode;
    an empty vote list yields label 0 (inference.py:176-179)."""
    n_seg = len(per_model_labels[0])
    out = []
    for j in range(n_seg):
        votes = [int(ml[j]) for ml in per_model_labels if int(ml[j]) != 0]
        if not votes:
            out.append(0)
            continue
        best, best_n = votes[0], 0
        seen = {}
        for v in votes:
            seen[v] = seen.get(v, 0) + 1
        for v in votes:                      # first-seen order
            if seen[v] > best_n:
                best, best_n = v, seen[v]
        out.append(best)
    return out


# --------------------------------------------------------------------------------------
# Adam (train.py:273) -- torch.optim.Adam semantics, eps outside the sqrt
# --------------------------------------------------------------------------------------

def adam_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """One torch.optim.Adam update (no weight decay, no amsgrad). step counts from 1."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# --------------------------------------------------------------------------------------
# pad_batch (train.py:183-205) -- the input contract
# --------------------------------------------------------------------------------------

def pad_batch(features, labels, batchsize=None, pad_value=-1):
    """features: list of (T_i, D); labels: list of (T_i,) int. Returns (x, x_len, target)."""
    lens = [f.shape[0] for f in features]
    B = batchsize or len(features)
    T = max(lens)
    D = features[0].shape[1]
    x = np.zeros((B, T, D), dtype=np.float32)
    y = np.full((B, T), pad_value, dtype=np.int64)
    for i, l in enumerate(lens):
        x[i, :l] = features[i][:l]
        y[i, :l] = labels[i][:l]
    return x, lens, y.reshape(-1)
