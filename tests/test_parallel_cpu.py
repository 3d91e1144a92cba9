"""CPU tier, world_size 2 over gloo: the data-parallel host logic (shard planner, local padding rule,
global loss divisor, bucketed sum-all-reduce in backward-stage order) reproduces the single-process
gradients.  The per-rank compute is the numpy oracle here (no GPU in this tier); the code under test is
pytorch_video_action_b200/parallel.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat_order(params):
    from oracle import mstcn_oracle as O
    dim, S, L, C, K = O.infer_config(params)
    keys = []
    for pre in O.stage_prefixes(S):
        keys += [pre + "conv_1x1.weight", pre + "conv_1x1.bias"]
        for l in range(L):
            keys += [f"{pre}layers.{l}.conv_dilated.weight", f"{pre}layers.{l}.conv_dilated.bias",
                     f"{pre}layers.{l}.conv_1x1.weight", f"{pre}layers.{l}.conv_1x1.bias"]
        keys += [pre + "conv_out.weight", pre + "conv_out.bias"]
    return keys


def _boundaries(params):
    """[0, layers(0,0), layers(1,0), ..., total] exactly as mstcn_bucket_boundary defines them."""
    from oracle import mstcn_oracle as O
    dim, S, L, C, K = O.infer_config(params)
    keys = _flat_order(params)
    off, acc = {}, 0
    for k in keys:
        off[k] = acc
        acc += params[k].size
    pre = O.stage_prefixes(S)
    return [0] + [off[p + "layers.0.conv_dilated.weight"] for p in pre] + [acc]


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_golden, split_golden
    from oracle import mstcn_oracle as O
    from pytorch_video_action_b200.parallel import shard_videos, local_pad_length, GradBucketReducer

    g = load_golden("small_eval")
    params, _ = split_golden(g)
    dim, S, L, C, K = O.infer_config(params)
    rng = np.random.default_rng(5)
    lens = [90, 71, 64, 33, 20, 7]
    Tg = max(lens)
    feats = [rng.standard_normal((n, dim)).astype(np.float32) for n in lens]
    labs = [rng.integers(0, K, n) for n in lens]
    n_valid_global = sum(lens)

    mine = shard_videos(lens, world)[rank]
    Tl = local_pad_length([lens[i] for i in mine], Tg)
    x = np.zeros((len(mine), Tl, dim), np.float32)
    y = np.full((len(mine), Tl), -1, np.int64)
    for j, i in enumerate(mine):
        x[j, :lens[i]] = feats[i]
        y[j, :lens[i]] = labs[i]
    ll = [lens[i] for i in mine]
    # the oracle insists on max(lens) == T like the reference; when the pad rule adds a frame, append an
    # inert full-length dummy video (all ops are per-video; its labels are -1 so it gets no gradient)
    if max(ll) < Tl:
        x = np.concatenate([x, np.zeros((1, Tl, dim), np.float32)], axis=0)
        y = np.concatenate([y, np.full((1, Tl), -1, np.int64)], axis=0)
        ll = ll + [Tl]
    out, cache = O.forward(params, x, ll, keep_cache=True)
    _, gout = O.cross_entropy(out, y.reshape(-1), n_valid=n_valid_global)
    grads = O.backward(cache, gout)
    keys = _flat_order(params)
    flat = torch.from_numpy(np.concatenate([grads[k].reshape(-1) for k in keys]).astype(np.float32))
    red = GradBucketReducer(flat, _boundaries(params))
    for s in range(S - 1, -1, -1):                      # the order MultiStageModel's backward reports stages in
        red.on_stage_done(s)
    red.finish()
    if rank == 0:
        np.save(os.path.join(tmp, "dp.npy"), flat.numpy())
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden, split_golden
    from oracle import mstcn_oracle as O
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    dp = np.load(os.path.join(str(tmp_path), "dp.npy"))

    g = load_golden("small_eval")
    params, _ = split_golden(g)
    dim, S, L, C, K = O.infer_config(params)
    rng = np.random.default_rng(5)
    lens = [90, 71, 64, 33, 20, 7]
    feats = [rng.standard_normal((n, dim)).astype(np.float32) for n in lens]
    labs = [rng.integers(0, K, n) for n in lens]
    x, _, y = O.pad_batch(feats, labs)
    out, cache = O.forward(params, x, lens)
    _, gout = O.cross_entropy(out, y)
    grads = O.backward(cache, gout)
    single = np.concatenate([grads[k].reshape(-1) for k in _flat_order(params)])
    err = np.abs(dp - single).max() / np.abs(single).max()
    assert err < 1e-4, err
