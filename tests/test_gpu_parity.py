"""GPU tier: the CUDA path (through the drop-in class and the C ABI) against the golden vectors of
the unmodified reference and against the numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): logits and gradients 1e-3 relative (max|a-b| / max|b| per
tensor), loss 1e-4 absolute, integer argmax / vote outputs bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden, split_golden, rel_err
from oracle import mstcn_oracle as O

pytestmark = pytest.mark.gpu

TOL_REL = 1e-3
TOL_LOSS = 1e-4


def _model_from_params(params, dev="cuda", tensor_cores=True):
    from pytorch_video_action_b200 import MultiStageModel
    dim, S, L, Cc, K = O.infer_config(params)
    net = MultiStageModel(dim, S, L, Cc, K)
    net.tensor_cores = tensor_cores
    net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()})   # strict, like train.py:264
    return net.to(dev)


def _run(net, x, lens, y, fused_loss=True):
    from pytorch_video_action_b200 import FrameCrossEntropy
    net.zero_grad()
    out = net(torch.from_numpy(x).cuda(), lens)
    crit = FrameCrossEntropy() if fused_loss else torch.nn.CrossEntropyLoss(ignore_index=-1)
    loss = crit(out, torch.from_numpy(y).cuda())
    loss.backward()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in net.named_parameters()}
    return out.detach().cpu().numpy(), float(loss), grads


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("fused_loss", [True, False])
@pytest.mark.parametrize("name", ["small_eval", "small_train", "deep_d_ge_T"])
def test_golden_forward_backward(name, fused_loss, tensor_cores):
    """tensor_cores=True: tcgen05 3xTF32 layer kernels (the default); False: exact fp32 FFMA kernels."""
    g = load_golden(name)
    params, ref_grads = split_golden(g)
    lens = [int(v) for v in g["lens"]]
    seed, off = (int(v) for v in g["dropout"])
    net = _model_from_params(params, tensor_cores=tensor_cores)
    if seed >= 0:
        net.train()
        net.set_dropout_state(seed, off)
    else:
        net.eval()
    out, loss, grads = _run(net, g["x"], lens, g["y"], fused_loss)
    assert rel_err(out, g["out"]) < TOL_REL
    assert abs(loss - float(g["loss"])) < TOL_LOSS
    errs = {k: rel_err(grads[k], ref_grads[k]) for k in ref_grads}
    worst = max(errs, key=errs.get)
    assert errs[worst] < TOL_REL, (worst, errs[worst])
    from pytorch_video_action_b200 import frame_argmax
    _, idx = frame_argmax(torch.from_numpy(out).cuda())
    assert np.array_equal(idx.cpu().numpy(), g["argmax"])


def test_dropout_stream_matches_oracle_bits():
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    n = 1000
    out = torch.empty(n, 64, device="cuda")
    for seed, off, layer in [(0, 0, 0), (0x1234ABCD5678, 7, 11), (2 ** 63 + 5, 2 ** 33 + 1, 39)]:
        d = _cabi.MstcnDropout(1, 0, seed, off)
        _cabi.check(lib.mstcn_dropout_scale(C.byref(d), layer, n, _cabi.ptr(out), _cabi.stream_ptr()))
        assert np.array_equal(out.cpu().numpy(), O.dropout_scale(seed, off, layer, n))


CASES = [
    # dim, S, L, K, lens (T = max), train
    (16, 2, 3, 48, [130, 64, 1], False),          # T not a tile multiple, len=1, len on a tile edge
    (8, 1, 4, 2, [70], False),                    # single stage, ctor-default class count, B=1
    (12, 3, 2, 7, [63, 65, 64, 2], True),         # K not a multiple of 4, dropout on
    (20, 2, 9, 48, [200, 200], False),            # dilation up to 256 >= T
    (400, 4, 10, 48, [300, 257, 120], True),      # the reference's real shape, short videos
    (4, 1, 1, 64, [129], True),                   # one stage of one layer (no backward chain), the maximum class count
    (8, 2, 1, 1, [128, 128], False),              # L = 1, a single class, lengths exactly on the tile edge
    (8, 2, 2, 5, [1], False),                     # T = 1
    (8, 2, 3, 6, [40 - i for i in range(24)], True),   # many short videos (B = 24)
]


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("dim,S,L,K,lens,train", CASES)
def test_against_oracle(dim, S, L, K, lens, train, tensor_cores):
    from pytorch_video_action_b200 import MultiStageModel
    torch.manual_seed(7)
    net = MultiStageModel(dim, S, L, 64, K).cuda()
    net.tensor_cores = tensor_cores
    params = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
    B, T = len(lens), max(lens)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, T, dim)).astype(np.float32)
    y = rng.integers(0, K, size=(B, T)).astype(np.int64)
    for b, l in enumerate(lens):
        x[b, l:] = 0
        y[b, l:] = -1
    seed, off = 99, 5
    if train:
        net.train()
        net.set_dropout_state(seed, off)
        drop = lambda li, n: O.dropout_scale(seed, off, li, n)   # noqa: E731
    else:
        net.eval()
        drop = None
    from parity import adopt_kinks
    out, loss, grads = _run(net, x, lens, y.reshape(-1))
    ref_out, cache = O.forward(params, x, lens, train_dropout=drop)
    ref_loss, gout = O.cross_entropy(ref_out, y.reshape(-1))
    assert rel_err(out, ref_out) < TOL_REL
    assert abs(loss - float(ref_loss)) < TOL_LOSS
    # sub-gradient choices at ReLU / max kinks are taken from the run under test (tests/parity.py)
    relu = [[h.cpu().numpy() for h in st] for st in net.saved_relu_outputs()]
    winner = np.argmax(net.stage_logits().cpu().numpy(), axis=0)
    n_relu, n_win = adopt_kinks(cache, relu, winner, lens)
    assert n_relu <= 50 and n_win <= 50
    ref_grads = O.backward(cache, gout)
    errs = {k: rel_err(grads[k], ref_grads[k]) for k in ref_grads}
    worst = max(errs, key=errs.get)
    assert errs[worst] < TOL_REL, (worst, errs[worst])
    # frames beyond each video's length are exactly zero (every stage's logits are masked)
    o = out.reshape(B, T, K)
    for b, l in enumerate(lens):
        assert not o[b, l:].any()


def test_grad_accumulation_and_foreign_grads():
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
    torch.manual_seed(1)
    net = MultiStageModel(8, 2, 2, 64, 5).cuda().eval()
    x = torch.randn(2, 40, 8, device="cuda")
    y = torch.randint(0, 5, (80,), device="cuda")
    crit = FrameCrossEntropy()
    net.zero_grad()
    crit(net(x, [40, 40]), y).backward()
    g1 = [p.grad.clone() for p in net.parameters()]
    crit(net(x, [40, 40]), y).backward()                      # no zero_grad: must accumulate
    for a, p in zip(g1, net.parameters()):
        assert torch.allclose(p.grad, 2 * a, rtol=1e-5, atol=1e-7)
    for p in net.parameters():                                # foreign grad tensors: added into
        p.grad = torch.ones_like(p)
    crit(net(x, [40, 40]), y).backward()
    for a, p in zip(g1, net.parameters()):
        assert torch.allclose(p.grad, a + 1, rtol=1e-5, atol=1e-6)


def test_inference_ensemble_bit_exact():
    """inference.py:113-179 with .eval(): per-frame argmax, segment votes (both rules), ensemble mode."""
    from pytorch_video_action_b200 import frame_argmax, segment_vote, ensemble_vote
    g = load_golden("inference_ensemble")
    K = int(g["cfg"][4])
    nets = []
    for mi in range(2):
        params = {k[len(f"p{mi}/"):]: v for k, v in g.items() if k.startswith(f"p{mi}/")}
        nets.append(_model_from_params(params).eval())
    for vi in range(int(g["n_videos"])):
        x = torch.from_numpy(g[f"v{vi}/x"]).cuda()
        seg = g[f"v{vi}/segments"]
        per_model = []
        for mi, net in enumerate(nets):
            with torch.no_grad():
                out = net(x, [x.shape[1]])
            assert rel_err(out.cpu().numpy(), g[f"v{vi}/out{mi}"]) < TOL_REL
            _, pred = frame_argmax(out)
            if vi % 3 == 0:          # other videos carry forced predictions (see make_golden.py)
                assert np.array_equal(pred.cpu().numpy(), g[f"v{vi}/argmax{mi}"])
            pred = torch.from_numpy(g[f"v{vi}/argmax{mi}"]).cuda()
            dev_votes = segment_vote(pred, seg, K, inference_fallback=False).cpu().tolist()
            inf_votes = segment_vote(pred, seg, K, inference_fallback=True).cpu().tolist()
            assert dev_votes == list(g[f"v{vi}/vote_dev"][mi])
            assert inf_votes == list(g[f"v{vi}/vote_inf"][mi])
            per_model.append(inf_votes)
        assert ensemble_vote(per_model) == list(g[f"v{vi}/final"])


def test_evaluate_video_matches_reference_loop():
    from pytorch_video_action_b200 import evaluate_video
    rng = np.random.default_rng(0)
    K, n = 12, 500
    labels = np.repeat(rng.integers(0, K, 20), 25)
    out = rng.standard_normal((n, K)).astype(np.float32)
    out[np.arange(n), labels] += 1.5
    cf, tf, cs, ts = evaluate_video(torch.from_numpy(out).cuda(), torch.from_numpy(labels).cuda(), K)
    pred = O.frame_argmax(out)[1]
    seq, bounds = O.label_runs(labels)
    votes = O.segment_vote(pred, bounds)
    assert (cf, tf, cs, ts) == (int((pred == labels).sum()), n, sum(int(a == b) for a, b in zip(votes, seq)), len(seq))


def test_fused_adam_matches_torch():
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, FusedAdam
    torch.manual_seed(2)
    a = MultiStageModel(8, 2, 2, 64, 5).cuda().eval()
    b = MultiStageModel(8, 2, 2, 64, 5).cuda().eval()
    b.load_state_dict(a.state_dict())
    oa = FusedAdam(a, lr=1e-3)
    ob = torch.optim.Adam(b.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    crit = FrameCrossEntropy()
    x = torch.randn(2, 50, 8, device="cuda")
    y = torch.randint(0, 5, (100,), device="cuda")
    for _ in range(3):
        for net, opt in ((a, oa), (b, ob)):
            opt.zero_grad()
            crit(net(x, [50, 50]), y).backward()
            opt.step()
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7), k


def test_padded_batch_differs_from_solo_like_the_reference():
    """the unmasked stage-input conv makes a video's logits depend on whether a
    padded frame follows it.  The oracle (pinned to the reference) shows the same dependence."""
    from pytorch_video_action_b200 import MultiStageModel
    torch.manual_seed(4)
    net = MultiStageModel(16, 2, 3, 64, 6).cuda().eval()
    params = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 90, 16)).astype(np.float32)
    x[1, 50:] = 0
    with torch.no_grad():
        both = net(torch.from_numpy(x).cuda(), [90, 50]).cpu().numpy().reshape(2, 90, 6)
        solo = net(torch.from_numpy(x[1:2, :50].copy()).cuda(), [50]).cpu().numpy().reshape(50, 6)
    ref_both, _ = O.forward(params, x, [90, 50], keep_cache=False)
    ref_solo, _ = O.forward(params, x[1:2, :50], [50], keep_cache=False)
    assert rel_err(both.reshape(-1, 6), ref_both) < TOL_REL and rel_err(solo, ref_solo) < TOL_REL
    gap, ref_gap = np.abs(both[1, :50] - solo).max(), np.abs(ref_both.reshape(2, 90, 6)[1, :50] - ref_solo).max()
    assert gap > 1e-4 and abs(gap - ref_gap) < 1e-3 * max(1.0, ref_gap)


def test_error_behaviour():
    from pytorch_video_action_b200 import MultiStageModel
    net = MultiStageModel(8, 2, 2, 64, 5)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 10, 8), [10])                       # CPU tensor: no CPU path
    net = net.cuda()
    x = torch.zeros(2, 10, 8, device="cuda")
    with pytest.raises(IndexError):
        net(x, [10])                                           # networks.py:309
    with pytest.raises(RuntimeError):
        net(x, [9, 8])                                         # networks.py:333 broadcast mismatch
    with pytest.raises(RuntimeError):
        net(x.double(), [10, 10])
    with pytest.raises(NotImplementedError):
        MultiStageModel(8, 2, 2, 32, 5)


def test_full_size_properties():
    """BASELINE config 2 (B=8, T_pad=4000, D=400, 4x10x64, K=48, train mode): size-independent checks."""
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
    lens = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700]
    B, T, K = 8, 4000, 48
    torch.manual_seed(0)
    net = MultiStageModel(400, 4, 10, 64, K).cuda().train()
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, 400, generator=g)
    y = torch.randint(1, K, (B, T), generator=g)
    for b, l in enumerate(lens):
        x[b, l:] = 0
        y[b, l:] = -1
    x, y = x.cuda(), y.flatten().cuda()
    crit = FrameCrossEntropy()

    def step():
        net.set_dropout_state(11, 0)
        net.zero_grad()
        out = net(x, lens)
        loss = crit(out, y)
        loss.backward()
        return out.detach().clone(), float(loss), net.flat_parameters()[1].clone()

    o1, l1, g1 = step()
    o2, l2, g2 = step()
    assert torch.equal(o1, o2) and l1 == l2 and torch.equal(g1, g2)          # deterministic, bit for bit
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all()
    o = o1.view(B, T, K)
    for b, l in enumerate(lens):
        assert not o[b, l:].any()                                            # masked frames are exactly 0
    # the loss is a mean over valid frames: compare with torch's CE on the same logits
    ref = torch.nn.functional.cross_entropy(o1, y, ignore_index=-1)
    assert abs(l1 - float(ref)) < TOL_LOSS
    # video independence: dropping the other videos (keeping one padded frame, fact 0.5) changes nothing
    net.eval()
    with torch.no_grad():
        full = net(x, lens).view(B, T, K)
        sub = net._forward_impl(x[6:7, :1241].contiguous(), [1240], strict_len=False).view(1241, K)
    assert rel_err(sub[:1240].cpu().numpy(), full[6, :1240].cpu().numpy()) < 1e-5


def test_config4_long_video_matches_oracle():
    """BASELINE configs[3]: B=1, T=16384, D=2048, 4x10x64, dilation up to 512 (halo-heavy), fwd+bwd."""
    from pytorch_video_action_b200 import MultiStageModel
    from parity import adopt_kinks
    dim, S, L, K, T = 2048, 4, 10, 48, 16384
    torch.manual_seed(11)
    net = MultiStageModel(dim, S, L, 64, K).cuda().eval()
    params = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
    rng = np.random.default_rng(4)
    x = rng.standard_normal((1, T, dim)).astype(np.float32)
    y = np.repeat(rng.integers(1, K, T // 256), 256).astype(np.int64)
    out, loss, grads = _run(net, x, [T], y)
    ref_out, cache = O.forward(params, x, [T])
    ref_loss, gout = O.cross_entropy(ref_out, y)
    assert rel_err(out, ref_out) < TOL_REL
    assert abs(loss - float(ref_loss)) < TOL_LOSS
    # argmax: exact wherever the oracle's own top-2 margin exceeds the fp32 noise floor of a 2048-term contraction
    # (16384 frames x 48 classes always hold a few near-ties; on those no two fp32 implementations agree)
    am, ar = np.argmax(out, 1), np.argmax(ref_out, 1)
    top2 = np.sort(ref_out, axis=1)[:, -2:]
    differ = np.nonzero(am != ar)[0]
    assert len(differ) <= 4 and np.all(top2[differ, 1] - top2[differ, 0] < 5e-5), (len(differ), differ[:8])
    relu = [[h.cpu().numpy() for h in st] for st in net.saved_relu_outputs()]
    winner = np.argmax(net.stage_logits().cpu().numpy(), axis=0)
    n_relu, n_win = adopt_kinks(cache, relu, winner, [T])
    assert n_relu <= 200 and n_win <= 200
    ref_grads = O.backward(cache, gout)
    errs = {k: rel_err(grads[k], ref_grads[k]) for k in ref_grads}
    worst = max(errs, key=errs.get)
    assert errs[worst] < TOL_REL, (worst, errs[worst])


def test_config3_batch64_video_independence():
    """BASELINE configs[2] per-GPU work at G=1: 64 videos in one batch.  Every op is per-video, so a video's
    logits and the batch gradient must not depend on how the batch is cut (the same property the
    data-parallel sharding relies on)."""
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
    lens = sorted([4000, 3892, 3600, 3100, 2600, 2000, 1240, 700] * 8, reverse=True)
    B, T, K = len(lens), max(lens), 48
    torch.manual_seed(0)
    net = MultiStageModel(400, 4, 10, 64, K).cuda().eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, 400, generator=g)
    y = torch.randint(1, K, (B, T), generator=g)
    for b, l in enumerate(lens):
        x[b, l:] = 0
        y[b, l:] = -1
    x, y = x.cuda(), y.flatten().cuda()
    crit = FrameCrossEntropy()
    net.zero_grad()
    out = net(x, lens)
    crit(out, y).backward()
    g_full = net.flat_parameters()[1].clone()
    assert torch.isfinite(out).all() and torch.isfinite(g_full).all()
    # the same 64 videos as two half-batches whose gradients are accumulated (global divisor)
    n_valid = sum(lens)
    net.zero_grad()
    outs = []
    for lo, hi in ((0, 32), (32, 64)):
        ll = lens[lo:hi]
        Tl = T if max(ll) == T else max(ll) + 1                      # keep one padded frame (fact 0.5)
        o = net._forward_impl(x[lo:hi, :Tl].contiguous(), ll, strict_len=False)
        yy = y.view(B, T)[lo:hi, :Tl].contiguous().flatten()
        crit(o, yy, n_valid=n_valid).backward()
        outs.append((o.detach(), Tl))
    g_split = net.flat_parameters()[1]
    assert rel_err(g_split.cpu().numpy(), g_full.cpu().numpy()) < 1e-4
    o_full = out.detach().view(B, T, K)
    for (o, Tl), (lo, hi) in zip(outs, ((0, 32), (32, 64))):
        assert rel_err(o.view(hi - lo, Tl, K).cpu().numpy(), o_full[lo:hi, :Tl].cpu().numpy()) < 1e-5


def test_device_pad_batch_matches_reference_collate():
    """DeviceFeatureStore.pad_batch == the reference's pad_batch (train.py:183-205), bit for bit."""
    from pytorch_video_action_b200 import DeviceFeatureStore
    rng = np.random.default_rng(5)
    lens = [37, 128, 5, 260, 1, 64]
    feats = [rng.standard_normal((n, 16)).astype(np.float32) for n in lens]
    labs = [rng.integers(0, 9, n) for n in lens]
    store = DeviceFeatureStore(feats, labs)
    for idx in ([3, 0, 5], [4], [1, 1, 2, 0]):
        x, x_len, y = store.pad_batch(idx)
        xo, lo, yo = O.pad_batch([feats[i] for i in idx], [labs[i] for i in idx])
        assert x_len == lo
        assert np.array_equal(x.cpu().numpy(), xo) and np.array_equal(y.cpu().numpy(), yo)
    x, x_len, y, lens_dev = store.pad_batch([2, 0], pad_to=40, with_lens_tensor=True)      # a data-parallel shard's padding
    assert x.shape == (2, 40, 16) and lens_dev.tolist() == [5, 37]
    assert not x[0, 5:].any() and (y.view(2, 40)[0, 5:] == -1).all()
    with pytest.raises(ValueError):
        store.pad_batch([3], pad_to=100)
    with pytest.raises(RuntimeError):
        DeviceFeatureStore(feats, labs, device="cpu")


def _stage_logits_case(S, B, T, K, lens, seed):
    rng = np.random.default_rng(seed)
    z = (rng.standard_normal((S, B, T, K)) * 2.5).astype(np.float32)
    y = rng.integers(0, K, (B, T))
    for b, n in enumerate(lens):
        z[:, b, n:] = 0.0                   # per-stage logits are masked (networks.py:333)
        y[b, n:] = -1
    return z.reshape(S, B * T, K), y.reshape(-1)


@pytest.mark.parametrize("S,B,T,K,lens", [(4, 3, 50, 48, [50, 31, 1]), (1, 1, 1, 3, [1]), (2, 2, 130, 64, [130, 129])])
def test_fused_paper_loss_matches_torch_autograd(S, B, T, K, lens):
    """L2 (parity unpinned: no reference result exists): fused CE + T-MSE fwd+bwd vs the torch restatement in fp64."""
    from pytorch_video_action_b200 import MsTcnLoss
    from oracle import torch_port as TP
    z, y = _stage_logits_case(S, B, T, K, lens, 3)
    zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    ref = TP.ms_tcn_paper_loss(zt, torch.from_numpy(y), lens) if T > 1 else sum(
        torch.nn.functional.cross_entropy(zt[s], torch.from_numpy(y), ignore_index=-1) for s in range(S))
    ref.backward()
    zg = torch.tensor(z, device="cuda", requires_grad=True)
    crit = MsTcnLoss()
    loss = crit(zg, torch.from_numpy(y).cuda(), lens)
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))
    assert rel_err(zg.grad.cpu().numpy(), zt.grad.numpy()) < 1e-5


@pytest.mark.parametrize("tensor_cores", [True, False])
def test_forward_stages_with_paper_loss_end_to_end(tensor_cores):
    """forward_stages -> MsTcnLoss -> backward (per-stage gradients, no max routing) vs torch autograd over the CPU
    port of the reference with the same loss.  Parity unpinned (the loss is not in the reference); tolerance 1e-3."""
    from pytorch_video_action_b200 import MultiStageModel, MsTcnLoss
    from oracle import torch_port as TP
    dim, S, L, K = 16, 3, 4, 7
    lens = [150, 97]
    B, T = len(lens), max(lens)
    torch.manual_seed(11)
    net = MultiStageModel(dim, S, L, 64, K).eval()
    P = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    rng = np.random.default_rng(2)
    x = np.zeros((B, T, dim), np.float32)
    y = np.full((B, T), -1, np.int64)
    for b, n in enumerate(lens):
        x[b, :n] = rng.standard_normal((n, dim)); y[b, :n] = rng.integers(0, K, n)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y.reshape(-1))
    zs = TP.forward(P, xt, lens, S, L, K, train=False, per_stage=True)
    ref = TP.ms_tcn_paper_loss(zs, yt, lens)
    ref.backward()
    net = net.cuda()
    net.tensor_cores = tensor_cores
    out = net.forward_stages(xt.cuda(), lens)
    assert out.shape == (S, B * T, K)
    assert rel_err(out.detach().cpu().numpy(), zs.detach().numpy()) < 1e-3
    loss = MsTcnLoss()(out, yt.cuda(), lens)
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))
    for name, p in net.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), P[name].grad.numpy()) < 1e-3, name
    with torch.no_grad(), pytest.raises(RuntimeError):
        net.forward_stages(xt.cuda(), lens)


def test_graphed_train_step_equals_eager_step():
    """GraphedTrainStep (CUDA-graph replay; forward -> fused CE -> backward called directly, 1/n_valid handed to the
    backward as a device scalar) leaves the same loss and gradients as net(x) -> FrameCrossEntropy -> backward()."""
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep
    lens = [300, 211, 64]
    B, T, dim, K = len(lens), max(lens), 32, 11
    torch.manual_seed(3)
    net = MultiStageModel(dim, 3, 5, 64, K).cuda().eval()           # eval: no dropout, so both runs see the same model
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, dim, generator=g)
    y = torch.randint(0, K, (B, T), generator=g)
    for b, n in enumerate(lens):
        x[b, n:] = 0; y[b, n:] = -1
    x, y = x.cuda(), y.flatten().cuda()
    crit = FrameCrossEntropy()
    net.zero_grad()
    loss = crit(net(x, lens), y)
    loss.backward()
    ref = net.flat_parameters()[1].clone()
    xin, yin = x.clone(), y.clone()
    step = GraphedTrainStep(net, crit, lens, x, y, n_valid=sum(lens), inputs=[(xin, yin)])
    got = step.replay(0)                                             # a capture that reads its inputs in place
    torch.cuda.synchronize()
    assert float(got) == float(loss) and torch.equal(net.flat_parameters()[1], ref)
    for _ in range(2):                                               # replays are idempotent
        got = step(x, y)
        torch.cuda.synchronize()
        assert float(got) == float(loss)
        assert torch.equal(net.flat_parameters()[1], ref)
    x2 = x.clone(); x2[0, :50] += 1.0                                # new data through the static buffers
    l2 = step(x2, y)
    torch.cuda.synchronize()
    g2 = net.flat_parameters()[1].clone()
    assert not torch.equal(g2, ref)                                  # the new batch really went through
    net.zero_grad()
    l2_ref = crit(net(x2, lens), y); l2_ref.backward()
    assert float(l2) == float(l2_ref) and torch.equal(g2, net.flat_parameters()[1])



def test_fused_adam_steplr_and_checkpoint_roundtrip():
    """train.py:273-274,334-335: Adam + StepLR driven exactly as the reference does; the fused optimizer's state_dict
    loads into a fresh FusedAdam AND into torch.optim.Adam (same layout), and training continues identically."""
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, FusedAdam
    torch.manual_seed(5)
    nets = [MultiStageModel(8, 2, 2, 64, 5).cuda().eval() for _ in range(3)]
    for n in nets[1:]:
        n.load_state_dict(nets[0].state_dict())
    crit = FrameCrossEntropy()
    x = torch.randn(2, 50, 8, device="cuda")
    y = torch.randint(0, 5, (100,), device="cuda")

    def train(net, opt, sched, n):
        for _ in range(n):
            opt.zero_grad()
            crit(net(x, [50, 50]), y).backward()
            opt.step()
            sched.step()

    a, b, c = nets
    oa = FusedAdam(a, lr=1e-2)
    sa = torch.optim.lr_scheduler.StepLR(oa, step_size=2, gamma=0.5)
    ob = torch.optim.Adam(b.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    sb = torch.optim.lr_scheduler.StepLR(ob, step_size=2, gamma=0.5)
    train(a, oa, sa, 3)
    train(b, ob, sb, 3)
    assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"] == 5e-3
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        # same gradients bit for bit (the kernels are deterministic); the two Adam implementations round differently:
        # 4e-7 .. 6e-7 absolute after three steps at lr 1e-2 (tools/dbg_adam2.py), i.e. 6e-5 of one step's movement
        assert torch.allclose(pa, pb, rtol=1e-5, atol=5e-6), k
    # checkpoint: fused -> fresh fused (model c) and fused -> torch.optim.Adam (model b's optimizer)
    sd = oa.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and int(sd["state"][0]["step"]) == 3
    c.load_state_dict(a.state_dict())
    oc = FusedAdam(c, lr=1.0)
    oc.load_state_dict(sd)
    assert oc.param_groups[0]["lr"] == 5e-3 and oc.step_count == 3
    sc = torch.optim.lr_scheduler.StepLR(oc, step_size=2, gamma=0.5, last_epoch=-1)
    sc.last_epoch, sc._step_count = sa.last_epoch, sa._step_count
    import copy
    ob.load_state_dict(copy.deepcopy(sd))        # (a checkpoint goes through torch.save/load; load_state_dict itself aliases the tensors it is given)
    train(a, oa, sa, 2)
    train(c, oc, sc, 2)
    train(b, ob, sb, 2)
    for (k, pa), pb, pc in zip(a.named_parameters(), b.parameters(), c.parameters()):
        assert torch.equal(pa, pc), k
        # torch's foreach Adam and the fused kernel round differently; five steps at lr <= 1e-2 leave ~2e-6 (measured)
        assert torch.allclose(pa, pb, rtol=1e-5, atol=2e-5), k


def test_graphed_step_with_optimizer_in_the_graph():
    """GraphedTrainStep(optimizer=FusedAdam): the Adam update runs inside the captured graph with its step count and
    learning rate on the device; three replays with a StepLR change in between equal three eager
    zero_grad/forward/loss/backward/step iterations (train.py:305-335), and the checkpoint carries the step count."""
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, FusedAdam, GraphedTrainStep
    torch.manual_seed(11)
    a = MultiStageModel(16, 2, 3, 64, 7).cuda().eval()          # eval: no dropout, so both runs see the same function
    b = MultiStageModel(16, 2, 3, 64, 7).cuda().eval()
    b.load_state_dict(a.state_dict())
    crit = FrameCrossEntropy()
    lens = [150, 90]
    x = torch.randn(2, 150, 16, device="cuda")
    x[1, 90:] = 0
    y = torch.randint(0, 7, (2, 150), device="cuda")
    y[1, 90:] = -1
    y = y.flatten()
    oa, ob = FusedAdam(a, lr=1e-2), FusedAdam(b, lr=1e-2)
    sa = torch.optim.lr_scheduler.StepLR(oa, step_size=2, gamma=0.1)
    sb = torch.optim.lr_scheduler.StepLR(ob, step_size=2, gamma=0.1)
    w0 = b.flat_parameters()[0].clone()
    step = GraphedTrainStep(b, crit, lens, x, y, n_valid=sum(lens), optimizer=ob)
    assert torch.equal(b.flat_parameters()[0], w0) and ob.step_count == 0      # warm-up steps were rolled back
    for _ in range(3):
        oa.zero_grad()
        crit(a(x, lens), y).backward()
        oa.step()
        sa.step()
        step(x, y)
        sb.step()
    torch.cuda.synchronize()
    assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"] == pytest.approx(1e-3)
    assert ob.step_count == 3 and int(ob.state_dict()["state"][0]["step"]) == 3
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-7), k
    assert not torch.equal(b.flat_parameters()[0], w0)


def test_ragged_batch_uploader_builds_the_reference_collate():
    """RaggedBatchUploader: the batch's valid frames shipped as one ragged pinned block and padded on the device equals the
    reference's host-side pad_batch (train.py:183-205) bit for bit, also into fixed buffers with a longer pad length."""
    from pytorch_video_action_b200 import RaggedBatchUploader
    g = torch.Generator().manual_seed(3)
    lens, dim = [37, 64, 5, 20], 12
    feats = [torch.randn(n, dim, generator=g) for n in lens]
    labs = [torch.randint(0, 9, (n,), generator=g) for n in lens]
    rx, ry = torch.cat(feats).pin_memory(), torch.cat(labs).pin_memory()
    for T in (max(lens), max(lens) + 3):
        ref_x = torch.zeros(len(lens), T, dim)
        ref_y = torch.full((len(lens), T), -1, dtype=torch.int64)
        for b, n in enumerate(lens):
            ref_x[b, :n] = feats[b]
            ref_y[b, :n] = labs[b]
        up = RaggedBatchUploader(lens, dim, pad_to=T)
        x, y = up.upload(rx, ry)
        torch.cuda.synchronize()
        assert torch.equal(x.cpu(), ref_x) and torch.equal(y.cpu(), ref_y.flatten())
        xb = torch.full((len(lens), T, dim), 7.0, device="cuda")
        yb = torch.full((len(lens) * T,), 5, dtype=torch.int64, device="cuda")
        x2, y2 = up.upload(rx, ry, out=(xb, yb))
        torch.cuda.synchronize()
        assert x2 is xb and torch.equal(xb.cpu(), ref_x) and torch.equal(yb.cpu(), ref_y.flatten())
        assert up.h2d_bytes == sum(lens) * dim * 4 + sum(lens) * 8
    with pytest.raises(ValueError):
        up.upload(rx[:-1], ry)


def test_forward_mask_entry_point():
    """(x, mask) spelling of canonical MS-TCN: identical to forward(x, x_len) with the mask's lengths."""
    from pytorch_video_action_b200 import MultiStageModel
    torch.manual_seed(3)
    net = MultiStageModel(16, 2, 3, 64, 6).cuda().eval()
    lens = [90, 50, 1]
    x = torch.randn(3, 90, 16, device="cuda")
    mask = torch.zeros(3, 6, 90, device="cuda")
    for b, n in enumerate(lens):
        x[b, n:] = 0
        mask[b, :, :n] = 1
    with torch.no_grad():
        ref = net(x, lens)
        assert torch.equal(net.forward_mask(x, mask), ref)             # (B, K, T) mask as networks.py:307-309 builds it
        assert torch.equal(net.forward_mask(x, mask[:, 0, :]), ref)    # (B, T) mask


def test_two_models_on_two_streams_do_not_starve_each_other():
    """chain launches of two models on different streams of one GPU used to spin on each other until the
    bounded wait trapped.  The chain lane (mstcn_capi.cu) / ordered replays (graph.py) serialise them."""
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep
    lens = [1500, 1200, 900, 800, 640, 512, 300, 129] * 3           # 24 videos: > 148 tiles per layer, every SM busy
    B, T, dim, K = len(lens), max(lens), 32, 11
    torch.manual_seed(3)
    nets = [MultiStageModel(dim, 2, 6, 64, K).cuda().eval() for _ in range(2)]
    nets[1].load_state_dict(nets[0].state_dict())
    x = torch.randn(B, T, dim, device="cuda")
    y = torch.randint(0, K, (B * T,), device="cuda")
    for b, n in enumerate(lens):
        x[b, n:] = 0
        y.view(B, T)[b, n:] = -1
    crit = FrameCrossEntropy()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(5):
        for net, st in zip(nets, streams):
            with torch.cuda.stream(st):
                net.zero_grad()
                crit(net(x, lens), y).backward()
    torch.cuda.synchronize()
    assert torch.equal(nets[0].flat_parameters()[1], nets[1].flat_parameters()[1])
    # the same through graph replays on two streams
    steps = [GraphedTrainStep(n, crit, lens, x, y, n_valid=sum(lens)) for n in nets]
    for rep in range(5):
        for stp, st in zip(steps, streams):
            with torch.cuda.stream(st):
                stp(x, y)
    torch.cuda.synchronize()
    assert torch.equal(nets[0].flat_parameters()[1], nets[1].flat_parameters()[1])


def test_loss_label_contract():
    """only -1 is ignored; other out-of-range labels make the loss NaN (torch asserts); an all-ignored batch
    gives NaN like torch's empty mean; a second backward through the same graph is not double-scaled."""
    from pytorch_video_action_b200 import FrameCrossEntropy
    crit = FrameCrossEntropy()
    z = torch.randn(64, 7, device="cuda", requires_grad=True)
    y = torch.randint(0, 7, (64,), device="cuda")
    y[::5] = -1
    loss = crit(z, y)
    ref = torch.nn.functional.cross_entropy(z.detach(), y, ignore_index=-1)
    assert abs(float(loss) - float(ref)) < 1e-6
    g1, = torch.autograd.grad(loss, z, retain_graph=True)
    g2, = torch.autograd.grad(loss, z)
    assert torch.equal(g1, g2)
    ybad = y.clone(); ybad[1] = 7
    assert torch.isnan(crit(z, ybad))
    ybad[1] = -2
    assert torch.isnan(crit(z, ybad))
    assert torch.isnan(crit(z, torch.full_like(y, -1)))
    with pytest.raises(NotImplementedError):
        FrameCrossEntropy(ignore_index=-100)
