"""CPU tier: the C-ABI library loads and exports every symbol the header declares (no compute
calls without a GPU), and the host-side layout / planner logic."""
import ctypes as C

import numpy as np
import pytest
import torch


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    declared = _cabi.header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_cabi._SIGNATURES)
    assert lib.mstcn_abi_version() == 2


def test_param_layout_matches_state_dict_order():
    from pytorch_video_action_b200 import MultiStageModel, _cabi
    lib = _cabi.lib()
    for dim, S, L, K in [(400, 4, 10, 48), (400, 4, 20, 2), (16, 1, 3, 5)]:
        net = MultiStageModel(dim, S, L, 64, K)
        d = _cabi.MstcnDims(dim, S, L, 64, K)
        acc = 0
        params = list(net.named_parameters())
        assert lib.mstcn_param_tensors(C.byref(d)) == len(params)
        for i, (_, p) in enumerate(params):
            assert lib.mstcn_param_offset(C.byref(d), i) == acc
            acc += p.numel()
        assert lib.mstcn_param_count(C.byref(d)) == acc
        b = net.bucket_boundaries()
        assert b[0] == 0 and b[-1] == acc and b == sorted(b) and len(b) == S + 2
    d = _cabi.MstcnDims(400, 4, 10, 64, 48)
    assert lib.mstcn_param_count(C.byref(d)) == 708032
    assert lib.mstcn_param_count(C.byref(_cabi.MstcnDims(400, 4, 20, 64, 48))) == 1368512


def test_rejects_unsupported_dims_loudly():
    from pytorch_video_action_b200 import MultiStageModel, _cabi
    lib = _cabi.lib()
    assert lib.mstcn_param_count(C.byref(_cabi.MstcnDims(400, 4, 10, 32, 48))) == -1
    assert b"num_f_maps" in lib.mstcn_last_error()
    for bad in [dict(num_f_maps=32), dict(n_class=65), dict(dim=401)]:
        with pytest.raises(NotImplementedError):
            MultiStageModel(**{**dict(dim=400, num_stages=2, num_layers=2, num_f_maps=64, n_class=4), **bad})


def test_same_seed_init_and_state_dict_keys():
    from pytorch_video_action_b200 import MultiStageModel
    torch.manual_seed(0)
    a = MultiStageModel(400, n_class=48)                          # train.py:252 call shape
    keys = list(a.state_dict().keys())
    assert len(keys) == 4 * (4 + 4 * 20)
    assert keys[0] == "stage1.conv_1x1.weight" and keys[-1] == "stages.2.conv_out.bias"
    assert a.state_dict()["stage1.layers.3.conv_dilated.weight"].shape == (64, 64, 3)
    assert a.state_dict()["stages.0.conv_1x1.weight"].shape == (64, 48, 1)
    g = __import__("conftest").load_golden("small_eval")
    # golden weights were drawn by the reference under manual_seed(0): same ctor order -> same values
    torch.manual_seed(0)
    b = MultiStageModel(24, 3, 4, 64, 48)
    for k, v in b.state_dict().items():
        assert np.array_equal(v.numpy(), g["p/" + k]), k


def test_cpu_input_raises_instead_of_falling_back():
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, frame_argmax
    net = MultiStageModel(8, 2, 2, 64, 5)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 10, 8), [10])
    with pytest.raises(RuntimeError):
        FrameCrossEntropy()(torch.zeros(4, 5), torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError):
        frame_argmax(torch.zeros(4, 5))


def test_shard_planner():
    from pytorch_video_action_b200.parallel import shard_videos, local_pad_length
    lens = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700] * 8
    for w in (1, 2, 4, 8):
        shards = shard_videos(lens, w)
        assert sorted(i for s in shards for i in s) == list(range(64))
        sums = [sum(lens[i] for i in s) for s in shards]
        assert max(sums) - min(sums) <= 0.02 * max(sums)
        assert all(len(s) == 64 // w for s in shards)
    assert local_pad_length([100, 50], 100) == 100
    assert local_pad_length([90, 50], 100) == 91


def test_ensemble_vote_host_logic():
    from pytorch_video_action_b200.postprocess import ensemble_vote
    from oracle import mstcn_oracle as O
    votes = [[3, 0, 5, 0], [4, 0, 5, 7], [4, 0, 2, 0]]
    assert ensemble_vote(votes) == O.ensemble_vote(votes) == [4, 0, 5, 7]
    assert ensemble_vote([[1], [2]]) == [1]                      # first-seen wins a tie (Python >= 3.8 mode)


def test_fused_adam_is_a_torch_optimizer():
    """train.py:273-274: StepLR(optimizer, ...) must accept the fused optimizer (it is a torch.optim.Optimizer with
    one param group whose lr the scheduler rewrites)."""
    from pytorch_video_action_b200 import MultiStageModel, FusedAdam
    net = MultiStageModel(8, 2, 2, 64, 5)
    opt = FusedAdam(net, lr=1e-3)
    assert isinstance(opt, torch.optim.Optimizer)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    assert len(opt.param_groups) == 1 and len(opt.param_groups[0]["params"]) == len(list(net.parameters()))
    for _ in range(2):
        sched.step()
    assert abs(opt.param_groups[0]["lr"] - 5e-4) < 1e-12
    with pytest.raises(TypeError):
        FusedAdam(net.parameters())
    sd = opt.state_dict()
    assert sd["param_groups"][0]["betas"] == (0.9, 0.999) and sd["param_groups"][0]["eps"] == 1e-8


def test_grad_bucket_reducer_follows_the_current_buffer():
    """the reducer must sum the CURRENT flat gradient buffer, not one captured at construction."""
    from pytorch_video_action_b200.parallel import GradBucketReducer
    bufs = [torch.zeros(10), torch.ones(10)]
    cur = [0]
    red = GradBucketReducer(lambda: bufs[cur[0]], [0, 4, 10], overlap=True)
    assert red.bucket(1).data_ptr() == bufs[0][4:].data_ptr()
    cur[0] = 1
    assert red.bucket(1).data_ptr() == bufs[1][4:].data_ptr() and red.bucket(0).numel() == 4
    red2 = GradBucketReducer(bufs[0], [0, 4, 10])              # default: one late all-reduce of the whole buffer
    assert not red2.overlap and red2.bucket(0).numel() == 10
