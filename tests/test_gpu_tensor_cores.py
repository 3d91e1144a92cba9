"""GPU tier: the tcgen05 / TMEM / TMA layer kernel (error-compensated 3xTF32) against the exact fp32
FFMA kernel, the numpy oracle and the reference's golden vectors.  Same north-star tolerances."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden, split_golden, rel_err
from oracle import mstcn_oracle as O

pytestmark = pytest.mark.gpu


def _packed_ptr(net, stage, layer, which):
    from pytorch_video_action_b200 import _cabi
    off = _cabi.lib().mstcn_packed_offset(C.byref(net._dims), stage, layer, which)
    assert off >= 0
    return C.c_void_p(net._packed.data_ptr() + 4 * off)


@pytest.mark.parametrize("B,T,lens,d,train", [
    (1, 128, [128], 1, False),
    (2, 300, [300, 131], 4, False),
    (3, 257, [257, 128, 5], 64, True),
    (2, 1000, [1000, 640], 512, True),
    (1, 97, [97], 128, False),            # both side taps entirely out of range (d >= T)
])
def test_tc_layer_forward_matches_fp32_kernel(B, T, lens, d, train):
    from pytorch_video_action_b200 import MultiStageModel, _cabi
    lib = _cabi.lib()
    torch.manual_seed(0)
    net = MultiStageModel(16, 2, 3, 64, 8).cuda()
    net.tensor_cores = True
    net.pack_ffma_operands = True      # these tests also call the FFMA kernels with this model's packed operands
    with torch.no_grad():
        net(torch.zeros(1, 8, 16, device="cuda"), [8])           # packs fp32 operands + tensor-core images
    torch.manual_seed(1)
    x = torch.randn(B, T, 64, device="cuda") * 1.7
    lens_dev = torch.tensor(lens, dtype=torch.int32, device="cuda")
    drop = _cabi.MstcnDropout(1 if train else 0, 0, 77, 3)
    st = _cabi.stream_ptr()
    y0, h0 = torch.full_like(x, 9.0), torch.full_like(x, 9.0)
    y1, h1 = torch.full_like(x, 7.0), torch.full_like(x, 7.0)
    s, l = 1, 2
    _cabi.check(lib.mstcn_layer_fwd(_cabi.ptr(x), _cabi.ptr(y0), _cabi.ptr(h0), _cabi.ptr(lens_dev), B, T, d,
                                    _packed_ptr(net, s, l, 3), _packed_ptr(net, s, l, 4), _packed_ptr(net, s, l, 5),
                                    _packed_ptr(net, s, l, 6), C.byref(drop), 5, st))
    _cabi.check(lib.mstcn_layer_fwd_tc(_cabi.ptr(x), _cabi.ptr(y1), _cabi.ptr(h1), _cabi.ptr(lens_dev), B, T, d,
                                       _packed_ptr(net, s, l, 12), _packed_ptr(net, s, l, 4), _packed_ptr(net, s, l, 6),
                                       C.byref(drop), 5, st))
    torch.cuda.synchronize()
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 2e-5
    for b, n in enumerate(lens):                                   # h is only defined on tiles that hold valid frames
        hi = min(T, (n + 127) // 128 * 128)
        hi0 = min(T, (n + 63) // 64 * 64)
        assert rel_err(h1[b, :hi0].cpu().numpy(), h0[b, :hi0].cpu().numpy()) < 2e-5
        assert not y1[b, n:].any()
        del hi


@pytest.mark.parametrize("name", ["small_eval", "small_train", "deep_d_ge_T"])
def test_tc_model_matches_reference_golden(name):
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, frame_argmax
    g = load_golden(name)
    params, ref_grads = split_golden(g)
    dim, S, L, Cc, K = O.infer_config(params)
    net = MultiStageModel(dim, S, L, Cc, K)
    net.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in params.items()})
    net = net.cuda()
    net.tensor_cores = True
    net.pack_ffma_operands = True      # these tests also call the FFMA kernels with this model's packed operands
    seed, off = (int(v) for v in g["dropout"])
    if seed >= 0:
        net.train(); net.set_dropout_state(seed, off)
    else:
        net.eval()
    out = net(torch.from_numpy(g["x"]).cuda(), [int(v) for v in g["lens"]])
    loss = FrameCrossEntropy()(out, torch.from_numpy(g["y"]).cuda())
    loss.backward()
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < 1e-3
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4
    errs = {k: rel_err(p.grad.cpu().numpy(), ref_grads[k]) for k, p in net.named_parameters()}
    worst = max(errs, key=errs.get)
    assert errs[worst] < 1e-3, (worst, errs[worst])
    assert np.array_equal(frame_argmax(out)[1].cpu().numpy(), g["argmax"])
    print(f"{name}: logits rel {rel_err(out.detach().cpu().numpy(), g['out']):.2e} worst grad {worst} {errs[worst]:.2e}")


@pytest.mark.parametrize("B,T,lens,d", [
    (1, 128, [128], 1),
    (2, 300, [300, 131], 4),
    (3, 257, [257, 128, 5], 64),
    (2, 1000, [1000, 640], 512),
    (1, 97, [97], 128),
])
def test_tc_input_gradient_matches_fp32_kernel(B, T, lens, d):
    """gx from the tcgen05 kernel vs the FFMA layer backward on the same gu / gy."""
    from pytorch_video_action_b200 import MultiStageModel, _cabi
    lib = _cabi.lib()
    torch.manual_seed(0)
    net = MultiStageModel(16, 2, 3, 64, 8).cuda()
    net.tensor_cores = True
    net.pack_ffma_operands = True      # these tests also call the FFMA kernels with this model's packed operands
    with torch.no_grad():
        net(torch.zeros(1, 8, 16, device="cuda"), [8])
    torch.manual_seed(2)
    x = torch.randn(B, T, 64, device="cuda")
    h = torch.relu(torch.randn(B, T, 64, device="cuda"))
    gy = torch.randn(B, T, 64, device="cuda")
    lens_dev = torch.tensor(lens, dtype=torch.int32, device="cuda")
    drop = _cabi.MstcnDropout(0, 0, 0, 0)
    st = _cabi.stream_ptr()
    s, l = 1, 1
    gx0, gu = torch.full_like(x, 3.0), torch.full_like(x, 3.0)
    gw = [torch.zeros(64 * 64 * 3, device="cuda"), torch.zeros(64, device="cuda"),
          torch.zeros(64 * 64, device="cuda"), torch.zeros(64, device="cuda")]
    scratch = torch.empty(lib.mstcn_layer_bwd_scratch_floats(), device="cuda")
    _cabi.check(lib.mstcn_layer_bwd(_cabi.ptr(x), _cabi.ptr(h), _cabi.ptr(gy), _cabi.ptr(gx0), _cabi.ptr(gu),
                                    _cabi.ptr(lens_dev), B, T, d, _packed_ptr(net, s, l, 7), _packed_ptr(net, s, l, 8),
                                    C.byref(drop), 0, _cabi.ptr(gw[0]), _cabi.ptr(gw[1]), _cabi.ptr(gw[2]),
                                    _cabi.ptr(gw[3]), _cabi.ptr(scratch), 0, st))
    gx1 = torch.full_like(x, 5.0)
    _cabi.check(lib.mstcn_layer_bwd_gx_tc(_cabi.ptr(gu), _cabi.ptr(gy), _cabi.ptr(gx1), _cabi.ptr(lens_dev), B, T, d,
                                          _packed_ptr(net, s, l, 13), st))
    torch.cuda.synchronize()
    assert rel_err(gx1.cpu().numpy(), gx0.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("B,T,lens,L,train", [
    (2, 300, [300, 131], 3, False),
    (3, 1100, [1100, 640, 5], 10, True),       # dilations up to 512: taps reach over several tiles, many padding tiles
    (20, 260, [260 - 7 * i for i in range(20)], 4, True),   # more tasks per layer than CTAs on the GPU
    (1, 97, [97], 1, False),                    # a chain of one layer
])
def test_stage_chain_launch_is_bit_identical_to_per_layer_launches(B, T, lens, L, train):
    """mstcn_stage_fwd_tc (all layers of a stage, tile-level dataflow inside one launch) == L x mstcn_layer_fwd_tc."""
    from pytorch_video_action_b200 import MultiStageModel, _cabi
    lib = _cabi.lib()
    torch.manual_seed(0)
    net = MultiStageModel(16, 2, L, 64, 8).cuda()
    net.tensor_cores = True
    net.pack_ffma_operands = True      # these tests also call the FFMA kernels with this model's packed operands
    with torch.no_grad():
        net(torch.zeros(1, 8, 16, device="cuda"), [8])
    torch.manual_seed(2)
    N = B * T
    lens_dev = torch.tensor(lens, dtype=torch.int32, device="cuda")
    drop = _cabi.MstcnDropout(1 if train else 0, 0, 99, 11)
    st = _cabi.stream_ptr()
    s = 1
    x0 = torch.randn(N, 64, device="cuda")
    for b, n in enumerate(lens):
        x0[b * T + n:(b + 1) * T] = 0.3                     # the stage input is not masked (bias on padded frames)
    ref = torch.full(((L + 1) * N, 64), 5.0, device="cuda"); ref[:N] = x0
    href = torch.full((L * N, 64), 5.0, device="cuda")
    for l in range(L):
        _cabi.check(lib.mstcn_layer_fwd_tc(_cabi.ptr(ref[l * N:]), _cabi.ptr(ref[(l + 1) * N:]), _cabi.ptr(href[l * N:]),
                                           _cabi.ptr(lens_dev), B, T, 1 << l, _packed_ptr(net, s, l, 12),
                                           _packed_ptr(net, s, l, 4), _packed_ptr(net, s, l, 6), C.byref(drop), s * L + l, st))
    got = torch.full(((L + 1) * N, 64), -5.0, device="cuda"); got[:N] = x0
    hgot = torch.full((L * N, 64), -5.0, device="cuda")
    flags = torch.full((L * B * ((T + 127) // 128),), 7, dtype=torch.int32, device="cuda")    # stale flags must not matter
    for _ in range(2):                                       # second launch: flags are cleared by the entry point itself
        _cabi.check(lib.mstcn_stage_fwd_tc(C.byref(net._dims), _cabi.ptr(net._packed), s, _cabi.ptr(got), _cabi.ptr(hgot),
                                           _cabi.ptr(lens_dev), B, T, C.byref(drop), _cabi.ptr(flags), st))
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    for l in range(L):                                       # h is written on tiles that hold valid frames only
        for b, n in enumerate(lens):
            hi = min(T, (n + 127) // 128 * 128)
            assert torch.equal(hgot[l * N + b * T: l * N + b * T + hi], href[l * N + b * T: l * N + b * T + hi])


@pytest.mark.parametrize("B,T,lens,dim", [(2, 300, [300, 131], 16), (3, 257, [257, 1, 129], 400), (1, 128, [128], 2048),
                                           (5, 130, [130, 2, 2, 2, 2], 36)])
def test_tc_projection_matches_fp32_kernel(B, T, lens, dim):
    """mstcn_proj_fwd_tc (tcgen05, 3xTF32, features read in place by TMA) vs the exact fp32 FFMA projection."""
    from pytorch_video_action_b200 import MultiStageModel, _cabi
    lib = _cabi.lib()
    torch.manual_seed(0)
    net = MultiStageModel(dim, 2, 2, 64, 8).cuda()
    net.tensor_cores = True
    net.pack_ffma_operands = True      # these tests also call the FFMA kernels with this model's packed operands
    with torch.no_grad():
        net(torch.zeros(1, 8, dim, device="cuda"), [8])
    torch.manual_seed(4)
    x = torch.randn(B, T, dim, device="cuda") * 1.3
    for b, n in enumerate(lens):
        x[b, n:] = 0                                    # pad_batch zero-pads the features
    lens_dev = torch.tensor(lens, dtype=torch.int32, device="cuda")
    y0 = torch.full((B * T, 64), 9.0, device="cuda")
    y1 = torch.full((B * T, 64), 7.0, device="cuda")
    y2 = torch.full((B * T, 64), 5.0, device="cuda")
    st = _cabi.stream_ptr()
    _cabi.check(lib.mstcn_proj_fwd(_cabi.ptr(x), B * T, dim, _packed_ptr(net, 0, 0, 0), _packed_ptr(net, 0, 0, 1), _cabi.ptr(y0), st))
    _cabi.check(lib.mstcn_proj_fwd_tc(_cabi.ptr(x), B * T, dim, _packed_ptr(net, 0, 0, 14), _packed_ptr(net, 0, 0, 1),
                                      _cabi.ptr(lens_dev), T, _cabi.ptr(y1), st))
    _cabi.check(lib.mstcn_proj_fwd_tc(_cabi.ptr(x), B * T, dim, _packed_ptr(net, 0, 0, 14), _packed_ptr(net, 0, 0, 1),
                                      None, 0, _cabi.ptr(y2), st))
    torch.cuda.synchronize()
    assert rel_err(y1.cpu().numpy(), y0.cpu().numpy()) < 2e-5
    assert rel_err(y2.cpu().numpy(), y0.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("B,T,lens,dim", [(2, 300, [300, 131], 16), (3, 257, [257, 1, 129], 400), (1, 640, [640], 2048),
                                           (5, 130, [130, 2, 2, 2, 2], 36), (4, 1000, [1000, 700, 64, 63], 400)])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_tc_projection_weight_gradient_matches_fp32_kernel_and_fp64(B, T, lens, dim, accumulate):
    """mstcn_proj_wgrad_tc (tc_wgrad_kernel in projection mode: features read in place by TMA in 64-feature chunks, exact
    4-term tf32) vs the fp32 FFMA kernel and a float64 contraction: dW = g^T x and db = sum_t g over ALL frames -- the
    stage-1 conv is unmasked (networks.py:330), so the gradient rows beyond x_len count (their features are zero, their g is not)."""
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    torch.manual_seed(11)
    x = torch.randn(B, T, dim, device="cuda") * 1.3
    for b, n in enumerate(lens):
        x[b, n:] = 0
    g = torch.randn(B * T, 64, device="cuda") * 0.7
    lens_dev = torch.tensor(lens, dtype=torch.int32, device="cuda")
    st = _cabi.stream_ptr()
    init_w, init_b = torch.randn(64, dim, device="cuda"), torch.randn(64, device="cuda")
    gw0, gb0, gw1, gb1 = init_w.clone(), init_b.clone(), init_w.clone(), init_b.clone()
    s0 = torch.empty(lib.mstcn_proj_bwd_scratch_floats(dim), device="cuda")
    s1 = torch.full((lib.mstcn_proj_wgrad_tc_scratch_floats(dim),), float("nan"), device="cuda")
    _cabi.check(lib.mstcn_proj_bwd(_cabi.ptr(x), _cabi.ptr(g), B * T, dim, _cabi.ptr(gw0), _cabi.ptr(gb0), _cabi.ptr(s0), accumulate, st))
    _cabi.check(lib.mstcn_proj_wgrad_tc(_cabi.ptr(x), _cabi.ptr(g), _cabi.ptr(lens_dev), B, T, dim, _cabi.ptr(gw1), _cabi.ptr(gb1),
                                        _cabi.ptr(s1), accumulate, st))
    torch.cuda.synchronize()
    ref_w = (g.double().t() @ x.view(B * T, dim).double()) + (init_w.double() if accumulate else 0)
    ref_b = g.double().sum(0) + (init_b.double() if accumulate else 0)
    assert rel_err(gw1.cpu().numpy(), ref_w.cpu().numpy()) < 2e-6
    assert rel_err(gb1.cpu().numpy(), ref_b.cpu().numpy()) < 2e-6
    assert rel_err(gw1.cpu().numpy(), gw0.cpu().numpy()) < 1e-5
    assert rel_err(gb1.cpu().numpy(), gb0.cpu().numpy()) < 1e-5
