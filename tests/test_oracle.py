"""CPU tier: the numpy oracle replayed against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  Tolerances: 1e-3 relative for logits/grads and
1e-4 absolute for the loss are the north-star bars; the oracle is held to 2e-5 / 1e-5."""
import numpy as np
import pytest

from conftest import load_golden, split_golden, rel_err
from oracle import mstcn_oracle as O


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = O.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_dropout_scale_statistics():
    s = O.dropout_scale(123, 0, 3, 4096)
    assert set(np.unique(s)) == {0.0, 2.0}
    assert abs(s.mean() - 1.0) < 0.02
    assert not np.array_equal(s, O.dropout_scale(123, 1, 3, 4096))
    assert not np.array_equal(s, O.dropout_scale(123, 0, 4, 4096))


@pytest.mark.parametrize("name", ["small_eval", "small_train", "deep_d_ge_T"])
def test_forward_backward_matches_reference(name):
    g = load_golden(name)
    params, grads = split_golden(g)
    lens = [int(v) for v in g["lens"]]
    seed, off = (int(v) for v in g["dropout"])
    drop = None if seed < 0 else (lambda li, n: O.dropout_scale(seed, off, li, n))
    out, cache = O.forward(params, g["x"], lens, train_dropout=drop, dtype=np.float32)
    assert rel_err(out, g["out"]) < 2e-5
    loss, gout = O.cross_entropy(out, g["y"])
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    mine = O.backward(cache, gout)
    assert set(mine) == set(grads)
    worst = max(rel_err(mine[k], grads[k]) for k in grads)
    assert worst < 2e-4, worst
    _, idx = O.frame_argmax(out)
    assert np.array_equal(idx, g["argmax"])


def test_float64_oracle_agrees():
    g = load_golden("small_eval")
    params, grads = split_golden(g)
    lens = [int(v) for v in g["lens"]]
    out, cache = O.forward(params, g["x"], lens, dtype=np.float64)
    assert rel_err(out, g["out"]) < 1e-5
    _, gout = O.cross_entropy(out, g["y"])
    mine = O.backward(cache, gout)
    assert max(rel_err(mine[k], grads[k]) for k in grads) < 1e-4


def test_votes_match_reference_snippets():
    g = load_golden("inference_ensemble")
    nv = int(g["n_videos"])
    for vi in range(nv):
        seg = g[f"v{vi}/segments"]
        per_model = []
        for mi in range(2):
            pred = g[f"v{vi}/argmax{mi}"]
            assert O.segment_vote(pred, seg, False) == list(g[f"v{vi}/vote_dev"][mi])
            inf = O.segment_vote(pred, seg, True)
            assert inf == list(g[f"v{vi}/vote_inf"][mi])
            # where the reference's unstable argsort has no tie at position [1] it must agree too
            free = g[f"v{vi}/vote_inf_tiefree{mi}"]
            unst = g[f"v{vi}/vote_inf_unstable{mi}"]
            assert all(a == b for a, b, f in zip(inf, unst, free) if f)
            per_model.append(inf)
        assert O.ensemble_vote(per_model) == list(g[f"v{vi}/final"])


def test_inference_argmax_matches_reference():
    g = load_golden("inference_ensemble")
    for mi in range(2):
        params = {k[len(f"p{mi}/"):]: v for k, v in g.items() if k.startswith(f"p{mi}/")}
        for vi in range(0, int(g["n_videos"]), 3):          # other videos carry forced predictions
            x = g[f"v{vi}/x"]
            out, _ = O.forward(params, x, [x.shape[1]], keep_cache=False)
            assert rel_err(out, g[f"v{vi}/out{mi}"]) < 2e-5
            assert np.array_equal(O.frame_argmax(out)[1], g[f"v{vi}/argmax{mi}"])


def test_label_runs_and_pad_batch():
    seq, bounds = O.label_runs([3, 3, 3, 5, 5, 1])
    assert seq == [3, 5, 1] and bounds == [0, 3, 5, 6]
    x, lens, y = O.pad_batch([np.ones((3, 4), np.float32), np.ones((5, 4), np.float32)],
                             [np.array([1, 1, 2]), np.array([4, 4, 4, 4, 4])])
    assert x.shape == (2, 5, 4) and lens == [3, 5]
    assert list(y[:5]) == [1, 1, 2, -1, -1] and float(x[0, 3:].sum()) == 0.0


def test_adam_matches_torch():
    import torch
    torch.manual_seed(0)
    p = torch.randn(257, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    pn, m, v = p.detach().numpy().astype(np.float64), np.zeros(257), np.zeros(257)
    for step in range(1, 4):
        g = torch.randn(257)
        p.grad = g.clone()
        opt.step()
        pn, m, v = O.adam_step(pn, g.numpy().astype(np.float64), m, v, step)
    assert np.abs(pn - p.detach().numpy()).max() < 1e-6


@pytest.mark.parametrize("name", ["small_eval", "deep_d_ge_T"])
def test_torch_port_matches_reference(name):
    """The CPU-baseline port (oracle/torch_port.py) reproduces the unmodified reference's outputs."""
    import torch
    from oracle import torch_port as TP
    g = load_golden(name)
    params, grads = split_golden(g)
    dim, S, L, C, K = (int(v) for v in g["cfg"])
    P = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in params.items()}
    out, loss = TP.train_step(P, torch.from_numpy(g["x"]), [int(v) for v in g["lens"]],
                              torch.from_numpy(g["y"]), S, L, K, train=False)
    assert rel_err(out.detach().numpy(), g["out"]) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert max(rel_err(P[k].grad.numpy(), grads[k]) for k in grads) < 1e-4
    # same-seed init equals the reference's (and hence the golden weights drawn under manual_seed)
    sd = TP.make_params(dim, S, L, K, seed=0 if name == "small_eval" else 3)
    assert all(np.array_equal(sd[k].numpy(), params[k]) for k in params)


def test_paper_loss_numpy_matches_torch_restatement():
    """L2 (canonical MS-TCN loss) is NOT in the reference (parity unpinned): its two restatements -- numpy and torch
    autograd, written from the paper's formula -- must at least agree with each other."""
    import torch
    from oracle import torch_port as TP
    torch.manual_seed(0)
    S, B, T, K = 3, 2, 9, 5
    z = torch.randn(S, B * T, K, dtype=torch.float64) * 3
    lens = [9, 6]
    y = torch.randint(0, K, (B * T,))
    y.view(B, T)[1, 6:] = -1
    a = float(TP.ms_tcn_paper_loss(z, y, lens))
    b = O.ms_tcn_paper_loss(z.numpy().reshape(S, B, T, K), y.numpy(), lens)
    assert abs(a - b) < 1e-12


def test_oracle_matches_reference_at_config2_full_size():
    """The oracle pinned at the HEADLINE shape: BASELINE configs[1] (B=8, T_pad=4000, D=400, 4x10x64, K=48, train mode
    with the Philox mask injected) against the unmodified reference's loss, logits, argmax and all 176 gradients
    (tests/golden/config2_train.npz, written by make_golden.py --full-size)."""
    from conftest import synth_config2, reference_init_params, CONFIG2_LENS
    g = load_golden("config2_train")
    dim, S, L, _, K = (int(v) for v in g["cfg"])
    _, params = reference_init_params(dim, S, L, K, int(g["wseed"]))
    x, y = synth_config2(int(g["xseed"]))
    seed, off = (int(v) for v in g["dropout"])
    out, cache = O.forward(params, x.numpy(), CONFIG2_LENS, train_dropout=lambda li, n: O.dropout_scale(seed, off, li, n))
    assert rel_err(out[g["out_rows"]], g["out_sample"]) < 2e-5
    loss, gout = O.cross_entropy(out, y.numpy())
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    am = np.argmax(out, 1)
    differ = np.nonzero(am != g["argmax"])[0]
    assert set(differ) <= set(g["near_tie_rows"][g["near_tie_margin"] < 2e-5]), differ[:8]
    # sub-gradient choices at ReLU / stage-max kinks: the reference's own, recorded in the fixture (tests/parity.py)
    from parity import adopt_recorded_kinks
    n_relu, n_win = adopt_recorded_kinks(cache, g)
    print(f"config 2: {n_relu} ReLU and {n_win} stage-max choices adopted from the reference (all on true kinks)")
    assert n_relu <= 200 and n_win <= 200
    mine = O.backward(cache, gout)
    _, grads = split_golden(g)
    assert set(mine) == set(grads) and len(grads) == 176
    errs = {k: rel_err(mine[k], grads[k]) for k in grads}
    worst = max(errs, key=errs.get)
    assert errs[worst] < 1e-3, (worst, errs[worst])
