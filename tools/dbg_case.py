import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
from oracle import mstcn_oracle as O
dim, S, L, K, lens, train = 400, 4, 10, 48, [300, 257, 120], True
if len(sys.argv) > 1 and sys.argv[1] == "eval": train = False
def run(flags):
    torch.manual_seed(7)
    net = MultiStageModel(dim, S, L, 64, K).cuda()
    net._dims.flags = flags
    params = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
    B, T = len(lens), max(lens)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, T, dim)).astype(np.float32)
    y = rng.integers(0, K, size=(B, T)).astype(np.int64)
    for b, l in enumerate(lens):
        x[b, l:] = 0; y[b, l:] = -1
    seed, off = 99, 5
    if train: net.train(); net.set_dropout_state(seed, off)
    else: net.eval()
    net.zero_grad()
    out = net(torch.from_numpy(x).cuda(), lens)
    loss = FrameCrossEntropy()(out, torch.from_numpy(y.reshape(-1)).cuda())
    loss.backward()
    return params, x, y, out.detach().cpu().numpy(), {k: p.grad.cpu().numpy() for k, p in net.named_parameters()}, net.stage_logits().cpu().numpy()
params, x, y, o0, g0, sl0 = run(0)
drop = (lambda li, n: O.dropout_scale(99, 5, li, n)) if train else None
ref_out, cache = O.forward(params, x, lens, train_dropout=drop, dtype=np.float64)
_, gout = O.cross_entropy(ref_out, y.reshape(-1))
gref = O.backward(cache, gout)
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
for name, flags in [("ffma", 0), ("tc fwd + ffma bwd", 3), ("tc all", 1)]:
    _, _, _, o, g, sl = run(flags)
    errs = {k: rel(g[k], gref[k]) for k in gref}
    worst = sorted(errs, key=errs.get)[-3:]
    win = np.argmax(sl, axis=0); wref = cache["winner"].reshape(win.shape)
    print(f"{name:22s} logits {rel(o, ref_out):.2e}  winner flips {(win != wref).sum()} of {win.size}  worst grads", [(k, f"{errs[k]:.1e}") for k in worst])
