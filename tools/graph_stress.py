"""Determinism stress of the CAPTURED step with changing data: two input batches A / B replayed alternately through the
same workspace (a stale tile read would pick up the other batch's values); every replay's loss and gradients must equal
the first replay of the same batch bit for bit.  python tools/graph_stress.py [--videos 8] [--iters 3000]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep
ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=8)
ap.add_argument("--iters", type=int, default=3000)
ap.add_argument("--poison", action="store_true", help="fill the whole workspace (tile flags included) with NaN before every replay")
a = ap.parse_args()
dev = torch.device("cuda", 0)
lens = sorted(LENS * (a.videos // 8), reverse=True)
torch.manual_seed(0)
net = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
batches = [tuple(t.to(dev) for t in synth_batch(lens, 400, NCLASS, 1234 + i)) for i in range(2)]
net.set_dropout_state(77, 0)
g = GraphedTrainStep(net, crit, lens, batches[0][0], batches[0][1], n_valid=sum(lens), inputs=batches)
ref = [None, None]
bad = 0
for i in range(a.iters):
    k = i & 1
    net._drop_counter.fill_(3)
    if a.poison:
        net.last_workspace[0].fill_(float('nan'))
    l = g.replay(k)
    torch.cuda.synchronize()
    cur = (float(l), net.flat_parameters()[1].clone())
    if ref[k] is None:
        ref[k] = cur
    elif cur[0] != ref[k][0] or not torch.equal(cur[1], ref[k][1]):
        bad += 1
        d = (cur[1] - ref[k][1]).abs()
        names = [n for (n, p), o in zip(net.named_parameters(), net.grad_offsets()) if float(d[o:o + p.numel()].max()) > 0 or not bool(torch.isfinite(cur[1][o:o + p.numel()]).all())]
        nonfinite = int((~torch.isfinite(cur[1])).sum())
        print(f"   non-finite grads {nonfinite}; {len(names)} tensors differ: {names[:4]} ... {names[-3:]}")
        print(f"replay {i} (batch {k}): MISMATCH loss {cur[0]!r} vs {ref[k][0]!r}; max grad diff {float(d.max()):.3e} (rel {float(d.max() / ref[k][1].abs().max()):.2e})")
print(f"{a.iters} alternating graph replays, B={len(lens)}: {bad} mismatching")
