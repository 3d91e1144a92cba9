"""Determinism stress of the eager train step (flag-linked kernels, side-stream weight gradients): the same batch N times,
every gradient compared bit for bit with the first run.  python tools/race_stress.py [--videos 64] [--iters 20] [--warm-graph]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep
ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warm-graph", action="store_true", help="first run graph-replayed config-2 steps of another model (test order)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
crit = FrameCrossEntropy()
if a.warm_graph:
    torch.manual_seed(1)
    n2 = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
    x2, y2 = [t.to(dev) for t in synth_batch(LENS, 400, NCLASS, 7)]
    g = GraphedTrainStep(n2, crit, LENS, x2, y2, n_valid=sum(LENS))
    for _ in range(20):
        g(x2, y2)
    torch.cuda.synchronize()
    del g, n2
lens = sorted(LENS * (a.videos // 8), reverse=True)
torch.manual_seed(0)
net = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
x, y = [t.to(dev) for t in synth_batch(lens, 400, NCLASS, 1234)]
ref = None
bad = 0
b = net.bucket_boundaries()
for i in range(a.iters):
    net.set_dropout_state(2024, 9)
    net.zero_grad()
    loss = crit(net(x, lens), y)
    loss.backward()
    torch.cuda.synchronize()
    gfl = net.flat_parameters()[1].clone()
    if ref is None:
        ref, ref_l = gfl, float(loss.detach())
    elif not torch.equal(gfl, ref) or float(loss.detach()) != ref_l:
        bad += 1
        diff = (gfl - ref).abs()
        where = [j for j in range(len(b) - 1) if float(diff[b[j]:b[j + 1]].max()) > 0]
        names = [k for (k, p), o in zip(net.named_parameters(), net.grad_offsets()) if float(diff[o:o + p.numel()].max()) > 0]
        top = [k for k in names if any(k.startswith(pre) for pre in ("stages.2.", "stages.1.", "stages.0.", "stage1."))]
        groups = {}
        for k in names:
            groups.setdefault(k.split(".layers.")[0].rsplit(".conv", 1)[0], []).append(k)
        print("   differing tensors per stage:", {g: len(v) for g, v in groups.items()}, "| stage input convs differing:", [k for k in names if ".layers." not in k])
        print(f"iter {i}: MISMATCH loss {float(loss.detach())!r} vs {ref_l!r}; max grad diff {float(diff.max()):.3e} (rel {float(diff.max() / ref.abs().max()):.2e}); buckets {where}; tensors {names[:6]}")
print(f"{a.iters} iterations, B={len(lens)}: {bad} mismatching")
