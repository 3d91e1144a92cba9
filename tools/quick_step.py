"""Graph-replay step time only (A/B of kernel variants: MSTCN_B200_LIB=variants/libmstcn_X.so python tools/quick_step.py [--config 2|3|4])."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dim, lens = 400, list(LENS)
if a.config == 3:
    lens = sorted(LENS * 8, reverse=True)
elif a.config == 4:
    dim, lens = 2048, [16384]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = MultiStageModel(dim, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
nrot = 4 if a.config != 3 else 2
res = [tuple(t.to(dev) for t in synth_batch(lens, dim, NCLASS, 1234 + i)) for i in range(nrot)]
g = GraphedTrainStep(net, crit, lens, res[0][0], res[0][1], n_valid=sum(lens), inputs=res)
for i in range(5):
    g.replay(i % nrot)
torch.cuda.synchronize()
best = []
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = g.replay(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / a.steps)
print(f"lib={os.environ.get('MSTCN_B200_LIB', 'default')} config={a.config} ms/step={min(best):.4f} (all {[round(b, 4) for b in best]}) "
      f"Mframes/s={sum(lens) / min(best) / 1e3:.2f} loss={float(loss):.5f}")
