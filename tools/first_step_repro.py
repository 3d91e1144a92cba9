"""Reproduction harness for a first-step discrepancy: (optional prelude of config-2 eager + graph-replay steps, as the test
file orders them) -> a NEW 64-video model -> ONE eager train step -> hashes of selected gradients.  Run it several times and
compare the lines."""
import argparse, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep
ap = argparse.ArgumentParser()
ap.add_argument("--prelude", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
crit = FrameCrossEntropy()
if a.prelude:
    torch.manual_seed(1)
    n2 = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
    x2, y2 = [t.to(dev) for t in synth_batch(LENS, 400, NCLASS, 7)]
    n2.set_dropout_state(5, 1)
    n2.zero_grad(); l = crit(n2(x2, LENS), y2); l.backward(); torch.cuda.synchronize()
    g = GraphedTrainStep(n2, crit, LENS, x2, y2, n_valid=sum(LENS), inputs=[(x2.clone(), y2.clone())])
    for _ in range(3):
        g.replay(0)
    torch.cuda.synchronize()
    del g, n2
    n3 = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
    g = GraphedTrainStep(n3, crit, LENS, x2, y2, n_valid=sum(LENS), inputs=[(x2, y2)])
    for _ in range(50):
        n3._drop_counter.fill_(3); g.replay(0); torch.cuda.synchronize()
    del g, n3
lens = sorted(LENS * 8, reverse=True)
torch.manual_seed(0)
net = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
gen = torch.Generator().manual_seed(3)
x = torch.randn(64, 4000, 400, generator=gen)
y = torch.randint(1, NCLASS, (64, 4000), generator=gen)
for b, l in enumerate(lens):
    x[b, l:] = 0; y[b, l:] = -1
for i in range(a.steps):
    net.set_dropout_state(2024, 9)
    net.zero_grad()
    out = net(x.cuda(), lens)
    loss = crit(out, y.flatten().cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in net.named_parameters()}
    h = lambda t: hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()[:10]
    allh = hashlib.sha1(net.flat_parameters()[1].cpu().numpy().tobytes()).hexdigest()[:10]
    print(f"step {i}: loss {float(loss.detach())!r} all {allh} stages.1.conv_1x1.w {h(grads['stages.1.conv_1x1.weight'])} "
          f"stages.0.conv_1x1.w {h(grads['stages.0.conv_1x1.weight'])} stages.1.conv_1x1.b {h(grads['stages.1.conv_1x1.bias'])} "
          f"|g| {float(grads['stages.1.conv_1x1.weight'].abs().sum()):.6f}", flush=True)
