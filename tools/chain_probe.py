"""Forward-only and forward+backward time of the config-2 step (graph replay), for the chain launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
dev = torch.device("cuda")
x, y = synth_batch(LENS, DIM, NCLASS, 1234); x, y = x.to(dev), y.to(dev)
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
torch.manual_seed(0)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
def fwd():
    with torch.no_grad(): return net(x, LENS)
def step():
    for p in net.parameters(): p.grad = None
    loss = crit(net(x, LENS), y); loss.backward(); return loss
for _ in range(3): fwd(); step()
s = torch.cuda.Stream()
for name, fn in (("fwd(no grad)", fwd), ("fwd+bwd", step)):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s): fn()
    torch.cuda.synchronize()
    print(f"{name}: {timeit(g.replay):.3f} ms")
