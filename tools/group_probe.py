"""Forward-only / backward-only timing under graph replay for stream_groups = 1, 2, 4."""
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
for groups in (1, 2, 4):
    torch.manual_seed(0)
    net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
    net.stream_groups = groups; net.backward_stream_groups = groups
    crit = FrameCrossEntropy()
    def fwd():
        with torch.no_grad(): return net(x, LENS)
    def step():
        for p in net.parameters(): p.grad = None
        loss = crit(net(x, LENS), y); loss.backward(); return loss
    for _ in range(3): fwd(); step()
    s = torch.cuda.Stream()
    res = {}
    for name, fn in (("fwd(no grad)", fwd), ("fwd+bwd", step)):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            fn(); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s): fn()
        torch.cuda.synchronize()
        res[name] = timeit(g.replay)
    print(f"groups {groups}: " + "  ".join(f"{k} {v:.3f} ms" for k, v in res.items()))
