import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel
dev = torch.device("cuda")
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
torch.manual_seed(0)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).eval()
for name, lens, groups in (("full g1", LENS, 1), ("full g2", LENS, 2), ("videos 0-2 g1", LENS[:3], 1), ("videos 3-7 g1", LENS[3:], 1),
                           ("one video 4000", LENS[:1], 1), ("one video 700", LENS[-1:], 1)):
    x, _ = synth_batch(lens, DIM, NCLASS, 1); x = x.to(dev)
    net.stream_groups = groups
    def fwd():
        with torch.no_grad(): return net(x, lens)
    for _ in range(3): fwd()
    s = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fwd(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s): fwd()
    torch.cuda.synchronize()
    print(f"{name:18s} tiles {sum((l+127)//128 for l in lens):4d}  fwd {timeit(g.replay)*1e3:8.1f} us")
