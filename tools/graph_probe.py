"""Does CUDA-graph replay of the whole train step beat stream launches? (launch-overhead probe)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
dev = torch.device("cuda")
torch.manual_seed(0)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
net.stream_groups = int(os.environ.get("GROUPS", "2"))
crit = FrameCrossEntropy()
x, y = synth_batch(LENS, DIM, NCLASS, 1234); x, y = x.to(dev), y.to(dev)
def step():
    for p in net.parameters(): p.grad = None
    loss = crit(net(x, LENS), y); loss.backward(); return loss
for _ in range(5): step()
torch.cuda.synchronize()
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_cpu_issue = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu_issue / n * 1e3
print("stream launches: gpu %.3f ms/step, cpu issue %.3f ms/step" % timeit(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        loss = step()
torch.cuda.synchronize()
print("graph replay   : gpu %.3f ms/step, cpu issue %.3f ms/step" % timeit(g.replay))
