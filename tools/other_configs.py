"""Step time (fwd + CE + bwd, eager launches) of BASELINE configs 3 (B=64 on one GPU) and 4 (B=1, T=16384, D=2048)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
dev = torch.device("cuda")
def run(name, dim, lens, K=48):
    B, T = len(lens), max(lens)
    torch.manual_seed(0)
    net = MultiStageModel(dim, 4, 10, 64, K).to(dev).train()
    x = torch.randn(B, T, dim, device=dev)
    y = torch.randint(1, K, (B, T), device=dev)
    for b, n in enumerate(lens):
        x[b, n:] = 0; y[b, n:] = -1
    y = y.flatten()
    crit = FrameCrossEntropy()
    def step():
        net.zero_grad()
        loss = crit(net(x, lens), y, n_valid=sum(lens)); loss.backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: B={B} T={T} D={dim} valid frames {sum(lens)}: {ms:.3f} ms/step = {sum(lens) / ms / 1e3:.2f} M valid frames/s")
base = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700]
run("config 2 (eager)", 400, base)
run("config 3 (B=64, one batch)", 400, sorted(base * 8, reverse=True))
run("config 4", 2048, [16384])
