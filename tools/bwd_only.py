"""Three training steps of the config-2 batch (profiling target for the backward kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
dev = torch.device("cuda")
x, y = synth_batch(LENS, DIM, NCLASS, 1234); x, y = x.to(dev), y.to(dev)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
for _ in range(3):
    net.zero_grad()
    loss = crit(net(x, LENS), y); loss.backward()
torch.cuda.synchronize()
print("ok")
