"""Three training-mode forwards of the config-2 batch (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel
dev = torch.device("cuda")
x, y = synth_batch(LENS, DIM, NCLASS, 1234); x = x.to(dev)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
with torch.no_grad():
    for _ in range(3): net(x, LENS)
torch.cuda.synchronize()
print("ok")
