"""Standalone launches of the fused dilated-residual forward kernel on the BASELINE config-2 shape
(the same call bench.py times for its `roofline` object); run under ncu for the per-launch DRAM traffic."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pytorch_video_action_b200 import MultiStageModel
dev = torch.device("cuda")
net = MultiStageModel(bench.DIM, bench.STAGES, bench.LAYERS, bench.FMAPS, bench.NCLASS).to(dev).train()
x, _ = bench.synth_batch(bench.LENS, bench.DIM, bench.NCLASS, 1)
with torch.no_grad():
    net(x.to(dev), bench.LENS)
t = bench.time_layer_kernel(net, x.to(dev), bench.LENS, 3)
print(f"avg launch {t * 1e6:.2f} us -> {512.0 * sum(bench.LENS) / t / 1e9:.1f} GB/s algorithmic")
