"""Per-role wait / work clock totals of CTA 0 of the last stage weight-gradient launch (config-2 step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, _cabi
lib = _cabi.lib()
dev = torch.device("cuda")
x, y = synth_batch(LENS, DIM, NCLASS, 1234); x, y = x.to(dev), y.to(dev)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
def step():
    net.zero_grad()
    loss = crit(net(x, LENS), y); loss.backward()
for _ in range(3): step()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib.mstcn_debug_tc_timing(_cabi.ptr(buf))
step()
torch.cuda.synchronize()
lib.mstcn_debug_tc_timing(None)
t = buf.cpu().tolist()
print(f"producer : waited bempty {t[32]:8d}  aempty {t[33]:8d} clk   A loads {t[34]}  B loads {t[35]}")
print(f"MMA warp : waited bready {t[36]:8d}  aready {t[37]:8d} clk   total {t[38]:8d}")
print(f"transform: waited bfull  {t[39]:8d}  afull  {t[40]:8d} clk   loop  {t[41]:8d}   wait done {t[42]:8d}  end {t[43]:8d}")
