import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, FusedAdam
torch.manual_seed(5)
nets = [MultiStageModel(8, 2, 2, 64, 5).cuda().eval() for _ in range(3)]
for n in nets[1:]:
    n.load_state_dict(nets[0].state_dict())
crit = FrameCrossEntropy()
x = torch.randn(2, 50, 8, device="cuda")
y = torch.randint(0, 5, (100,), device="cuda")
def train(net, opt, n):
    for _ in range(n):
        opt.zero_grad(); crit(net(x, [50, 50]), y).backward(); opt.step()
a, b, c = nets
oa = FusedAdam(a, lr=1e-2); ob = torch.optim.Adam(b.parameters(), lr=1e-2); oc = torch.optim.Adam(c.parameters(), lr=1e-2)
# gradient determinism first
crit(a(x, [50, 50]), y).backward(); crit(b(x, [50, 50]), y).backward()
print("grad equal a vs b:", all(torch.equal(p.grad, q.grad) for p, q in zip(a.parameters(), b.parameters())))
train(a, oa, 3); train(b, ob, 3); train(c, oc, 3)
worst = max(((pa - pb).abs().max().item(), k) for (k, pa), pb in zip(a.named_parameters(), b.parameters()))
print("fused vs torch Adam after 3 steps: worst abs diff", worst)
worst = max(((pc - pb).abs().max().item(), k) for (k, pc), pb in zip(c.named_parameters(), b.parameters()))
print("torch vs torch Adam after 3 steps: worst abs diff", worst)
