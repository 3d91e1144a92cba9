import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, FusedAdam
torch.manual_seed(5)
nets = [MultiStageModel(8, 2, 2, 64, 5).cuda().eval() for _ in range(3)]
for n in nets[1:]:
    n.load_state_dict(nets[0].state_dict())
crit = FrameCrossEntropy()
x = torch.randn(2, 50, 8, device="cuda"); y = torch.randint(0, 5, (100,), device="cuda")
def maxdiff(m, n):
    return max(float((pa - pb).abs().max()) for pa, pb in zip(m.parameters(), n.parameters()))
def train(net, opt, sched, n, tag):
    for _ in range(n):
        opt.zero_grad(); l = crit(net(x, [50, 50]), y); l.backward()
        g = float(net.flat_parameters()[1].abs().sum())
        opt.step(); sched.step()
        print(f"  {tag}: loss {float(l.detach()):.6f} |g| {g:.6f} lr-> {opt.param_groups[0]['lr']:.2e} step {opt.state_dict()['state'][0]['step']}")
a, b, c = nets
oa = FusedAdam(a, lr=1e-2); sa = torch.optim.lr_scheduler.StepLR(oa, step_size=2, gamma=0.5)
ob = torch.optim.Adam(b.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8); sb = torch.optim.lr_scheduler.StepLR(ob, step_size=2, gamma=0.5)
train(a, oa, sa, 3, "a"); train(b, ob, sb, 3, "b")
print("after 3: a-b", maxdiff(a, b))
sd = oa.state_dict()
c.load_state_dict(a.state_dict())
oc = FusedAdam(c, lr=1.0); oc.load_state_dict(sd)
sc = torch.optim.lr_scheduler.StepLR(oc, step_size=2, gamma=0.5, last_epoch=-1)
sc.last_epoch, sc._step_count = sa.last_epoch, sa._step_count
print("a-c params after load", maxdiff(a, c), "a-b", maxdiff(a, b))
ob.load_state_dict(copy.deepcopy(sd))
print("a-b after ob.load", maxdiff(a, b))
train(a, oa, sa, 2, "a"); print("a-b after a trained", maxdiff(a, b), "a-c", maxdiff(a, c))
train(c, oc, sc, 2, "c"); print("a-c", maxdiff(a, c))
train(b, ob, sb, 2, "b"); print("a-b", maxdiff(a, b))
