"""torchrun ... tools/dp_quick.py [--allreduce peer|nccl] [--nvls] [--overlap on|off] : device-timed graph-replay step of the
weak-scaling workload (8 videos per rank), max over ranks -- the A/B harness for the gradient all-reduce variants."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep
from pytorch_video_action_b200.parallel import DataParallelMSTCN

ap = argparse.ArgumentParser()
ap.add_argument("--allreduce", default="peer")
ap.add_argument("--nvls", default="auto")
ap.add_argument("--overlap", default="on")
ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
torch.manual_seed(0)
net = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
dp = DataParallelMSTCN(net, crit, overlap=a.overlap == "on", allreduce=a.allreduce, nvls={"auto": None, "on": True, "off": False}[a.nvls])
res = [tuple(t.to(dev) for t in synth_batch(LENS, 400, NCLASS, 1234 + 100 * rank + i)) for i in range(4)]
g = GraphedTrainStep(net, crit, LENS, res[0][0], res[0][1], n_valid=sum(LENS) * world, dp=dp, inputs=res)
for i in range(5):
    g.replay(i % 4)
best = []
for r in range(3):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        g.replay(i % 4)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best.append(float(t))
if rank == 0:
    print(f"world={world} allreduce={dp.allreduce} overlap={a.overlap} skip={os.environ.get('MSTCN_DP_SKIP', '0')} "
          f"ms/step={min(best):.4f} {[round(b, 4) for b in best]} Mframes/s={sum(LENS) * world / min(best) / 1e3:.2f}", flush=True)
del g
import gc; gc.collect(); torch.cuda.synchronize()
dist.barrier()
os._exit(0)
