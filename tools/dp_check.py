"""torchrun --nproc-per-node G tools/dp_check.py : G-rank summed gradients (video-sharded batch, bucketed NCCL
all-reduce overlapped with backward) vs the same global batch on one GPU (SURVEY.md 8e parity test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
from pytorch_video_action_b200.parallel import DataParallelMSTCN, shard_videos, local_pad_length

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
dim, K = 400, 48
# 16 DISTINCT lengths: every rank gets its own length list, and every rank but one pads to local max + 1 (fact 0.5)
lens = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700, 3950, 3700, 3333, 2800, 2222, 1500, 999, 129]
rng = np.random.default_rng(0)
feats = [rng.standard_normal((n, dim)).astype(np.float32) for n in lens]
labs = [rng.integers(1, K, n) for n in lens]
Tg, nvalid = max(lens), sum(lens)

def batch(idx, T):
    x = np.zeros((len(idx), T, dim), np.float32); y = np.full((len(idx), T), -1, np.int64)
    for j, i in enumerate(idx):
        x[j, :lens[i]] = feats[i]; y[j, :lens[i]] = labs[i]
    return torch.from_numpy(x).to(dev), torch.from_numpy(y.reshape(-1)).to(dev), [lens[i] for i in idx]

torch.manual_seed(0)
net = MultiStageModel(dim, 4, 10, 64, K).to(dev).eval()
crit = FrameCrossEntropy()
dp = DataParallelMSTCN(net, crit)
mine = shard_videos(lens, world)[rank]
Tl = local_pad_length([lens[i] for i in mine], Tg)
x, y, ll = batch(mine, Tl)
if world > 1:
    every = [None] * world
    dist.all_gather_object(every, (ll, Tl))
    assert len({tuple(e[0]) for e in every}) == world, "ranks must hold distinct length lists"
    assert sum(1 for e in every if e[1] == max(e[0]) + 1) >= world - 1, "the local pad rule (+1 frame) must be exercised"
net.zero_grad()
loss = dp.forward_backward(x, ll, y, nvalid)
g_dp = net.flat_parameters()[1].clone()
dist.all_reduce(loss)
if rank == 0:
    print(f"default gradient all-reduce: {dp.allreduce}" + (f" (peer setup failed: {dp.fallback_reason})" if dp.fallback_reason else ""))
# every rank holds identical bits after the sum
chk = g_dp.double().sum().reshape(1).clone()
allchk = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allchk, chk)
assert all(float(c) == float(allchk[0]) for c in allchk), "ranks disagree on the summed gradient"
# the other all-reduce paths give the same sum (different summation order: 1e-6)
for mode, nvls in (("nccl", False), ("auto", False)):
    dpm = DataParallelMSTCN(net, crit, allreduce=mode, nvls=nvls)
    net.zero_grad()
    dpm.forward_backward(x, ll, y, nvalid)
    gm = net.flat_parameters()[1]
    e = float((gm - g_dp).abs().max() / g_dp.abs().max())
    if rank == 0:
        print(f"all-reduce {dpm.allreduce} (requested {mode}, nvls={nvls}): grad rel diff vs default {e:.1e}")
    assert e < 1e-5, (mode, nvls, e)
    del dpm
net.zero_grad()
dp = DataParallelMSTCN(net, crit)
dp.forward_backward(x, ll, y, nvalid)
rep = float((net.flat_parameters()[1] - g_dp).abs().max() / g_dp.abs().max())
if rank == 0:
    print(f"default all-reduce repeated: rel diff {rep:.1e}" + (" (bit-identical)" if rep == 0 else ""))
assert rep < 1e-6, rep
# a second micro-step WITHOUT zero_grad accumulates: 2x the reduced gradient, not world x G1 + G2
dp.forward_backward(x, ll, y, nvalid)
acc_err = float((net.flat_parameters()[1] - 2 * g_dp).abs().max() / g_dp.abs().max())
assert acc_err < 1e-6, acc_err
# the same step as a captured CUDA graph (fused loss head, per-stage hook or the single late all-reduce): identical gradients
from pytorch_video_action_b200 import GraphedTrainStep
for overlap in (True, False):
    dpg = DataParallelMSTCN(net, crit, overlap=overlap)
    step = GraphedTrainStep(net, crit, ll, x, y, n_valid=nvalid, dp=dpg)
    lg = step(x, y).clone()
    torch.cuda.synchronize()
    dist.all_reduce(lg)
    gerr = float((net.flat_parameters()[1] - g_dp).abs().max() / g_dp.abs().max())
    if rank == 0:
        print(f"graphed DP step ({dpg.allreduce}, overlap={overlap}): loss {float(lg):.6f}  grad rel diff vs eager DP {gerr:.1e}")
    assert gerr < 1e-6 and abs(float(lg) - float(loss)) < 1e-5
    del step
    torch.cuda.synchronize()
if rank == 0:
    x, y, ll = batch(list(range(len(lens))), Tg)
    net.zero_grad()
    l1 = crit(net(x, ll), y); l1.backward()
    g1 = net.flat_parameters()[1]
    b = net.bucket_boundaries()
    errs = [float((g_dp[b[i]:b[i+1]] - g1[b[i]:b[i+1]]).abs().max() / g1[b[i]:b[i+1]].abs().max()) for i in range(len(b) - 1)]
    print(f"world {world}: loss dp {float(loss):.6f} single {float(l1):.6f}  per-bucket grad rel err {['%.1e' % e for e in errs]}")
    assert abs(float(loss) - float(l1)) < 1e-4 and max(errs) < 1e-3
dist.barrier()
dist.destroy_process_group()
