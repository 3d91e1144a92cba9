"""Where does a non-deterministic replay first go wrong?  Two batches replayed alternately (as tools/graph_stress.py); the whole
workspace of the first replay of each batch is kept, and a replay whose loss / gradients differ is compared with it plane by
plane in forward order: the first differing plane names the kernel, the differing frames / channels its tile and warp.
python tools/locate_race.py [--videos 8] [--iters 6000] [--max-reports 6]"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, GraphedTrainStep, _cabi
ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=8)
ap.add_argument("--iters", type=int, default=6000)
ap.add_argument("--max-reports", type=int, default=6)
a = ap.parse_args()
dev = torch.device("cuda", 0)
lens = sorted(LENS * (a.videos // 8), reverse=True)
B, T = len(lens), max(lens)
torch.manual_seed(0)
net = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
crit = FrameCrossEntropy()
batches = [tuple(t.to(dev) for t in synth_batch(lens, 400, NCLASS, 1234 + i)) for i in range(2)]
net.set_dropout_state(77, 0)
g = GraphedTrainStep(net, crit, lens, batches[0][0], batches[0][1], n_valid=sum(lens), inputs=batches)
ws = net.last_workspace[0]
lib = _cabi.lib()


def planes():
    """(name, offset, columns) of every forward plane, in the order the forward writes them"""
    out = []
    for s in range(STAGES):
        for l in range(LAYERS + 1):
            out.append((f"stage {s} input of layer {l}" if l < LAYERS else f"stage {s} last activation", 0, s, l, 64))
            if l < LAYERS:
                out.append((f"stage {s} relu output h of layer {l}", 1, s, l, 64))
        out.append((f"stage {s} logits", 2, s, 0, NCLASS))
        out.append((f"stage {s} softmax*mask q", 3, s, 0, 64))
    res = []
    for name, what, s, l, cols in out:
        off = lib.mstcn_workspace_offset(C.byref(net._dims), B, T, 1, what, s, l)
        if off >= 0:
            res.append((name, off, cols))
    return sorted(res, key=lambda r: 0) if False else res


PL = planes()
ref = [None, None]
bad = 0
for i in range(a.iters):
    k = i & 1
    net._drop_counter.fill_(3)
    l = g.replay(k)
    torch.cuda.synchronize()
    cur = (float(l), net.flat_parameters()[1].clone())
    if ref[k] is None:
        ref[k] = (cur[0], cur[1], ws.clone())
        continue
    if cur[0] == ref[k][0] and torch.equal(cur[1], ref[k][1]):
        continue
    bad += 1
    if bad > a.max_reports:
        continue
    print(f"replay {i} (batch {k}): loss {cur[0]!r} vs {ref[k][0]!r}; forward planes that differ, in forward order:")
    shown = 0
    for name, off, cols in PL:
        x, r = ws[off: off + B * T * cols].view(B, T, cols), ref[k][2][off: off + B * T * cols].view(B, T, cols)
        ne = x != r
        if not bool(ne.any()):
            continue
        fr = ne.any(dim=2).nonzero()          # (video, frame) pairs
        ch = ne.any(dim=0).any(dim=0).nonzero().flatten().tolist()
        vids = sorted(set(fr[:, 0].tolist()))
        desc = []
        for v in vids[:3]:
            t = fr[fr[:, 0] == v][:, 1]
            desc.append(f"video {v} (len {lens[v]}): {t.numel()} frames in [{int(t.min())}, {int(t.max())}] = tiles {int(t.min()) // 128}..{int(t.max()) // 128}")
        maxd = float((x - r).abs().max())
        print(f"   {name}: {int(ne.sum())} elements, max |diff| {maxd:.3e}, channels {ch[:8]}{'...' if len(ch) > 8 else ''} ({len(ch)} of {cols}); " + "; ".join(desc))
        shown += 1
        if shown >= 4:
            break
    if shown == 0:
        print("   no forward plane differs: the backward went wrong")
print(f"{a.iters} alternating graph replays, B={B}: {bad} mismatching")
