"""Per-task trace of the last forward chain launch (stage S-1): dependency latency and task duration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, _cabi
import os
if os.environ.get('TRACE_LENS'):
    LENS = [int(v) for v in os.environ['TRACE_LENS'].split(',')]
lib = _cabi.lib()
dev = torch.device("cuda")
x, y = synth_batch(LENS, DIM, NCLASS, 1234); x = x.to(dev)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
B, T = x.shape[:2]
tpv = (T + 127) // 128; nt = B * tpv
buf = torch.zeros(LAYERS * nt * 8, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3): net(x, LENS)
    lib.mstcn_debug_chain_trace(_cabi.ptr(buf))
    net(x, LENS)
    torch.cuda.synchronize()
    lib.mstcn_debug_chain_trace(None)
t = buf.cpu().numpy().reshape(LAYERS, nt, 8).astype(np.int64)
lens = np.array(LENS)
valid = np.array([[ti * 128 < lens[b] for ti in range(tpv)] for b in range(B)]).reshape(-1)
pub = t[:, :, 3]
t00 = pub[0][pub[0] > 0].min()
for l in range(LAYERS - 1):
    v = valid & (pub[l] > 0)
    line = f"layer {l}: published {int(pub[l][v].min() - t00):7d} .. {int(pub[l][v].max() - t00):7d} ns"
    if l > 0:
        poll, ready, g1 = t[l, :, 0], t[l, :, 1], t[l, :, 2]
        vv = valid & (ready > 0)
        d = 1 << l
        lat = []
        for b in range(B):
            for ti in range(tpv):
                i = b * tpv + ti
                if not vv[i]: continue
                deps = set()
                for k in (-1, 0, 1):
                    tf = ti * 128 + k * d
                    if tf + 127 < 0 or tf >= T: continue
                    deps.add(max(tf, 0) // 128); deps.add(min(tf + 127, T - 1) // 128)
                last = max(pub[l - 1][b * tpv + j] for j in deps)
                lat.append((ready[i] - last, ready[i] - poll[i], g1[i] - ready[i], pub[l][i] - g1[i],
                            t[l, i, 4] - ready[i], t[l, i, 5] - t[l, i, 4], t[l, i, 6] - t[l, i, 5], g1[i] - t[l, i, 5],
                            t[l, i, 7] - g1[i], pub[l][i] - t[l, i, 7]))
        lat = np.array(lat)
        line += ("  | dep->ready med %5d max %5d  | polled med %5d  | ready->g1 med %5d  | g1->publish med %5d max %5d"
                 % (np.median(lat[:, 0]), lat[:, 0].max(), np.median(lat[:, 1]), np.median(lat[:, 2]),
                    np.median(lat[:, 3]), lat[:, 3].max()))
        line += "\n      ready->tma %5d  tma->landed %5d  landed->lo parked %5d  landed->g1 %5d  g1->g2 %5d  g2->publish %5d" % tuple(
            np.median(lat[:, j]) for j in range(4, 10))
    print(line)

# critical path: walk back from the last-published tile of the last traced layer through its latest dependency
print("\ncritical path (times in ns relative to the first publish; cta = task % 148):")
l = LAYERS - 2
i = int(np.argmax(np.where(valid, pub[l], 0)))
while l >= 1:
    b, ti = divmod(i, tpv)
    d = 1 << l
    deps = set()
    for k in (-1, 0, 1):
        tf = ti * 128 + k * d
        if tf + 127 < 0 or tf >= T: continue
        deps.add(max(tf, 0) // 128); deps.add(min(tf + 127, T - 1) // 128)
    j = max(deps, key=lambda q: pub[l - 1][b * tpv + q])
    r = t[l, i]
    task = l * nt + i
    print(f"  layer {l} tile ({b},{ti:2d}) cta {task % 148:3d}: poll {r[0]-t00:7d} ready {r[1]-t00:7d} (+{r[1]-pub[l-1][b*tpv+j]:5d} after dep ({b},{j}) "
          f"{'valid' if valid[b*tpv+j] else 'PAD'}) tma {r[4]-r[1]:5d} landed {r[5]-r[4]:5d} lo {r[6]-r[5]:5d} g1 {r[2]-r[5]:5d} g2 {r[7]-r[2]:5d} pub {r[3]-r[7]:5d}  total {r[3]-r[1]:5d}")
    i = b * tpv + j
    l -= 1

print("\nper-video completion time of each layer (ns, max over the video's valid tiles) and mean cadence:")
for b in range(B):
    idx = [b * tpv + ti for ti in range(tpv) if valid[b * tpv + ti]]
    ends = [int(pub[l][idx].max() - t00) for l in range(LAYERS - 1)]
    tot = [int(np.median(t[l, idx, 3] - t[l, idx, 1])) for l in range(1, LAYERS - 1)]
    print(f"  video {b} ({len(idx):2d} tiles): " + " ".join(f"{e:6d}" for e in ends) + f"   cadence {(ends[-1]-ends[0])/(len(ends)-1):6.0f}  median ready->publish {int(np.median(tot))}")
