"""Per-CTA timeline of consecutive tiles of a forward chain launch (throughput regime): where one CTA spends its time
between tiles.  TRACE stamps: 0 poll start, 1 deps seen, 4 first TMA issued, 5 centre tap ready for the tensor core, 6 all
taps parked, 2 tap GEMM done, 7 1x1 GEMM done, 3 published.   python tools/chain_period.py [--videos 64] [--cta 5] [--layer 3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_batch, LENS, DIM, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, _cabi
ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=64)
ap.add_argument("--cta", type=int, default=5)
ap.add_argument("--layer", type=int, default=3)
a = ap.parse_args()
lens = sorted(LENS * (a.videos // 8), reverse=True)
lib = _cabi.lib()
dev = torch.device("cuda")
x, y = synth_batch(lens, DIM, NCLASS, 1234); x = x.to(dev)
net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
B, T = x.shape[:2]
tpv = (T + 127) // 128; nt = B * tpv
buf = torch.zeros(LAYERS * nt * 8, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3): net(x, lens)
    lib.mstcn_debug_chain_trace(_cabi.ptr(buf))
    net(x, lens)
    torch.cuda.synchronize()
    lib.mstcn_debug_chain_trace(None)
t = buf.cpu().numpy().reshape(LAYERS * nt, 8).astype(np.int64)
grid = min(148, LAYERS * nt)
tasks = [k for k in range(a.cta, LAYERS * nt, grid) if k // nt == a.layer and t[k, 2] > 0]
print(f"B={B} T={T} tiles/layer={nt} grid={grid}; CTA {a.cta}, layer {a.layer}: {len(tasks)} compute tiles")
t0 = t[tasks[0], 1]
prev_pub = None
print("tile   deps   tma  ctr_rdy parked   g1     g2    pub   | period(g1) | g1-parked  g2-g1  pub-g2")
pg1 = None
for k in tasks:
    r = t[k]
    print(f"{k % nt:5d} {r[1]-t0:6d} {r[4]-t0:6d} {r[5]-t0:6d} {r[6]-t0:6d} {r[2]-t0:6d} {r[7]-t0:6d} {r[3]-t0:6d} | "
          f"{(r[2]-pg1) if pg1 else 0:6d}     | {r[2]-r[6]:6d} {r[7]-r[2]:6d} {r[3]-r[7]:6d}")
    pg1 = r[2]
g1s = np.array([t[k, 2] for k in tasks])
print("median period between tap-GEMM completions:", int(np.median(np.diff(g1s))), "ns")
allg1 = t[:, 2][t[:, 2] > 0]
print("launch span (first deps seen -> last publish):", int(t[:, 3].max() - t[:, 1][t[:, 1] > 0].min()), "ns for", len(allg1), "compute tiles")
