"""Does any kernel read a workspace row that nobody wrote?  The train step is run on a workspace pre-filled with zeros, with
1e30 and with NaN (MSTCN_POISON_WS); loss and every gradient must be bit-identical.  python tools/poison_check.py [--videos 8]"""
import argparse, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--videos", type=int, default=8)
ap.add_argument("--child", default=None)
ap.add_argument("--eval", action="store_true")
a = ap.parse_args()
if a.child is None:
    outs = {}
    for val in ("0", "1e30", "nan"):
        env = dict(os.environ, MSTCN_POISON_WS=val)
        r = subprocess.run([sys.executable, __file__, "--videos", str(a.videos), "--child", val] + (["--eval"] if a.eval else []),
                           env=env, capture_output=True, text=True)
        outs[val] = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "ERR " + r.stderr[-300:]
        print(val, "->", outs[val])
    print("IDENTICAL" if len(set(outs.values())) == 1 else "DIFFERENT: some kernel reads rows nobody wrote")
    sys.exit(0)
import hashlib
import torch
from bench import synth_batch, LENS, STAGES, LAYERS, FMAPS, NCLASS
from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy
dev = torch.device("cuda", 0)
lens = sorted(LENS * (a.videos // 8), reverse=True)
torch.manual_seed(0)
net = MultiStageModel(400, STAGES, LAYERS, FMAPS, NCLASS).to(dev)
net.train(not a.eval)
x, y = [t.to(dev) for t in synth_batch(lens, 400, NCLASS, 1234)]
net.set_dropout_state(2024, 9)
net.zero_grad()
out = net(x, lens)
loss = FrameCrossEntropy()(out, y)
loss.backward()
torch.cuda.synchronize()
g = net.flat_parameters()[1]
bad = [k for (k, p), o in zip(net.named_parameters(), net.grad_offsets()) if not torch.isfinite(g[o:o + p.numel()]).all()]
h = hashlib.sha1(g.cpu().numpy().tobytes()).hexdigest()[:16]
ho = hashlib.sha1(out.detach().cpu().numpy().tobytes()).hexdigest()[:16]
print(f"loss {float(loss.detach())!r} out {ho} grads {h} non-finite tensors {bad[:8]}")
