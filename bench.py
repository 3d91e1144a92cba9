"""MS-TCN train-step benchmark (BASELINE.json metric: train frames/sec, fwd+bwd; % HBM roofline per
layer kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|eager-gpu]
                    [--config 2|3|4|5] [--scaling weak|strong]

Default (N == 1) workload = BASELINE configs[1] ("config 2"): 4 stages x 10 layers x 64 ch, 48 classes, D=400, batch of
8 padded/masked videos (T_pad=4000, lens from segment.txt quantiles, 21 132 valid frames), train mode (dropout on),
synthetic N(0,1) features and piecewise-constant labels, default-init weights.
N > 1 (torchrun, one rank per GPU):
  --scaling weak   (default) every rank runs that batch with its own feature seeds = configs[2] with a global batch of
                   8N videos; gradients summed by one all-reduce per step;
  --scaling strong configs[2] as SURVEY.md 8d states it: a FIXED global batch of 64 videos (8 copies of the config-2
                   lengths, sorted by length like BucketBatchSampler) sharded 64/N per rank by parallel.shard_videos,
                   each rank padded by parallel.local_pad_length; N == 1 runs the 64 videos as one batch.
Other single-GPU workloads: --config 3 (the 64 videos as one batch), --config 4 (B=1, T=16384, D=2048),
--config 5 (inference ensemble: 2 checkpoints x 32 segment.txt-shaped videos, argmax + segment vote + mode).

A step = zero_grad -> forward -> CrossEntropy(ignore_index=-1) -> backward; the Adam step is
timed separately (`with_adam`).  `value` has inputs resident in HBM; `e2e` goes through the public API with pinned
HOST buffers (H2D of features+labels and D2H of the loss inside the timed region) -- two host feeds are measured, the
reference's padded batch (`e2e_padded_host`) and the valid frames only, padded on the device (`e2e_ragged_host`); `e2e` is
the faster of the two in this run and names it (`feed`); `e2e_resident_feed` draws each
batch from a DeviceFeatureStore (the dataset lives in HBM; only the video indices cross PCIe).
--impl reference times the reference's CPU path on the host cores: the UNMODIFIED reference class when build() could
vendor it into the git-ignored oracle/_ref/ (kind "reference"), else the torch-CPU port in oracle/torch_port.py.
--impl eager-gpu times that same reference class run eagerly on the B200 (cuDNN / ATen) -- the bar to beat on the same
box -- for configs 1 / 2 / 4 with TF32 on and off.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LENS = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700]
DIM, STAGES, LAYERS, FMAPS, NCLASS = 400, 4, 10, 64, 48
METRIC = "mstcn_train_frames_per_sec_fwd_bwd"
UNIT = "valid frames/s"
N_ROTATE = 4          # distinct resident input batches rotated through (4 x 51 MB > 126 MB L2)


def synth_batch(lens, dim, n_class, seed, T=None):
    import torch
    g = torch.Generator().manual_seed(seed)
    B, T = len(lens), (T or max(lens))
    x = torch.randn(B, T, dim, generator=g)
    y = torch.full((B, T), -1, dtype=torch.long)
    for b, l in enumerate(lens):
        x[b, l:] = 0                                          # pad_batch zero-fills (train.py:188)
        t = 0
        while t < l:                                          # piecewise-constant runs, classes 1..K-1
            run = int(torch.randint(30, 401, (1,), generator=g))
            y[b, t:min(l, t + run)] = int(torch.randint(1, n_class, (1,), generator=g))
            t += run
    return x, y.flatten()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and therefore the pinned host buffers it first-touches) to the NUMA node its GPU hangs off.
    Returns a short description for the JSON line.  A single-node host leaves nothing to bind."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bus is None:
            out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bus = out[-12:].lower() if out else None
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if len(nodes) <= 1 or not bus:
            return f"{len(nodes)} NUMA node(s): nothing to bind"
        node = -1
        for cand in (bus, "0000:" + bus[-7:]):
            pth = f"/sys/bus/pci/devices/{cand}/numa_node"
            if os.path.exists(pth):
                node = int(open(pth).read().strip())
                break
        if node < 0:
            return "GPU NUMA node unknown: not bound"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return f"bound to NUMA node {node} ({len(cpus)} cpus)"
    except Exception as e:        # noqa: BLE001 -- best effort, never fatal
        return f"not bound ({type(e).__name__})"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# The reference itself (baseline arms only; never on the product path)
# ------------------------------------------------------------------------------------------------
def load_reference_class():
    """The UNMODIFIED reference MultiStageModel from oracle/_ref/networks.py (vendored by __graft_entry__.build() in the
    build container; git-ignored, travels with the snapshot), or None when it is absent."""
    path = os.path.join(ROOT, "oracle", "_ref", "networks.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("mstcn_reference_networks", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MultiStageModel


def make_reference_stepper(dim, lens, device, train=True, seed=0, xseed=1234):
    """One `zero_grad -> forward -> CrossEntropyLoss(ignore_index=-1) -> backward` step exactly as train.py:305-328 drives
    the model, on `device`.  Returns (step_fn, kind, valid_frames)."""
    import torch
    cls = load_reference_class()
    x, y = synth_batch(lens, dim, NCLASS, xseed)
    x, y = x.to(device), y.to(device)
    lens = list(lens)
    if cls is not None:
        torch.manual_seed(seed)
        net = cls(dim, STAGES, LAYERS, FMAPS, NCLASS).to(device)          # train.py:252 (+ explicit 4x10 as BASELINE says)
        net.train(train)
        crit = torch.nn.CrossEntropyLoss(ignore_index=-1)                  # train.py:266-267

        def step():
            net.zero_grad()                                                # train.py:305 (optimizer.zero_grad)
            out = net(x, lens)                                             # :308
            loss = crit(out, y)                                            # :326
            loss.backward()                                                # :328
            return loss
        return step, "reference", sum(lens)
    from oracle import torch_port as TP
    P = {k: v.to(device).requires_grad_(True) for k, v in TP.make_params(dim, STAGES, LAYERS, NCLASS, seed=seed).items()}
    if torch.device(device).type != "cpu":
        raise RuntimeError("oracle/_ref/networks.py is missing (run __graft_entry__.build() where /root/reference exists); "
                           "the torch port is a CPU baseline only")

    def step():
        return TP.train_step(P, x, lens, y, STAGES, LAYERS, NCLASS, train=train)[1]
    return step, "port", sum(lens)


def cpu_reference_frames_per_s(steps, warmup, threads=None):
    """The reference's CPU path on a bounded sample of the workload: the 2000-frame video of the config-2 batch alone
    (B=1, T=2000, D=400 = BASELINE configs[0]), train mode."""
    import torch
    step, kind, frames = make_reference_stepper(DIM, [2000], "cpu")
    tried = None
    if threads is None:
        # "all the host threads it can use": more threads are not always faster on a 64-channel model (and an
        # over-committed host can be far slower with all of them), so the thread count is calibrated -- one warm-up and
        # two timed steps per candidate -- and the timed run uses the fastest.  The candidates are reported.
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        tried = {}
        for n in sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), 1}, reverse=True):
            torch.set_num_threads(n)
            step()
            best = float("inf")
            for _ in range(2):
                t0 = time.perf_counter()
                step()
                best = min(best, time.perf_counter() - t0)
            tried[n] = frames / best
        threads = max(tried, key=tried.get)
    torch.set_num_threads(threads)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"fps": frames / med, "best": frames / min(times), "median_s": med, "threads": threads, "kind": kind,
            "threads_tried": tried}


CPU_SAMPLE = "B=1 T=2000 D=400 video of the config-2 batch (= BASELINE configs[0]), fwd+CE+bwd, dropout on"


def cpu_baseline_entry(steps, warmup, with_single_thread=True):
    r = cpu_reference_frames_per_s(steps, warmup)
    what = "the unmodified reference class (oracle/_ref/networks.py)" if r["kind"] == "reference" else "torch-CPU port of the reference"
    e = {"value": r["fps"], "unit": UNIT, "cores": r["threads"], "kind": r["kind"], "best": r["best"],
         "cpu_model": cpu_model(), "sample": f"{CPU_SAMPLE}, {what}"}
    if r.get("threads_tried"):
        e["threads_tried_frames_per_s"] = {str(k): v for k, v in r["threads_tried"].items()}
        e["host_cpus"] = max(r["threads_tried"])
    if with_single_thread:
        r1 = cpu_reference_frames_per_s(max(2, steps // 3), 1, threads=1)
        e["single_thread"] = {"value": r1["fps"], "cores": 1, "best": r1["best"]}
    return e, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warm = max(args.warmup, 1)
    entry, r = cpu_baseline_entry(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": r["median_s"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MS-TCN 4x10x64, K=48, D=400, B=8 padded videos T_pad=4000 (configs[1]); "
                               "reference arm steps a bounded B=1,T=2000 sample of it on the host CPU"},
        "cpu_baseline": entry,
        "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# The bar to beat on the same box: the reference run eagerly on the B200
# ------------------------------------------------------------------------------------------------
EAGER_CONFIGS = {
    1: ("configs[0] shape on the GPU: B=1, T=2000, D=400", 400, [2000]),
    2: ("configs[1]: B=8, T_pad=4000, D=400", 400, LENS),
    4: ("configs[3]: B=1, T=16384, D=2048", 2048, [16384]),
}


def eager_gpu_time(cfg, allow_tf32, cudnn_benchmark, steps, warmup, device="cuda:0"):
    """ms per fwd+CE+bwd step of the unmodified reference class on the GPU (CUDA events, after warm-up)."""
    import torch
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.benchmark = cudnn_benchmark
    _, dim, lens = EAGER_CONFIGS[cfg]
    step, kind, frames = make_reference_stepper(dim, lens, device)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": ms, "value": frames / ms * 1e3, "unit": UNIT, "allow_tf32": allow_tf32,
            "cudnn_benchmark": cudnn_benchmark, "loss": float(loss.detach())}


def gpu_eager_baseline(configs, steps, warmup, variants=((True, False), (False, False), (True, True))):
    """{config: [one entry per (allow_tf32, cudnn.benchmark) variant]} or {"unavailable": why}."""
    if load_reference_class() is None:
        return {"unavailable": "oracle/_ref/networks.py missing: build() was not run where /root/reference exists"}
    out = {"kind": "reference", "what": "the unmodified reference MultiStageModel run eagerly on this GPU (cuDNN/ATen), "
                                        "zero_grad -> forward -> CE -> backward, train mode, CUDA events"}
    for c in configs:
        out[f"config{c}"] = {"workload": EAGER_CONFIGS[c][0],
                             "runs": [eager_gpu_time(c, tf, bm, steps, warmup) for tf, bm in variants]}
    return out


def run_eager_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.cuda.set_device(0)
    steps, warm = max(3, min(args.steps, 20)), max(args.warmup, 3)
    res = gpu_eager_baseline([1, 2, 4], steps, warm)
    best = None
    if "config2" in res:
        best = max(res["config2"]["runs"], key=lambda r: r["value"])
    line = {"impl": "eager-gpu", "metric": METRIC, "value": best["value"] if best else None, "unit": UNIT, "n_gpus": 1,
            "steps": steps, "warmup": warm, "ms_per_step": best["ms_per_step"] if best else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 / tf32 (cuDNN)", "data": "synthetic",
            "config": {"workload": "the reference's own eager GPU path; headline = config 2 best variant"},
            "gpu_eager_baseline": res}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# per-kernel timing for the roofline entries
# ------------------------------------------------------------------------------------------------
def time_layer_kernel(net, x, lens, steps, n_frames=None):
    """Average launch duration of ONE fused dilated-residual forward launch (single layer).  Default: the 40
    (stage, layer) launches of the timed configuration; n_frames: one synthetic video batch of that many frames
    (B = n_frames / 16384 videos of 16384 frames) -- the large-N point SURVEY.md 8d asks for."""
    import ctypes as C
    import torch
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    if n_frames:
        T = 16384
        B = max(1, n_frames // T)
        lens = [T] * B
    B, T = len(lens), max(lens)
    N = B * T
    a = torch.randn(N, 64, device=x.device)
    yb = torch.empty_like(a)
    hb = torch.empty_like(a)
    lens_dev = torch.tensor(lens, dtype=torch.int32, device=x.device)
    S, L = net._dims.num_stages, net._dims.num_layers
    packed = net._packed
    dims = C.byref(net._dims)
    tcores = net.tensor_cores
    lay_off = [([int(lib.mstcn_packed_offset(dims, s, l, w)) for w in (3, 4, 5, 6, 12)], 1 << l, s * L + l)
               for s in range(S) for l in range(L)]
    drop = _cabi.MstcnDropout(1, 0, 7, 0)
    st = _cabi.stream_ptr()
    fsz = 4

    def launch(off, d, lid):
        pp = packed.data_ptr()
        if tcores:
            _cabi.check(lib.mstcn_layer_fwd_tc(_cabi.ptr(a), _cabi.ptr(yb), _cabi.ptr(hb), _cabi.ptr(lens_dev), B, T, d,
                                               C.c_void_p(pp + off[4] * fsz), C.c_void_p(pp + off[1] * fsz),
                                               C.c_void_p(pp + off[3] * fsz), C.byref(drop), lid, st))
            return
        _cabi.check(lib.mstcn_layer_fwd(_cabi.ptr(a), _cabi.ptr(yb), _cabi.ptr(hb), _cabi.ptr(lens_dev), B, T, d,
                                        C.c_void_p(pp + off[0] * fsz), C.c_void_p(pp + off[1] * fsz),
                                        C.c_void_p(pp + off[2] * fsz), C.c_void_p(pp + off[3] * fsz),
                                        C.byref(drop), lid, st))

    for off, d, lid in lay_off:
        launch(off, d, lid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, min(steps, 5))
    e0.record()
    for _ in range(reps):
        for off, d, lid in lay_off:
            launch(off, d, lid)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * len(lay_off)), N


def time_stage_chain(net, x, lens, steps):
    """Average duration of the stage chain launch (all num_layers fused dilated-residual layers of one stage in one
    persistent tcgen05 kernel -- the dominant kernel of the step), CUDA events on the launch stream, over the
    model's stages; planes are (L+1) x 8 MB + L x 8 MB per launch."""
    import ctypes as C
    import torch
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    B, T = len(lens), max(lens)
    N = B * T
    S, L = net._dims.num_stages, net._dims.num_layers
    planes = torch.randn((L + 1) * N, 64, device=x.device)
    hplanes = torch.empty(L * N, 64, device=x.device)
    flags = torch.zeros(L * B * ((T + 127) // 128), dtype=torch.int32, device=x.device)
    lens_dev = net._lens_device(lens, x.device)
    drop = _cabi.MstcnDropout(1, 0, 7, 0)
    st = _cabi.stream_ptr()

    def launch(s):
        _cabi.check(lib.mstcn_stage_fwd_tc(C.byref(net._dims), _cabi.ptr(net._packed), s, _cabi.ptr(planes), _cabi.ptr(hplanes),
                                           _cabi.ptr(lens_dev), B, T, C.byref(drop), _cabi.ptr(flags), st))

    for s in range(S):
        launch(s)
    torch.cuda.synchronize()
    reps = max(1, min(steps, 5))
    total = 0.0
    for _ in range(reps):
        for s in range(S):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            planes[:N].normal_()                      # fresh stage input; also pushes the previous result out of the way
            torch.cuda.synchronize()
            e0.record()
            launch(s)
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1) * 1e-3
    return total / (reps * S)


def time_backward_kernels(net, crit, x, y, lens, reps):
    """Average duration of the backward chain launch (tc_layer_kernel<2>, L-1 fused steps of a stage) and of the stage
    weight-gradient launch (tc_wgrad_kernel), each timed ALONE with CUDA events on its own launch stream inside a real
    eager train step (mstcn_debug_backward_timing drains the streams around them)."""
    import ctypes as C
    import torch
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    S = net._dims.num_stages
    chain, wgrad = [], []
    buf = (C.c_float * (2 * S))()
    for p in net.parameters():
        p.grad = None
    _cabi.check(lib.mstcn_debug_backward_timing(1))
    try:
        for _ in range(max(1, reps)):
            loss = crit(net(x, lens), y)
            loss.backward()
            torch.cuda.synchronize()
            _cabi.check(lib.mstcn_debug_backward_times(buf, 2 * S))
            chain += [buf[2 * s] * 1e-3 for s in range(S)]
            wgrad += [buf[2 * s + 1] * 1e-3 for s in range(S)]
            for p in net.parameters():
                p.grad = None
    finally:
        _cabi.check(lib.mstcn_debug_backward_timing(0))
    return sum(chain) / len(chain), sum(wgrad) / len(wgrad)


def tf32_peak(dev):
    """Measured dense TF32 tensor-core throughput of this GPU: cuBLAS fp32 matmul with TF32 allowed, 8192^3, best of 10
    (burst) and back to back for ~1 s (sustained) -- the TF32 counterpart of MEASURED_PEAKS.json's bf16 numbers."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    flop = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 400
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    sustained = flop * reps / (e0.elapsed_time(e1) * 1e-3)
    torch.backends.cuda.matmul.allow_tf32 = old
    return {"burst_tflops": best / 1e12, "sustained_tflops": sustained / 1e12,
            "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 10 / 400 back to back, CUDA events"}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def plan_workload(args, world, rank):
    """-> dict(lens (this rank), T (this rank's padded length), dim, valid_global, padded_global, videos_global, name)."""
    from pytorch_video_action_b200.parallel import shard_videos, local_pad_length
    cfg = args.config
    if cfg == 4:
        if world > 1:
            raise SystemExit("--config 4 is a single-GPU workload")
        return dict(lens=[16384], T=16384, dim=2048, valid_global=16384, padded_global=16384, videos_global=1,
                    name="B=1, T=16384, D=2048 (BASELINE configs[3], dilation to 512)")
    if cfg == 3 or args.scaling == "strong":
        glens = sorted(LENS * 8, reverse=True)                     # BucketBatchSampler sorts by length (data_utils.py:24)
        mine = [glens[i] for i in shard_videos(glens, world)[rank]]
        T = local_pad_length(mine, max(glens))
        return dict(lens=mine, T=T, dim=DIM, valid_global=sum(glens), padded_global=None, videos_global=64,
                    name=f"global batch of 64 padded/masked videos (8 x the config-2 lengths, length-sorted), D={DIM}, "
                         f"sharded {64 // world} videos per rank (BASELINE configs[2], SURVEY 8d config 3"
                         + (": the 64 videos as ONE batch" if world == 1 else "") + ")")
    return dict(lens=list(LENS), T=max(LENS), dim=DIM, valid_global=sum(LENS) * world, padded_global=8 * max(LENS) * world,
                videos_global=8 * world,
                name=f"per-GPU batch 8 padded/masked videos T_pad={max(LENS)} lens={LENS} (BASELINE configs[1]; x{world} ranks = configs[2])")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pytorch_video_action_b200 import (MultiStageModel, FrameCrossEntropy, FusedAdam, GraphedTrainStep,
                                           DeviceFeatureStore, RaggedBatchUploader)
    from pytorch_video_action_b200.parallel import DataParallelMSTCN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dbg = os.environ.get("MSTCN_BENCH_DEBUG")
    if dbg:                                   # hang diagnosis: dump every thread's Python stack after `dbg` seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(dbg), exit=True)

    def note(msg):
        if dbg:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    numa = bind_to_gpu_numa_node(local_rank)          # before any pinned allocation: first touch lands on the GPU's node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    wl = plan_workload(args, world, rank)
    lens, T, dim = wl["lens"], wl["T"], wl["dim"]
    B = len(lens)
    valid_global = wl["valid_global"]
    valid_local = sum(lens)

    torch.manual_seed(0)
    net = MultiStageModel(dim, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
    net.tensor_cores = not args.fp32_ffma
    crit = FrameCrossEntropy()
    opt = FusedAdam(net, lr=1e-3)
    dp = None
    if world > 1:
        dp = DataParallelMSTCN(net, crit, overlap=None if args.dp_overlap == "auto" else args.dp_overlap == "on",
                               allreduce=args.allreduce, nvls={"auto": None, "on": True, "off": False}[args.nvls])

    host = [synth_batch(lens, dim, NCLASS, 1234 + 100 * rank + i, T=T) for i in range(N_ROTATE)]
    host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]

    note("inputs resident")
    # device-side halves of the end-to-end H2D double buffer (below)
    dbuf = [(torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])) for _ in range(2)]
    graphed = None
    if not args.no_graph:
        # the whole step (fwd + CE + bwd, incl. the gradient all-reduce when world > 1) replayed as one CUDA graph;
        # one capture per input buffer (the resident batches and the two double-buffer halves), so a replay reads its
        # inputs where they already are
        graphed = GraphedTrainStep(net, crit, lens, resident[0][0], resident[0][1], n_valid=valid_global, dp=dp,
                                   inputs=resident + dbuf)
    slot_of = {id(t[0]): i for i, t in enumerate(resident + dbuf)}

    note("graph captured" if graphed is not None else "eager mode")

    graphed_adam = [None]

    def step(x, y, with_adam=False):
        if graphed is not None and with_adam:
            # the optimizer step inside the captured graph (device-side step count / learning rate): one replay =
            # zero_grad -> forward -> CE -> backward -> (gradient sum) -> Adam, as train.py:305-329 iterates
            if graphed_adam[0] is None:
                graphed_adam[0] = GraphedTrainStep(net, crit, lens, resident[0][0], resident[0][1], n_valid=valid_global, dp=dp,
                                                   inputs=resident, optimizer=opt)
            return graphed_adam[0].replay(slot_of[id(x)])
        if graphed is not None:
            loss = graphed.replay(slot_of[id(x)])
        else:
            opt.zero_grad()
            if dp is not None:
                loss = dp.forward_backward(x, lens, y, valid_global)
            else:
                loss = crit(net._forward_impl(x, lens, strict_len=False), y, n_valid=valid_global)
                loss.backward()
        if with_adam:
            opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(seconds):
        t = torch.tensor([seconds], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    W, K = max(args.warmup, 3), args.steps
    note("timing starts")
    for i in range(W):
        step(*resident[i % N_ROTATE])
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_dev = timed(lambda i: step(*resident[i % N_ROTATE]), K)
    clocks = sampler.stop() if sampler else None

    # end to end through the public API: every step's features + labels travel from pinned HOST buffers to the
    # device and the loss comes back to the host, all inside the timed region.  The H2D copy of step i+1 runs on
    # a copy stream under step i's compute (double-buffered device inputs), as a training loop would do it.
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch_host(i):
        hx, hy = host[i % N_ROTATE]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the step that last read this buffer pair is done
            dbuf[i % 2][0].copy_(hx, non_blocking=True)
            dbuf[i % 2][1].copy_(hy, non_blocking=True)
            ready[i % 2].record(copy_stream)

    # resident feed: the dataset lives in HBM (DeviceFeatureStore); a step ships B video indices and one
    # gather kernel builds the padded batch in the graph's input buffers -- what a real epoch loop over a device-resident
    # I3D feature set does
    store = DeviceFeatureStore([hx[b, :lens[b]] for hx, _ in host for b in range(B)],
                               [hy.view(B, T)[b, :lens[b]] for _, hy in host for b in range(B)], device=dev)
    lens_scratch = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2)]

    def prefetch_store(i):
        idx = [(i % N_ROTATE) * B + b for b in range(B)]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            store.pad_batch(idx, pad_to=T, out=(dbuf[i % 2][0], dbuf[i % 2][1], lens_scratch[i % 2]))
            ready[i % 2].record(copy_stream)

    # host feed without the padding: the collate's valid frames travel as one ragged pinned block per step and
    # pad_batch_kernel builds the reference's padded batch on the device (RaggedBatchUploader)
    ragged = [(torch.cat([hx[b, :lens[b]] for b in range(B)]).pin_memory(),
               torch.cat([hy.view(B, T)[b, :lens[b]] for b in range(B)]).pin_memory()) for hx, hy in host]
    uploaders = [RaggedBatchUploader(lens, dim, pad_to=T, device=dev) for _ in range(2)]

    def prefetch_ragged(i):
        rx, ry = ragged[i % N_ROTATE]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            uploaders[i % 2].upload(rx, ry, out=dbuf[i % 2])
            ready[i % 2].record(copy_stream)

    def e2e_run(n, prefetch):
        # every step's loss is copied to pinned host memory (D2H) and read by the host; the host reads step i-1's value
        # while step i runs, the way a training loop logs, so the read does not drain the GPU between steps
        losses, evs = [], []
        host_loss = torch.empty(n, dtype=torch.float32).pin_memory()
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i % 2])
            loss = step(*dbuf[i % 2])
            consumed[i % 2].record()
            host_loss[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)     # D2H read of the step's result
            ev = torch.cuda.Event()
            ev.record()
            evs.append(ev)
            if i >= 1:
                evs[i - 1].synchronize()
                losses.append(float(host_loss[i - 1]))
        evs[-1].synchronize()
        losses.append(float(host_loss[n - 1]))
        return losses

    def e2e_timed(prefetch):
        for e in consumed:
            e.record()
        e2e_run(3, prefetch)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(K, prefetch)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) * 1e-3)

    t_e2e_padded = e2e_timed(prefetch_host)
    t_e2e = e2e_timed(prefetch_ragged)
    note("e2e timed")
    t_e2e_store = e2e_timed(prefetch_store)
    # measured host->device copy rate of this rank's batch alone (names the host-feed limiter at N > 1)
    hx0, hy0 = host[0]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        dbuf[i % 2][0].copy_(host[i % N_ROTATE][0], non_blocking=True)
    e1.record()
    barrier()
    h2d_gbs = hx0.numel() * 4 * 8 / max_over_ranks(e0.elapsed_time(e1) * 1e-3) / 1e9
    step(*resident[0], with_adam=True)              # (captures the graph with the optimizer step outside the timed region)
    t_adam = timed(lambda i: step(*resident[i % N_ROTATE], with_adam=True), K)
    last_loss = float(step(*resident[0]).item())

    def shutdown():
        """Tear the process group down.  Captured graphs that hold NCCL kernels must be released first, and a
        communicator teardown that still stalls (seen with graph-captured collectives) must not hang the job."""
        nonlocal graphed
        if world == 1:
            return
        graphed = None
        graphed_adam[0] = None
        import gc
        import threading
        gc.collect()
        torch.cuda.synchronize()
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)
        if th.is_alive():
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        shutdown()
        return

    hbm, peak_kind = load_peaks()
    rooflines = {}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    ncu = {}
    if os.path.exists(tpath) and not args.fp32_ffma and args.config == 2 and args.scaling == "weak":
        with open(tpath) as f:
            ncu = json.load(f)
        traffic = ncu["dram_bytes_read_per_launch"] + ncu["dram_bytes_write_per_launch"]
    graphed_keep = graphed
    tf32 = None
    if args.fp32_ffma:
        t_layer, _ = time_layer_kernel(net, resident[0][0], lens, K)
        algo_bytes = 512.0 * valid_local            # 512 B per frame-layer, valid frames only
        kernel_name = "layer_fwd_kernel (fused dilated residual layer, fp32 FFMA)"
    else:
        t_layer = time_stage_chain(net, resident[0][0], lens, K)
        algo_bytes = 512.0 * valid_local * LAYERS   # one chain launch = all layers of a stage
        kernel_name = ("tc_layer_kernel<0> chain launch (the 10 fused dilated residual layers of a stage in one "
                       "persistent kernel, tcgen05 3xTF32 + TMA + TMEM)")
        tf32 = tf32_peak(dev)
        # the other two kernels that, with the forward chain, make ~70 % of the step; each timed alone inside a real step
        t_bchain, t_wgrad = time_backward_kernels(net, crit, resident[0][0], resident[0][1], lens, 2)
        # large-N point ("additionally time each layer kernel at N = 256 k and 1 M frames"): one single-layer
        # launch over 64 videos of 16384 frames -- bandwidth / tensor rate without the per-tile dependency latency
        t_big, n_big = time_layer_kernel(net, resident[0][0], lens, 2, n_frames=1 << 20)
        def ncu_of(key):
            e = ncu.get(key)
            return (e["gpu_time_us"] * 1e-6, e["dram_bytes_read_per_launch"] + e["dram_bytes_write_per_launch"]) if e else (None, None)

        def entry(algo_bytes, t_events, key, flop_tf32=None, note=None):
            """achieved / frac from the kernel's duration in the committed ncu capture of this workload when there is one
            (stored constant, config 2 only), beside the duration measured in this run with CUDA events around the launch
            inside a drained step (which adds the launch's cold start and is an upper bound)."""
            t_ncu, tr = ncu_of(key) if key else (None, None)
            t = t_ncu or t_events
            e = {"bound": "hbm", "avg_launch_us": t * 1e6, "duration_source": "ncu capture in profiles/ (stored constant)" if t_ncu else "CUDA events, this run",
                 "event_timed_alone_us": t_events * 1e6, "algorithmic_bytes_per_launch": algo_bytes,
                 "achieved": algo_bytes / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": algo_bytes / t / 1e9 / hbm, "traffic": tr,
                 "traffic_source": "ncu --set full, stored constant" if tr else None}
            if flop_tf32:
                e["tensor_frac"] = flop_tf32 / t / 1e12 / tf32["sustained_tflops"]
            if note:
                e["note"] = note
            return e
        rooflines = {
            "tc_layer_kernel<2> backward chain launch (L-1 fused gx(l)+gu(l-1) steps of a stage)":
                entry(768.0 * valid_local * (LAYERS - 1), t_bchain, "tc_layer_kernel<2> chain",
                      flop_tf32=3 * 32768.0 * valid_local * (LAYERS - 1)),
            "tc_wgrad_kernel (all weight gradients of a stage)":
                entry(1280.0 * valid_local * LAYERS, t_wgrad, "tc_wgrad_kernel", flop_tf32=4 * 4 * 2.0 * 64 * 64 * valid_local * LAYERS),
            f"tc_layer_kernel<0> single-layer launch at N = {n_big} frames (no cross-layer dependency)":
                entry(768.0 * n_big, t_big, None, flop_tf32=3 * 32768.0 * n_big,
                      note="training-mode launch: reads x, writes y AND h = 768 B per frame actually moved (512 B by the 8d accounting)"),
        }
    del graphed_keep
    achieved = algo_bytes / t_layer / 1e9
    cpu_entry = cpu_baseline_entry(10, 2)[0] if world == 1 else None      # reported on rank 0 at N = 1 only (the contract)
    eager = None
    if args.config == 2 and args.scaling == "weak" and not args.no_eager_baseline:
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline([2], 5, 3, variants=((True, False), (False, False)))

    if args.fp32_ffma:
        launches_per_step = 1 + 1 + STAGES * LAYERS + STAGES + 2 + 2 * STAGES + 4 * STAGES * LAYERS + 2
    else:
        # operand packing x2; forward: projection, per stage (chain + tail); fused loss head + its finalize; backward:
        # per stage (tail with the top layer's gu fused in, chain, layer-0 gx, weight gradients, two reductions), projection
        # gradient + reduction.  (Host-launch mode adds stage max, the separate CE kernels, gradient routing and torch's glue.)
        per_stage_bwd = 7 if os.environ.get("MSTCN_FUSE_GU") == "0" else 6
        launches_per_step = 2 + 1 + 2 * STAGES + 2 + per_stage_bwd * STAGES + 2
    hx, hy = host[0]
    strong = args.scaling == "strong" or args.config == 3
    roof = {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
            "peak_kind": peak_kind, "traffic": traffic,
            "traffic_source": "stored constant: dram__bytes_read.sum + dram__bytes_write.sum of this launch from the ncu --set full "
                              "capture in profiles/r02_final_tc_chain_fwd_full_raw.csv (not re-measured in this run)" if traffic is not None else None,
            "avg_launch_us": t_layer * 1e6, "algorithmic_bytes_per_launch": algo_bytes}
    if tf32 is not None:
        # tensor roofline beside the HBM one: the chain executes 3 TF32 products per algorithmic one (3xTF32)
        roof["tensor_frac"] = 3 * 32768.0 * valid_local * LAYERS / t_layer / 1e12 / tf32["sustained_tflops"]
        roof["tensor_peak_tf32"] = tf32
        roof["other_kernels"] = rooflines
    line = {
        "metric": METRIC, "value": valid_global * K / t_dev, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "strong" if (strong and world > 1) or args.scaling == "strong" else "weak",
        "vs_baseline": None,
        "dtype": "f32 (fp32 FFMA)" if args.fp32_ffma else "f32-equivalent (3xTF32 on tcgen05 for every forward / input-gradient GEMM, exact 4-term tf32 for every weight gradient)",
        "data": "synthetic",
        "config": {"workload": f"MS-TCN {STAGES}x{LAYERS}x{FMAPS}, K={NCLASS}, {wl['name']}, train mode (dropout on), fwd+CE+bwd",
                   "global_batch_videos": wl["videos_global"], "valid_frames_per_step": valid_global,
                   "padded_frames_per_step": wl["padded_global"],
                   "parallelism": f"dp{world}" + ("" if world == 1 else
                                                  f" (gradient sum: {dp.allreduce}, "
                                                  + ("per-stage buckets under the backward)" if getattr(dp.reducer, "overlap", False) else "one all-reduce after the backward)")),
                   "launch": "host launches" if args.no_graph else "CUDA-graph replay of the step (one capture per resident input buffer: replays read the inputs in place)",
                   "l2": f"{N_ROTATE} resident input batches rotated ({N_ROTATE * hx.numel() * 4 / 1e6:.0f} MB > 126 MB L2); "
                         "the saved activations (0.8 GB at config 2) stream through per step, no explicit flush",
                   "host_numa": numa},
        "with_adam": {"value": valid_global * K / t_adam, "unit": UNIT, "ms_per_step": t_adam / K * 1e3,
                      "what": "the same step with the fused Adam update inside the captured graph (train.py:329)"},
        "e2e": None,                # filled below: the faster of the two HOST-buffer feeds measured in this run
        "e2e_ragged_host": {"value": valid_global * K / t_e2e, "unit": UNIT, "ms_per_step": t_e2e / K * 1e3,
                            "h2d_bytes_per_step": uploaders[0].h2d_bytes, "d2h_bytes_per_step": 4,
                            "what": "pinned HOST buffers -> device inside the timed region, every step: the batch's valid frames as "
                                    "one ragged block (RaggedBatchUploader: the reference's pad_batch with the H2D copy first and "
                                    "the padding on the device), on a copy stream under the previous step; the loss is read back "
                                    "every step.  Ships 2/3 of the bytes: the better feed when the host memory system is the "
                                    "limit (8 ranks on one NUMA node)"},
        "e2e_padded_host": {"value": valid_global * K / t_e2e_padded, "unit": UNIT, "ms_per_step": t_e2e_padded / K * 1e3,
                            "h2d_bytes_per_step": hx.numel() * 4 + hy.numel() * 8, "d2h_bytes_per_step": 4,
                            "what": "pinned HOST buffers -> device inside the timed region, every step: the host-padded (B, T_pad, D) "
                                    "tensor + labels exactly as the reference's collate builds them (train.py:183-205,301-302), "
                                    "copied on a copy stream under the previous step; the loss is read back every step"},
        "e2e_resident_feed": {"value": valid_global * K / t_e2e_store, "unit": UNIT, "ms_per_step": t_e2e_store / K * 1e3,
                              "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 4,
                              "what": "DeviceFeatureStore.pad_batch(indices) gathers each step's batch from the HBM-resident "
                                      "dataset into the graph's input buffers (on a side stream under the previous step)"},
        "gpu_launches": launches_per_step * K,
        "roofline": roof,
        "clocks": clocks, "loss": last_loss,
    }
    best = "e2e_padded_host" if line["e2e_padded_host"]["value"] >= line["e2e_ragged_host"]["value"] else "e2e_ragged_host"
    line["e2e"] = dict(line[best], feed=best, h2d_copy_rate_gbs_per_rank_all_ranks_copying=h2d_gbs)
    if cpu_entry is not None:
        line["cpu_baseline"] = cpu_entry
    if wl["padded_global"]:
        line["padded_frames_per_s"] = wl["padded_global"] * K / t_dev
    if eager is not None:
        line["gpu_eager_baseline"] = eager
    print(json.dumps(line), flush=True)
    shutdown()


def run_config5(args):
    """BASELINE configs[4]: inference-only ensemble (inference.py:113-179 with .eval()): 2 checkpoints, 32 videos with
    the segment.txt length distribution, batch 1 per call: forward x2 -> per-frame argmax -> segment vote -> mode."""
    import torch
    from pytorch_video_action_b200 import MultiStageModel, frame_argmax, segment_vote, ensemble_vote
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    nets = []
    for ws in (0, 1):
        torch.manual_seed(ws)
        nets.append(MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).eval())
    g = torch.Generator().manual_seed(55)
    # 32 lengths spread like segment.txt's spans (min 156 / median 1240 / p90 3892 / max 8191), 2..16 segments each
    spans = [156, 400, 575, 705, 775, 820, 900, 960, 1010, 1100, 1180, 1240, 1240, 1300, 1382, 1500, 1620, 1700, 1850, 2000,
             2200, 2400, 2600, 2800, 3100, 3400, 3600, 3892, 4007, 4123, 6000, 8191]
    videos = []
    for T in spans:
        nseg = int(torch.randint(2, 17, (1,), generator=g))
        cuts = sorted(set(int(v) for v in torch.randint(1, T, (nseg - 1,), generator=g)))
        videos.append((torch.randn(1, T, DIM, generator=g).mul_(3.0).to(dev), [0] + cuts + [T]))
    frames = sum(spans)

    from pytorch_video_action_b200 import ensemble_predict

    def run_all():
        # batch 1 per call with T = the video's own length (inference.py:78); votes gathered on the device, one D2H
        return ensemble_predict(nets, [x for x, _ in videos], [seg for _, seg in videos], NCLASS)
    W, K = max(args.warmup, 3), args.steps
    for _ in range(W):
        run_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        run_all()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    line = {"metric": "mstcn_inference_ensemble_frames_per_sec", "value": frames * K / t, "unit": "video frames/s (each through both checkpoints)",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": t / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32-equivalent (3xTF32 on tcgen05)", "data": "synthetic",
            "config": {"workload": f"MS-TCN {STAGES}x{LAYERS}x{FMAPS}, K={NCLASS}, D={DIM}: inference ensemble of 2 checkpoints over 32 "
                                   "videos (segment.txt-shaped lengths), batch 1 per call, argmax + segment vote on the device, votes "
                                   "read back with one D2H per pass, mode on the host (BASELINE configs[4])", "launch": "host launches (variable T per call)"},
            "frames_per_pass": frames}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager-gpu"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE workload (SURVEY.md 8d numbering)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 8 videos per rank (default); strong = fixed global batch of 64 videos sharded 64/N")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "peer", "nccl"],
                    help="data parallel gradient sum: peer = mstcn_dp_allreduce over NVLink peer memory (default when the\n"
                         "ranks can map each other's buffers), nccl = ncclAllReduce")
    ap.add_argument("--nvls", default="auto", choices=["auto", "on", "off"],
                    help="peer all-reduce through the NVSwitch multicast mapping (multimem.ld_reduce / multimem.st); auto: when available")
    ap.add_argument("--dp-overlap", default="auto", choices=["auto", "on", "off"],
                    help="per-stage gradient buckets summed under the rest of the backward (auto = off: one sum behind the backward measured fastest)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from the host instead of CUDA-graph replay")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the gpu_eager_baseline leg of the default line")
    ap.add_argument("--fp32-ffma", action="store_true",
                    help="run the dilated layers on the exact fp32 FFMA kernels instead of tcgen05 3xTF32")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "eager-gpu":
        run_eager_gpu(args)
    elif args.config == 5:
        run_config5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
