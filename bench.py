"""MS-TCN train-step benchmark (BASELINE.json metric: train frames/sec, fwd+bwd; % HBM roofline per
layer kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N == 1 workload = BASELINE configs[1]: 4 stages x 10 layers x 64 ch, 48 classes, D=400, batch of 8
padded/masked videos (T_pad=4000, lens from segment.txt quantiles, 21 132 valid frames), train mode
(dropout on), synthetic N(0,1) features and piecewise-constant labels, default-init weights.
N > 1 (torchrun, one rank per GPU): every rank runs that batch (different feature seeds) = configs[2]
(global batch 8N videos), gradients summed with a bucketed NCCL all-reduce overlapped with backward
-> "scaling": "weak".

A step = zero_grad -> forward -> CrossEntropy(ignore_index=-1) -> backward; the
Adam step is timed separately (`with_adam`).  `value` has inputs resident in HBM; `e2e` goes through the
public API with pinned HOST buffers (H2D of features+labels and D2H of the loss inside the timed region).
--impl reference times the reference's CPU path (torch-CPU port in oracle/torch_port.py, since
/root/reference does not exist on the GPU box) on the host cores.
"""
import argparse
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


def synth_batch(lens, dim, n_class, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    B, T = len(lens), max(lens)
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


def cpu_port_frames_per_s(steps, warmup, threads=None):
    """The reference's CPU path (torch-CPU port) on a bounded sample of the workload: the 2000-frame
    video of the config-2 batch alone (B=1, T=2000, D=400 = BASELINE configs[0]), train mode."""
    import torch
    from oracle import torch_port as TP
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    P = TP.make_params(DIM, STAGES, LAYERS, NCLASS, seed=0)
    for p in P.values():
        p.requires_grad_(True)
    x, y = synth_batch([2000], DIM, NCLASS, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        TP.train_step(P, x, [2000], y, STAGES, LAYERS, NCLASS, train=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return 2000.0 / med, 2000.0 / min(times), med, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    fps, best, med, threads = cpu_port_frames_per_s(steps, max(args.warmup, 1))
    sample = "B=1 T=2000 D=400 video of the config-2 batch (= BASELINE configs[0]), fwd+CE+bwd, dropout on"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(args.warmup, 1), "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MS-TCN 4x10x64, K=48, D=400, B=8 padded videos T_pad=4000 (configs[1]); "
                               "reference arm steps a bounded B=1,T=2000 sample of it on the host CPU"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "best": best},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_layer_kernel(net, x, lens, steps):
    """Average launch duration of the fused dilated-residual forward kernel (the roofline target) over
    the 40 (stage, layer) launches of the timed configuration, CUDA events on the launch stream."""
    import ctypes as C
    import torch
    from pytorch_video_action_b200 import _cabi
    lib = _cabi.lib()
    B, T = len(lens), max(lens)
    N = B * T
    a = torch.randn(N, 64, device=x.device)
    yb = torch.empty_like(a)
    hb = torch.empty_like(a)
    lens_dev = net._lens_device(lens, x.device)
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
    return e0.elapsed_time(e1) * 1e-3 / (reps * len(lay_off))


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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pytorch_video_action_b200 import MultiStageModel, FrameCrossEntropy, FusedAdam, GraphedTrainStep
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

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    torch.manual_seed(0)
    net = MultiStageModel(DIM, STAGES, LAYERS, FMAPS, NCLASS).to(dev).train()
    net.tensor_cores = not args.fp32_ffma
    crit = FrameCrossEntropy()
    opt = FusedAdam(net, lr=1e-3)
    dp = DataParallelMSTCN(net, crit, overlap=args.dp_bucket_overlap) if world > 1 else None
    valid_local = sum(LENS)
    valid_global = valid_local * world
    T = max(LENS)

    host = [synth_batch(LENS, DIM, NCLASS, 1234 + 100 * rank + i) for i in range(N_ROTATE)]
    host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]

    note("inputs resident")
    # device-side halves of the end-to-end H2D double buffer (below)
    dbuf = [(torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])) for _ in range(2)]
    graphed = None
    if not args.no_graph:
        # the whole step (fwd + CE + bwd, incl. the NCCL bucket all-reduces when world > 1) replayed as one CUDA graph;
        # one capture per input buffer (the resident batches and the two double-buffer halves), so a replay reads its
        # inputs where they already are
        graphed = GraphedTrainStep(net, crit, LENS, resident[0][0], resident[0][1], n_valid=valid_global, dp=dp,
                                   inputs=resident + dbuf)
    slot_of = {id(t[0]): i for i, t in enumerate(resident + dbuf)}

    note("graph captured" if graphed is not None else "eager mode")

    def step(x, y, with_adam=False):
        if graphed is not None:
            loss = graphed.replay(slot_of[id(x)])
        else:
            opt.zero_grad()
            if dp is not None:
                loss = dp.forward_backward(x, LENS, y, valid_global)
            else:
                loss = crit(net(x, LENS), y)
                loss.backward()
        if with_adam:
            opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

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

    def prefetch(i):
        hx, hy = host[i % N_ROTATE]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])          # the step that last read this buffer pair is done
            dbuf[i % 2][0].copy_(hx, non_blocking=True)
            dbuf[i % 2][1].copy_(hy, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_run(n):
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

    for e in consumed:
        e.record()
    e2e_run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(K)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t)
    note("e2e timed")
    t_adam = timed(lambda i: step(*resident[i % N_ROTATE], with_adam=True), K)
    last_loss = float(step(*resident[0]).item())

    def shutdown():
        """Tear the process group down.  Captured graphs that hold NCCL kernels must be released first, and a
        communicator teardown that still stalls (seen with graph-captured collectives) must not hang the job."""
        nonlocal graphed
        if world == 1:
            return
        graphed = None
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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath) and not args.fp32_ffma:
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
    if args.fp32_ffma:
        t_layer = time_layer_kernel(net, resident[0][0], LENS, K)
        algo_bytes = 512.0 * valid_local            # 512 B per frame-layer, valid frames only
    else:
        t_layer = time_stage_chain(net, resident[0][0], LENS, K)
        algo_bytes = 512.0 * valid_local * LAYERS   # one chain launch = all layers of a stage
    achieved = algo_bytes / t_layer / 1e9
    cpu_fps, cpu_best, _, cpu_threads = cpu_port_frames_per_s(10, 2)

    if args.fp32_ffma:
        launches_per_step = 1 + 1 + STAGES * LAYERS + STAGES + 2 + 2 * STAGES + 4 * STAGES * LAYERS + 2
    else:
        # operand packing x2; forward: projection, per stage (chain + tail); fused loss head + its finalize; backward:
        # per stage (tail, top-layer gu, chain, layer-0 gx, weight gradients, two reductions), projection gradient +
        # reduction.  (Host-launch mode adds stage max, the separate CE kernels, gradient routing and torch's glue.)
        launches_per_step = 2 + 1 + 2 * STAGES + 2 + 7 * STAGES + 2
    hx, hy = host[0]
    line = {
        "metric": METRIC, "value": valid_global * K / t_dev, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (fp32 FFMA)" if args.fp32_ffma else "f32-equivalent (3xTF32 on tcgen05 for every forward / input-gradient GEMM, exact 4-term tf32 for weight gradients; fp32 FFMA only for the projection weight gradient)",
        "data": "synthetic",
        "config": {"workload": f"MS-TCN {STAGES}x{LAYERS}x{FMAPS}, K={NCLASS}, D={DIM}, per-GPU batch 8 padded/masked "
                               f"videos T_pad={T} lens={LENS} (BASELINE configs[1]; x{world} ranks = configs[2]), "
                               "train mode (dropout on), fwd+CE+bwd",
                   "global_batch_videos": 8 * world, "valid_frames_per_step": valid_global,
                   "padded_frames_per_step": 8 * T * world, "parallelism": f"dp{world}" + ("" if world == 1 else (" (bucketed all-reduce under the backward)" if args.dp_bucket_overlap else " (one gradient all-reduce after the backward)")),
                   "launch": "host launches" if args.no_graph else "CUDA-graph replay of the step (one capture per resident input buffer: replays read the inputs in place)",
                   "l2": f"{N_ROTATE} resident input batches rotated (205 MB > 126 MB L2); "
                         "0.8 GB of saved activations stream through per step, no explicit flush"},
        "padded_frames_per_s": 8 * T * world * K / t_dev,
        "with_adam": {"value": valid_global * K / t_adam, "unit": UNIT, "ms_per_step": t_adam / K * 1e3},
        "e2e": {"value": valid_global * K / t_e2e, "unit": UNIT, "ms_per_step": t_e2e / K * 1e3,
                "h2d_bytes_per_step": hx.numel() * 4 + hy.numel() * 8, "d2h_bytes_per_step": 4},
        "gpu_launches": launches_per_step * K,
        "roofline": {"kernel": ("layer_fwd_kernel (fused dilated residual layer, fp32 FFMA)" if args.fp32_ffma else
                                "tc_layer_kernel<0> chain launch (the 10 fused dilated residual layers of a stage in one "
                                "persistent kernel, tcgen05 3xTF32 + TMA + TMEM)"),
                     "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "peak_kind": peak_kind, "traffic": traffic, "avg_launch_us": t_layer * 1e6,
                     "algorithmic_bytes_per_launch": algo_bytes},
        "cpu_baseline": {"value": cpu_fps, "unit": UNIT, "cores": cpu_threads, "kind": "port", "best": cpu_best,
                         "sample": "B=1 T=2000 D=400 video of the config-2 batch (BASELINE configs[0]), "
                                   "fwd+CE+bwd, dropout on, torch-CPU port of the reference"},
        "clocks": clocks, "loss": last_loss,
    }
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dp-bucket-overlap", action="store_true",
                    help="data parallel: all-reduce one gradient bucket per stage under the rest of the backward instead of\n"
                         "one all-reduce after it (measured 1.4 %% slower at this step time: the NCCL kernels take SMs\n"
                         "from the chain launches)")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from the host instead of CUDA-graph replay")
    ap.add_argument("--fp32-ffma", action="store_true",
                    help="run the dilated layers on the exact fp32 FFMA kernels instead of tcgen05 3xTF32")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
