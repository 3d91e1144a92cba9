"""CPU tier: the reference arm of bench.py (`--impl reference`) honours the driver's JSON contract without a GPU, alone and
under torchrun (rank 0 alone runs and prints; the other ranks exit 0 without work)."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_reference_line(out, n_gpus):
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out                      # exactly ONE JSON line, from rank 0
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mstcn_train_frames_per_sec_fwd_bwd"
    assert d["unit"] == "valid frames/s" and d["higher_is_better"] is True and d["n_gpus"] == n_gpus
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    # the thread count is the fastest of the calibrated candidates, which are reported
    tried = cb["threads_tried_frames_per_s"]
    assert str(cb["cores"]) in tried and "1" in tried and cb["host_cpus"] == max(int(k) for k in tried)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    _check_reference_line(r.stdout, 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_reference_arm_under_torchrun_runs_on_rank_0_only():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    _check_reference_line(r.stdout, 2)
