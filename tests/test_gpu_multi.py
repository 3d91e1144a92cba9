"""GPU tier, N > 1: data-parallel gradients on real GPUs over NCCL (SURVEY.md 8e parity test).  Spawns
`torchrun --nproc-per-node 2 tools/dp_check.py`: 16 videos of distinct lengths sharded by parallel.shard_videos, each
rank padding to parallel.local_pad_length (so the +1-frame rule of fact 0.5 runs on hardware), eager and
graph-captured steps, bucketed and single all-reduce, gradient accumulation -- against the un-sharded batch on one GPU.
Skipped on a single-GPU box (the CPU tier covers the host logic with gloo, tests/test_parallel_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_match_single_gpu():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "per-bucket grad rel err" in r.stdout
