import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def split_golden(g, ppre="p/", gpre="g/"):
    params = {k[len(ppre):]: v for k, v in g.items() if k.startswith(ppre)}
    grads = {k[len(gpre):]: v for k, v in g.items() if k.startswith(gpre)}
    return params, grads


def rel_err(a, b):
    """max|a-b| / max|b| -- the per-tensor relative error the north star's 1e-3 refers to."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / den


CONFIG2_LENS = [4000, 3892, 3600, 3100, 2600, 2000, 1240, 700]      # SURVEY.md 8d config 2 (= bench.py LENS)


def synth_config2(seed=1234, lens=CONFIG2_LENS, dim=400, K=48):
    """The benchmark batch (bench.py synth_batch / tests/golden/make_golden.py synth_config2): regenerated from the
    seed, bit-identical on every machine (torch CPU generator)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    B, T = len(lens), max(lens)
    x = torch.randn(B, T, dim, generator=g)
    y = torch.full((B, T), -1, dtype=torch.long)
    for b, l in enumerate(lens):
        x[b, l:] = 0
        t = 0
        while t < l:
            run = int(torch.randint(30, 401, (1,), generator=g))
            y[b, t:min(l, t + run)] = int(torch.randint(1, K, (1,), generator=g))
            t += run
    return x, y.flatten()


def reference_init_params(dim, S, L, K, seed):
    """state_dict of the drop-in class under manual_seed(seed): equal to the reference's default init bit for bit
    (tests/test_cabi_cpu.py::test_same_seed_init_and_state_dict_keys)."""
    import torch
    from pytorch_video_action_b200 import MultiStageModel
    torch.manual_seed(seed)
    net = MultiStageModel(dim, S, L, 64, K)
    return net, {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
