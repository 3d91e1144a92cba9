"""CPU port of the reference's MS-TCN path on torch-CPU ops -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference's own implementation of this path IS PyTorch on the host (networks.py:298-347 run by
train.py:298-332 when no GPU is present), but /root/reference does not exist on the GPU box, so the
baseline that bench.py times there (`cpu_baseline`, `--impl reference`, kind "port") is this
functional restatement: the same ATen ops in the same order (conv1d -> relu -> conv1d -> dropout ->
residual*mask; conv_out*mask; softmax*mask; cat; permute; max over stages; CrossEntropyLoss), taking a
state_dict instead of nn.Modules.  Pinned against the golden vectors produced by the unmodified
reference (tests/test_oracle.py::test_torch_port_matches_reference).  Only tests/ and bench.py's
baseline legs may import it; the product path never does.
"""
import torch
import torch.nn.functional as F


def make_params(dim, num_stages, num_layers, n_class, num_f_maps=64, seed=0):
    """Default nn.Conv1d init in the reference's construction order (networks.py:299-303,324-327,339-340)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def conv(prefix, cin, cout, k):
        c = nn.Conv1d(cin, cout, k)
        sd[prefix + ".weight"], sd[prefix + ".bias"] = c.weight.detach().clone(), c.bias.detach().clone()

    for si in range(num_stages):
        pre = "stage1." if si == 0 else f"stages.{si - 1}."
        conv(pre + "conv_1x1", dim if si == 0 else n_class, num_f_maps, 1)
        for li in range(num_layers):
            conv(f"{pre}layers.{li}.conv_dilated", num_f_maps, num_f_maps, 3)
            conv(f"{pre}layers.{li}.conv_1x1", num_f_maps, num_f_maps, 1)
        conv(pre + "conv_out", num_f_maps, n_class, 1)
    return sd


def _stage(P, pre, num_layers, x, mask, train):
    out = F.conv1d(x, P[pre + "conv_1x1.weight"], P[pre + "conv_1x1.bias"])                    # networks.py:330
    for li in range(num_layers):
        d = 2 ** li
        h = F.relu(F.conv1d(out, P[f"{pre}layers.{li}.conv_dilated.weight"], P[f"{pre}layers.{li}.conv_dilated.bias"],
                            padding=d, dilation=d))                                            # :344
        o = F.conv1d(h, P[f"{pre}layers.{li}.conv_1x1.weight"], P[f"{pre}layers.{li}.conv_1x1.bias"])   # :345
        o = F.dropout(o, 0.5, training=train)                                                  # :346
        out = (out + o) * mask[:, 0:1, :]                                                      # :347
    return F.conv1d(out, P[pre + "conv_out.weight"], P[pre + "conv_out.bias"]) * mask[:, 0:1, :]         # :333


def forward(P, x, x_len, num_stages, num_layers, n_class, train=False, per_stage=False):
    """MultiStageModel.forward (networks.py:305-320)."""
    x = x.transpose(1, 2)
    mask = torch.zeros(x.shape[0], n_class, max(x_len), dtype=torch.float)
    for i in range(x.shape[0]):
        mask[i, :, :x_len[i]] = 1
    out = _stage(P, "stage1.", num_layers, x, mask, train)
    outputs = out.unsqueeze(0)
    for s in range(num_stages - 1):
        out = _stage(P, f"stages.{s}.", num_layers, F.softmax(out, dim=1) * mask[:, 0:1, :], mask, train)
        outputs = torch.cat((outputs, out.unsqueeze(0)), dim=0)
    outputs = outputs.permute(0, 1, 3, 2)
    outputs = outputs.contiguous().view(outputs.shape[0], outputs.shape[1] * outputs.shape[2], outputs.shape[3])
    if per_stage:
        return outputs                      # (S, B*T, K): what canonical MS-TCN returns (not the reference)
    return torch.max(outputs, 0)[0]


def train_step(P, x, x_len, labels, num_stages, num_layers, n_class, train=True):
    """zero_grad -> forward -> CrossEntropyLoss(ignore_index=-1) -> backward (train.py:305-328)."""
    for p in P.values():
        p.grad = None
    out = forward(P, x, x_len, num_stages, num_layers, n_class, train)
    loss = F.cross_entropy(out, labels, ignore_index=-1)
    loss.backward()
    return out, loss


def ms_tcn_paper_loss(stage_logits, labels, x_len, lam=0.15, tau=4.0):
    """Canonical MS-TCN loss (Farha & Gall, CVPR 2019) in plain torch, autograd-differentiable: the restatement the
    fused kernel is checked against (NOT in the reference -> parity unpinned).  stage_logits (S, B*T, K)."""
    S, N, K = stage_logits.shape
    B = len(x_len)
    T = N // B
    mask = torch.zeros(B, T, dtype=stage_logits.dtype)
    for i, n in enumerate(x_len):
        mask[i, :n] = 1
    total = 0
    for s in range(S):
        z = stage_logits[s]
        total = total + F.cross_entropy(z, labels, ignore_index=-1)
        logp = F.log_softmax(z.view(B, T, K), dim=2)
        diff = torch.clamp((logp[:, 1:] - logp[:, :-1].detach()) ** 2, min=0, max=tau * tau)
        total = total + lam * torch.mean(diff * mask[:, 1:, None])
    return total
