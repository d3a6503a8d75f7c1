"""Helpers shared by the -m gpu parity tests (all compute calls go through the C-ABI via ctypes)."""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pad64(c):
    return (c + 63) // 64 * 64


def nhwc_pad(x):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, pad64(c), dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def conv_taps(kh, kw, pad):
    return [(r - pad, s - pad, r * kw + s) for r in range(kh) for s in range(kw)]


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().cpu().flatten(), b.float().cpu().flatten(), dim=0).item()


def randomize(net, seed=0):
    """Non-degenerate parameters: N(0,.05) convs, affine ~ (1 +- .1, +- .1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if p.dim() == 4:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif name.endswith("weight"):
                p.copy_(1 + torch.randn(p.shape, generator=g) * 0.1)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
