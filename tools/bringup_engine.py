"""GPU bring-up of the generator engines against the CPU oracle: forward output and all parameter
gradients for a random output-gradient. Usage: python tools/bringup_engine.py unet++ 64 64 2"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
from tactile_gan_b200.generators.generators import create_gen  # noqa: E402


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "unet++"
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    torch.manual_seed(0)
    net = create_gen(kind, 3, 3, nf, True)
    if os.environ.get("LINEAR") == "1":   # identity activations: isolates graph plumbing from ReLU mask flips
        orc.ACT["relu"] = lambda t: t
        net._tg_debug_act = 0
    for name, p in net.named_parameters():
        if p.dim() == 4:
            torch.nn.init.normal_(p, 0.0, 0.05)
        elif name.endswith("weight"):
            torch.nn.init.normal_(p, 1.0, 0.1)
        else:
            torch.nn.init.normal_(p, 0.0, 0.1)
    sd = OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
    g = torch.Generator().manual_seed(1)
    x, _ = orc.synthetic_batch(g, n, size)
    gout = torch.randn(n, 3, size, size, generator=g) * 0.01
    orc.QUANT["on"] = os.environ.get("QUANT", "1") == "1"
    # oracle (CPU fp32, optionally with the bf16 storage points emulated)
    psd = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items() if not k.startswith("clstm"))
    ref = orc.gen_forward(kind, psd, x, True)
    names = list(psd.keys())
    rg = torch.autograd.grad(ref, [psd[k] for k in names], gout, allow_unused=True)
    rg = dict(zip(names, rg))
    # engine
    net = net.cuda()
    out = net(x.cuda())
    print(f"forward rel_l2={rel(out.cpu(), ref):.3e} max_abs={(out.cpu()-ref).abs().max().item():.3e}", flush=True)
    out.backward(gout.cuda())
    torch.cuda.synchronize()
    worst = 0
    for k, p in net.named_parameters():
        if rg.get(k) is None:
            continue
        r = rel(p.grad.cpu(), rg[k])
        worst = max(worst, r)
        cos = torch.nn.functional.cosine_similarity(p.grad.cpu().flatten(), rg[k].flatten(), dim=0).item()
        if r > 3e-2 or os.environ.get("VERBOSE"):
            print(f"  cos={cos:.4f}", end="")
            print(f"  grad {k:40s} rel_l2={r:.3e} |ref|={rg[k].norm().item():.3e}")
    print(f"backward worst grad rel_l2={worst:.3e}", flush=True)
    from tactile_gan_b200 import _C
    print("error flag", _C.error_flag())


if __name__ == "__main__":
    main()
