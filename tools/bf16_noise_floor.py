"""Where the fake_B difference between the CUDA path and the oracle comes from. CPU only.

Experiment: run the oracle's UNet++ forward with bf16 rounding at the CUDA path's storage points (oracle.QUANT) twice --
once as is, once with the INPUT perturbed by 1e-6 relative (the size of an fp32 summation-order difference). Any two
bf16-storage implementations of the same network differ at least that much before the first rounding. If the rounding
noise of deep conv+InstanceNorm stacks were stable, the two outputs would agree to ~1e-6; instead the roundings
decorrelate after a few layers and the outputs differ by about what CUDA-vs-oracle differ.
    python tools/bf16_noise_floor.py [nf] [size]
"""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import oracle as orc  # noqa: E402

if __name__ == "__main__":
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_dict_keys.pt"))
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    for gen in ("UNet++", "UNet", "BCDUNet"):
        g = torch.Generator().manual_seed(21)
        shp = OrderedDict((k, tuple(s if (i > 1 or len(v) < 4) else max(1, s * nf // 64) if s % 64 == 0 else s
                                    for i, s in enumerate(v))) for k, v in shapes[gen].items())
        sd = orc.init_state_dict(shp if nf != 64 else shapes[gen], g)
        sd = OrderedDict((k, v) for k, v in sd.items() if not k.startswith("clstm"))
        x = torch.rand(1, 3, size, size, generator=g) * 2 - 1
        with torch.no_grad():
            y32 = orc.gen_forward(gen, sd, x)
            orc.QUANT["on"] = True
            yq = orc.gen_forward(gen, sd, x)
            yq2 = orc.gen_forward(gen, sd, x * (1 + 1e-6))
            orc.QUANT["on"] = False
            y32b = orc.gen_forward(gen, sd, x * (1 + 1e-6))
        print(f"{gen:8s} nf={nf} {size}^2: bf16-storage vs fp32 {rel(yq, y32):.4f} | bf16-storage vs bf16-storage, input "
              f"*(1+1e-6) {rel(yq2, yq):.4f} | fp32 vs fp32, same perturbation {rel(y32b, y32):.2e} | "
              f"|fake_B| rms {y32.pow(2).mean().sqrt().item():.4f}", flush=True)
        # the same experiment on the parameter gradient of sum(out * G), with the ReLU masks / max-pool routes of run A
        # forced onto run B (what tests/parity_util.py does between the CUDA forward and the oracle): what is left is
        # the decorrelated bf16 rounding of two runs that differentiate the SAME piecewise-linear map
        import torch.nn.functional as Fn
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from parity_util import MaskFeed, PoolFeed
        gout = torch.randn(1, 3, size, size, generator=g) * 0.01
        names = [k for k, v in sd.items()]

        def grads(x_in, relu, pool):
            p = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
            orc.ACT["relu"], orc.POOL["max"] = relu, pool
            try:
                y = orc.gen_forward(gen, p, x_in)
                gr = torch.autograd.grad(y, [p[k] for k in names], gout, allow_unused=True)
            finally:
                orc.ACT["relu"], orc.POOL["max"] = Fn.relu, orc.max_pool_2x2
            return torch.cat([t.flatten() for t in gr if t is not None])

        masks, pools = [], []

        def rec_relu(t):
            y = Fn.relu(t)
            masks.append(y.detach() > 0)
            return y

        def rec_pool(t):
            o, i = Fn.max_pool2d(t, 2, 2, return_indices=True)
            pools.append(i)
            return o

        orc.QUANT["on"] = True
        ga = grads(x, rec_relu, rec_pool)
        gb = grads(x * (1 + 1e-6), MaskFeed(masks), PoolFeed(pools))
        orc.QUANT["on"] = False
        g32 = grads(x, MaskFeed(masks), PoolFeed(pools))
        print(f"         flat parameter gradient: bf16-storage run B (input *(1+1e-6), run A's masks) vs run A "
              f"{rel(gb, ga):.4f} | fp32 with run A's masks vs run A {rel(g32, ga):.4f}", flush=True)
