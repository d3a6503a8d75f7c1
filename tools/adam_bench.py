"""Fused Adam + re-pack on the generators' real parameter sets:  python tools/adam_bench.py  (one B200)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tactile_gan_b200.engine import build_generator_engine  # noqa: E402
from tactile_gan_b200.generators.generators import create_gen  # noqa: E402
from tactile_gan_b200.util import init_weights  # noqa: E402

if __name__ == "__main__":
    for gen in ("UNet++", "UNet", "BCDUNet"):
        net = create_gen(gen, 3, 3, 64, True).cuda()
        init_weights(net)
        size = 256 if gen == "UNet" else 64
        eng = build_generator_engine(gen, net, 1, size, size, True)
        st = eng.store
        st.grad_arena.normal_()
        nparam = sum(p.numel() for p in st.params)
        for _ in range(3):
            st.adam_step(1e-3, 0.9)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            st.adam_step(1e-3, 0.9)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"{gen:8s} {nparam / 1e6:5.1f} M params  {us:7.1f} us/step  {nparam * 32 / us / 1e3:6.0f} GB/s "
              f"(28 B/param + two bf16 packs)", flush=True)
        del eng, net
        torch.cuda.empty_cache()
