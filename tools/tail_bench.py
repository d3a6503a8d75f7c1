"""HBM roofline micro-benchmark of the InstanceNorm passes at the step's shapes (one process per TG_STREAM setting):
    TG_STREAM=1 python tools/tail_bench.py ; TG_STREAM=0 python tools/tail_bench.py
Each launch works on a different buffer set (3 sets, > L2 together), CUDA events over 12 launches."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tactile_gan_b200 import _C  # noqa: E402
from tactile_gan_b200._C import F, ptr  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6464.9


NCU = "--ncu" in sys.argv          # one launch per pass of the first shape only (for `ncu --set full`)
SLIM = "--slim" in sys.argv        # the backward passes in the slim form (one CTA per SM, 4 KiB chunks), alone on the GPU


def bench(fn, nbytes, reps=12):
    if NCU:
        reps = 1
    for i in range(0 if NCU else 3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    return us, nbytes / us / 1e3


if __name__ == "__main__":
    dev = "cuda"
    shapes = [(32, 256, 256, 64), (32, 128, 128, 128), (32, 64, 64, 256), (32, 32, 32, 512), (32, 16, 16, 1024),
              (64, 63, 63, 128), (64, 61, 61, 256), (64, 59, 59, 512), (4, 256, 256, 64), (4, 2, 2, 512)]
    print("TG_STREAM =", os.environ.get("TG_STREAM", "1"), " slim" if SLIM else "", " peak", PEAK, "GB/s")
    if SLIM:
        _C.lib().tg_in_stream_slim(1)
    tot_t, tot_b = 0.0, 0.0
    for n, h, w, c in (shapes[:1] if NCU else shapes):
        sets = 3 if n * h * w * c * 2 * 3 < 3e9 else 2
        bufs = [[torch.randn(n, h, w, c, device=dev).bfloat16() for _ in range(4)] for _ in range(sets)]
        mr = torch.rand(n, c, 2, device=dev) + 0.5
        red = torch.zeros(n, c, 2, device=dev)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        size = n * h * w * c * 2

        def fwd(i):
            b = bufs[i % sets]
            _C.call("in_act_fwd", ptr(b[0]), ptr(mr), ptr(gamma), ptr(beta), ptr(b[1]), None, 0, None, n, h, w, c, c, 3,
                    F(0.0))

        def reduce1(i):
            b = bufs[i % sets]
            _C.call("in_bwd_reduce", ptr(b[0]), ptr(b[1]), ptr(mr), ptr(gamma), ptr(beta), ptr(b[2]), None, 0, None, 0,
                    None, ptr(red), n, h, w, c, c, 3, F(0.0))

        def reduce2(i):
            b = bufs[i % sets]
            _C.call("in_bwd_reduce", ptr(b[0]), ptr(b[1]), ptr(mr), ptr(gamma), ptr(beta), ptr(b[2]), None, 0, ptr(b[3]),
                    1, None, ptr(red), n, h, w, c, c, 3, F(0.0))

        def apply1(i):
            b = bufs[i % sets]
            _C.call("in_bwd_apply_re", ptr(b[0]), ptr(b[1]), ptr(mr), ptr(gamma), ptr(beta), ptr(b[2]), None, 0, None, 0,
                    ptr(red), ptr(b[3]), n, h, w, c, c, 3, F(0.0), ptr(dg), ptr(db))

        upb = [torch.empty(n, 2 * h, 2 * w, c, device=dev, dtype=torch.bfloat16) for _ in range(2)] \
            if (n == 32 and h in (128, 64, 32)) else None

        def fwd_up(i):
            b = bufs[i % sets]
            _C.call("in_act_fwd", ptr(b[0]), ptr(mr), ptr(gamma), ptr(beta), ptr(b[1]), None, 0, ptr(upb[i % 2]), n, h, w,
                    c, c, 3, F(0.0))

        row = [f"{n:3d}x{h:3d}x{w:3d}x{c:4d} ({size / 1e6:6.1f} MB)"]
        passes = [("fwd", fwd, 2), ("reduce", reduce1, 2), ("reduce2", reduce2, 3), ("apply", apply1, 3)]
        if upb is not None:
            passes.append(("fwd+up", fwd_up, 6))
        for name, fn, k in passes:
            us, gbs = bench(fn, k * size)
            row.append(f"{name} {us:7.1f}us {gbs:6.0f} GB/s ({gbs / PEAK:.2f})")
            if n >= 32:
                tot_t += us
                tot_b += k * size
        print("  ".join(row), flush=True)
        del bufs, upb
        torch.cuda.empty_cache()
    print(f"aggregate (step shapes): {tot_b / tot_t / 1e3:.0f} GB/s = {tot_b / tot_t / 1e3 / PEAK:.3f} of peak")
