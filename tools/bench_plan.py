"""Micro-benchmark of single implicit-GEMM plans (diagnostic): python tools/bench_plan.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tactile_gan_b200 import _C  # noqa: E402

dev = "cuda"


def timeit(plan, reps=20):
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def conv_case(n, h, w, cin, cout, taps, stride=1, bias=False, act=0, pad=None):
    k = int(round(taps ** 0.5))
    pad = (k // 2) if pad is None else pad
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    x = torch.randn(n, h, w, cin, device=dev).bfloat16()
    wt = torch.randn(taps, cout, cin, device=dev).bfloat16()
    out = torch.zeros(n, ho, wo, cout, device=dev, dtype=torch.bfloat16)
    b = torch.randn(cout, device=dev) if bias else None
    tp = [(r - pad, s - pad, r * k + s) for r in range(k) for s in range(k)]
    plan = _C.conv_plan([dict(act=x, wgt=wt)], out, tp, stride=stride, bias=b, act=act)
    ms = timeit(plan)
    fl = 2.0 * n * ho * wo * taps * cin * cout
    print(f"conv n{n} {h}x{w} cin{cin} cout{cout} taps{taps} s{stride} bias={int(bias)} act={act}: {ms*1e3:8.1f} us "
          f"{fl/ms/1e9:8.1f} TF/s  {(x.numel()+out.numel())*2/ms/1e6:7.1f} GB/s")


def wgrad_case(n, h, w, cin, cout, k=3, pad=1):
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    x = torch.randn(n, h, w, cin, device=dev).bfloat16()
    dy = torch.randn(n, ho, wo, cout, device=dev).bfloat16()
    dw = torch.zeros(k * k, cout, cin, device=dev)
    tp = [(r - pad, s - pad, r * k + s) for r in range(k) for s in range(k)]
    plan = _C.wgrad_plan([x], dy, tp, dw)
    ms = timeit(plan)
    fl = 2.0 * n * ho * wo * k * k * cin * cout
    print(f"wgrad n{n} {h}x{w} cin{cin} cout{cout} k{k} p{pad}: {ms*1e3:8.1f} us {fl/ms/1e9:8.1f} TF/s")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "wgrad":
        for n in (16, 32, 48, 64):
            wgrad_case(n, 61, 61, 256, 512, pad=0)
        wgrad_case(64, 63, 63, 128, 256, pad=0)
        wgrad_case(32, 64, 64, 256, 256)
        wgrad_case(32, 32, 32, 512, 512)
        wgrad_case(32, 256, 256, 64, 64)
        wgrad_case(32, 256, 256, 384, 64)
        wgrad_case(32, 128, 128, 128, 128)
        wgrad_case(32, 128, 128, 640, 128)
        sys.exit(0)
    for args in [(32, 127, 127, 64, 64, 1), (32, 128, 128, 64, 64, 1), (32, 127, 127, 64, 128, 1),
                 (32, 59, 59, 512, 64, 1), (32, 59, 59, 64, 512, 1), (32, 256, 256, 64, 64, 1)]:
        conv_case(*args)
    conv_case(32, 127, 127, 64, 64, 1, bias=True, act=1)
    conv_case(32, 127, 127, 64, 64, 1, bias=True, act=0)
    conv_case(32, 127, 127, 64, 64, 1, bias=False, act=1)
    conv_case(32, 256, 256, 64, 64, 9, bias=True, act=3)
    conv_case(32, 256, 256, 64, 64, 9)
    conv_case(32, 256, 256, 64, 128, 9)
    conv_case(32, 128, 128, 128, 128, 9)
