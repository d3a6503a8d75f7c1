"""Does the slim InstanceNorm backward (tg_in_stream_slim) share an SM with a persistent weight-gradient GEMM?
Per level of UNet++ (batch 32): the weight gradient alone, the statistics + apply passes alone (fat ring, slim ring),
and both together -- the GEMM queued first on a side stream, the passes on the main stream -- for either form.
    python tools/slim_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tactile_gan_b200 import _C  # noqa: E402
from tactile_gan_b200._C import F, ptr  # noqa: E402

dev = "cuda"
TAPS = [(r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]


def timeit(fn, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def level(n, h, w, c, cin):
    """tensor of the passes: n x h x w x c; the GEMM: weight gradient of a 3x3 conv cin -> c on the same map"""
    x = torch.randn(n, h, w, cin, device=dev).bfloat16()
    dy = torch.randn(n, h, w, c, device=dev).bfloat16()
    dw = torch.zeros(9, c, cin, device=dev)
    plan = _C.wgrad_plan([x], dy, TAPS, dw)
    raw = torch.randn(n, h, w, c, device=dev).bfloat16()
    g1 = torch.randn(n, h, w, c, device=dev).bfloat16()
    dz = torch.zeros_like(raw)
    mr = torch.rand(n, c, 2, device=dev) + 0.5
    red = torch.zeros(n, c, 2, device=dev)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    s2 = torch.cuda.Stream()

    def tail():
        _C.call("in_bwd_reduce", ptr(raw), ptr(raw), ptr(mr), ptr(gamma), ptr(beta), ptr(g1), None, 0, None, 0, None,
                ptr(red), n, h, w, c, c, 3, F(0.0))
        _C.call("in_bwd_apply_re", ptr(raw), ptr(raw), ptr(mr), ptr(gamma), ptr(beta), ptr(g1), None, 0, None, 0, ptr(red),
                ptr(dz), n, h, w, c, c, 3, F(0.0), None, None)

    def both():
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(s2):
            s2.wait_event(ev)
            plan.run()
            ev2 = torch.cuda.Event()
            ev2.record()
        tail()
        torch.cuda.current_stream().wait_event(ev2)

    lib = _C.lib()
    t_w = timeit(plan.run)
    out = [f"{n}x{h}x{w}x{c} (wgrad {cin}->{c}): wgrad {t_w:6.1f} us"]
    for slim in (0, 1):
        lib.tg_in_stream_slim(slim)
        t_t = timeit(tail)
        t_b = timeit(both)
        out.append(f"{'slim' if slim else 'fat '}: passes {t_t:6.1f}  together {t_b:6.1f}  (sum {t_w + t_t:6.1f})")
    lib.tg_in_stream_slim(0)
    print("   ".join(out), flush=True)


if __name__ == "__main__":
    print("TG_SLIM_KB =", os.environ.get("TG_SLIM_KB", "32"), " TG_STREAM =", os.environ.get("TG_STREAM", "1"))
    level(32, 256, 256, 64, 64)
    level(32, 256, 256, 64, 192)
    level(32, 128, 128, 128, 128)
    level(32, 128, 128, 128, 384)
    level(32, 64, 64, 256, 256)
    level(32, 32, 32, 512, 512)
