"""Does a bandwidth-bound tail kernel co-run with a persistent implicit-GEMM kernel? (diagnostic)
Times {GEMM alone, tail alone, both back to back on one stream, both on two streams}."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tactile_gan_b200 import _C  # noqa: E402
from tactile_gan_b200._C import F, ptr  # noqa: E402

dev = "cuda"
n, h, w, c = 32, 256, 256, 64


def gemm_plan(kind):
    x = torch.randn(n, h, w, c, device=dev).bfloat16()
    taps = [(r - 1, s - 1, r * 3 + s) for r in range(3) for s in range(3)]
    if kind == "conv":
        wt = torch.randn(9, 64, c, device=dev).bfloat16()
        out = torch.zeros(n, h, w, 64, device=dev, dtype=torch.bfloat16)
        return _C.conv_plan([dict(act=x, wgt=wt)], out, taps)
    dy = torch.randn(n, h, w, 64, device=dev).bfloat16()
    dw = torch.zeros(9, 64, c, device=dev)
    return _C.wgrad_plan([x], dy, taps, dw)


def main():
    raw = torch.randn(n, h, w, c, device=dev).bfloat16()
    dn = torch.randn(n, h, w, c, device=dev).bfloat16()
    dz = torch.zeros_like(raw)
    mr = torch.rand(n, c, 2, device=dev) + 0.5
    red = torch.randn(n, c, 2, device=dev)
    gamma = torch.ones(c, device=dev)
    y = torch.zeros_like(raw)

    def apply():
        _C.call("in_bwd_apply", ptr(dn), ptr(raw), ptr(mr), ptr(gamma), ptr(red), ptr(dz), n, h * w, c, c, None, None)

    def act_fwd():
        _C.call("in_act_fwd", ptr(raw), ptr(mr), ptr(gamma), None, ptr(y), None, 0, None, n, h, w, c, c, 3, F(0.0))

    s2 = torch.cuda.Stream()

    def timeit(fn, reps=10):
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

    for kind in ("conv", "wgrad"):
        plan = gemm_plan(kind)
        for tname, tail in (("in_bwd_apply", apply), ("in_act_fwd", act_fwd)):
            def both_serial():
                plan.run()
                tail()

            def both_overlap():
                ev = torch.cuda.Event()
                ev.record()
                plan.run()
                with torch.cuda.stream(s2):
                    s2.wait_event(ev)
                    tail()
                    ev2 = torch.cuda.Event()
                    ev2.record()
                torch.cuda.current_stream().wait_event(ev2)

            def tail_first_overlap():
                ev = torch.cuda.Event()
                ev.record()
                with torch.cuda.stream(s2):
                    s2.wait_event(ev)
                    tail()
                    ev2 = torch.cuda.Event()
                    ev2.record()
                plan.run()
                torch.cuda.current_stream().wait_event(ev2)

            print(f"{kind:5s} + {tname:12s}: gemm {timeit(plan.run):7.1f} us  tail {timeit(tail):7.1f} us  "
                  f"serial {timeit(both_serial):7.1f} us  overlap(gemm first) {timeit(both_overlap):7.1f} us  "
                  f"overlap(tail first) {timeit(tail_first_overlap):7.1f} us")


if __name__ == "__main__":
    main()
