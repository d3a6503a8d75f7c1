"""Measured parity of the CUDA path against the oracle at BASELINE.json's shapes (nf=64, 256^2) -- prints the numbers
the tolerances of tests/test_baseline_shapes_gpu.py are set from. Run on the GPU box:
    python tools/parity_report.py > gpurun_out/parity_report.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import parity_util as pu  # noqa: E402

if __name__ == "__main__":
    quick = "--quick" in sys.argv
    torch.set_num_threads(os.cpu_count())
    print("host threads", torch.get_num_threads(), "device", torch.cuda.get_device_name(0))
    cases = [dict(gen="UNet++", batch=2), dict(gen="UNet++", batch=4), dict(gen="UNet", batch=4),
             dict(gen="BCDUNet", batch=2), dict(gen="BCDUNet", batch=2, binary_target=True),
             dict(gen="UNet++", batch=2, regularize=False), dict(gen="UNet++", batch=2, lambda_gp=0.0),
             dict(gen="UNet++", batch=2, lambda_per=0.0), dict(gen="UNet++", batch=2, label_smoothing=False),
             dict(gen="UNet++", batch=2, loss="ce", label_smoothing=False), dict(gen="UNet++", batch=2, loss="ce"),
             dict(gen="UNet++", batch=2, loss="hinge"), dict(gen="UNet++", batch=2, loss="w")]
    if quick:
        cases = [dict(c, size=64, nf=16) for c in cases[:4]]
    for c in cases:
        print(pu.fmt(pu.run_step_case(**c)), flush=True)
    print("forward UNet++ B=1 512^2 nf=64: rel-l2 %.5f max-abs %.5f" % pu.run_forward_case("UNet++", 1, 512), flush=True)
    print("forward BCDUNet B=64 256^2 nf=64, samples 0/31/63: rel-l2 %.5f max-abs %.5f"
          % pu.run_forward_case("BCDUNet", 64, 256, samples=(0, 31, 63)), flush=True)
    print("forward UNet B=4 256^2 nf=64: rel-l2 %.5f max-abs %.5f" % pu.run_forward_case("UNet", 4, 256), flush=True)
    for resync in (True, False):
        print("trajectory UNet++ nf=32 128^2 B=2, 8 steps, resync=%s (cuda/oracle)" % resync)
        print(pu.fmt_traj(pu.run_trajectory(nf=32, size=128, resync=resync)), flush=True)
    print("trajectory UNet++ nf=64 256^2 B=2, 4 steps, resync=True")
    print(pu.fmt_traj(pu.run_trajectory(nf=64, size=256, steps=4, resync=True)), flush=True)
