#!/usr/bin/env python
"""Benchmark of the tactile-gan G+D training step (BASELINE.json metric: train images/sec, UNet++ 256^2).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port, host cores)

One "step" = one full G+D iteration (reference train.py:99-168) on a per-GPU batch of 32 synthetic
256x256 pairs (configs[1]); N > 1 = one rank per GPU, NCCL gradient allreduce, weak scaling.
`value` is timed with inputs resident in HBM; `e2e` goes through TrainStep.step_from_host (pinned host
batch -> H2D -> step -> D2H of the loss scalars). `roofline` is measured live with CUDA events around
every implicit-GEMM launch of the timed steps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMG = {True: 994.0, False: 927.2}   # 2*(3*MAC_G + 15*MAC_D) with GP / without


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = self.samples[0][1]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(mx) if mx.replace('.', '').isdigit() else None, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_baseline(steps, warmup, batch, threads=None):
    """The oracle port (oracle/oracle.py, a restatement of the reference's PyTorch CPU path) on the host
    cores: UNet++ nf=64 + PatchD, version-2 stack with GP, `batch` images of 256^2 per step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_dict_keys.pt"))
    g = torch.Generator().manual_seed(21)
    sd_g = orc.init_state_dict(shapes["UNet++"], g)
    sd_d = orc.init_state_dict(shapes["patch"], g)
    cfg = orc.StepConfig(gen="UNet++", loss="ls", version=2)
    label = orc.make_real_label((batch, 1, 57, 57), True, generator=g)
    og, od = {}, {}
    times = []
    for i in range(warmup + steps):
        a, b = orc.synthetic_batch(g, batch, 256)
        alpha = torch.rand(batch, 1, generator=g)
        t = time.perf_counter()
        orc.train_step(sd_g, sd_d, og, od, a, b, label, alpha, cfg)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return {"value": batch / per_step, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"{steps} G+D steps of UNet++(nf=64)+PatchD, version-2 losses + GP, batch {batch}, 256x256, "
                      f"fp32 torch CPU after {warmup} warm-up", "s_per_step": per_step}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(0, min(args.warmup, 1))
    cb = cpu_baseline(steps, warmup, 1)
    line = {"impl": "reference", "metric": "train images/sec (G+D step) UNet++ 256^2", "value": cb["value"],
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "UNet++ nf=64 + PatchDiscriminator, version-2 loss stack with GP every step, "
                                   "256x256; CPU arm samples batch 1 per step (GPU arm: batch 32 per GPU)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def layer_table(ts, batch, path):
    """Diagnostic (not a bench number): two more steps with a CUDA-event pair around every launch;
    per-shape implicit-GEMM TFLOP/s and per-kernel tail time go to `path` as JSON."""
    from tactile_gan_b200 import _C
    _C.TIMING["records"].clear()
    _C.TIMING["tail_records"].clear()
    _C.TIMING["on"] = _C.TIMING["tail"] = True
    reps = 2
    for _ in range(reps):
        ts.step(*batch)
    torch.cuda.synchronize()
    _C.TIMING["on"] = _C.TIMING["tail"] = False
    gemm, tail = {}, {}
    for kind, flops, a, b, tag in _C.TIMING["records"]:
        t, f, c = gemm.get(tag, (0.0, 0.0, 0))
        gemm[tag] = (t + a.elapsed_time(b), f + flops, c + 1)
    for name, a, b in _C.TIMING["tail_records"]:
        t, c = tail.get(name, (0.0, 0))
        tail[name] = (t + a.elapsed_time(b), c + 1)
    _C.TIMING["records"].clear()
    _C.TIMING["tail_records"].clear()
    out = {"gemm": [{"tag": k, "ms_per_step": t / reps, "launches_per_step": c / reps,
                     "tflops": (f / (t / 1e3) / 1e12) if t > 0 else 0.0, "gflop_per_step": f / reps / 1e9}
                    for k, (t, f, c) in sorted(gemm.items(), key=lambda kv: -kv[1][0])],
           "tail": [{"kernel": k, "ms_per_step": t / reps, "launches_per_step": c / reps}
                    for k, (t, c) in sorted(tail.items(), key=lambda kv: -kv[1][0])]}
    out["gemm_ms_per_step"] = sum(r["ms_per_step"] for r in out["gemm"])
    out["tail_ms_per_step"] = sum(r["ms_per_step"] for r in out["tail"])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


def run_ours(args):
    from tactile_gan_b200 import _C
    from tactile_gan_b200.discriminators.discriminators import create_disc
    from tactile_gan_b200.generators.generators import create_gen
    from tactile_gan_b200.step import TrainStep
    from tactile_gan_b200.util import init_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    B, S = args.batch, args.size
    torch.manual_seed(21)
    netG = create_gen("UNet++", 3, 3, 64, True).to(dev)
    init_weights(netG)
    netD = create_disc("patch", 3, 3, 64, return_filter=True, activation=True).to(dev)
    init_weights(netD)
    ts = TrainStep(netG, netD, B, S, S, loss="ls", version=2, lambda_a=1.0, lambda_gp=0.01, lambda_per=1.0,
                   w_per=(0, .1, .3, .6), lr=1e-3, beta1=0.9, label_smoothing=True)
    g = torch.Generator().manual_seed(21 + rank)
    pool = 2
    host = [(torch.rand(B, 3, S, S, generator=g).mul_(2).sub_(1).pin_memory(),
             torch.rand(B, 3, S, S, generator=g).pin_memory()) for _ in range(pool)]
    devb = [(a.to(dev), b.to(dev)) for a, b in host]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    for i in range(args.warmup):
        ts.step(*devb[i % pool])
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    _C.COUNTERS["launches"] = 0
    _C.TIMING["records"].clear()
    _C.TIMING["on"] = True
    ms = timed(lambda i: ts.step(*devb[i % pool]), args.steps)
    _C.TIMING["on"] = False
    launches = _C.COUNTERS["launches"]
    sampler.stop_flag = True
    losses = ts.loss_dict()
    # per-kernel-kind roofline from the CUDA events recorded around every implicit-GEMM launch
    agg = {}
    for kind, flops, a, b, _tag in _C.TIMING["records"]:
        t, f, c = agg.get(kind, (0.0, 0.0, 0))
        agg[kind] = (t + a.elapsed_time(b), f + flops, c + 1)
    _C.TIMING["records"].clear()
    if args.layers and rank == 0:
        layer_table(ts, devb[0], args.layers)
    # end-to-end through the public call with host buffers
    for i in range(min(2, args.warmup)):
        ts.step_from_host(*host[i % pool])
    ms_e2e = timed(lambda i: ts.step_from_host(*host[i % pool]), args.steps)
    sampler.join(timeout=2)

    if rank == 0:
        hbm, tf_burst, tf_sus, which = peaks()
        imgs = B * world * args.steps
        line = {"metric": "train images/sec (G+D step) UNet++ 256^2", "value": imgs / (ms / 1e3), "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"UNet++ nf=64 + PatchDiscriminator, version-2 loss stack (LSGAN + L1 + pan + GP "
                                       f"every step), batch {B}/GPU, {S}x{S}, random-init weights",
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "per-step working set (>10 GB of activations) exceeds the 126 MB L2; no explicit flush"},
                "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s",
                        "h2d_bytes_per_step": 2 * B * 3 * S * S * 4, "d2h_bytes_per_step": 32},
                "gpu_launches": launches, "clocks": sampler.summary(), "losses_last_step": losses,
                "step_tflops": GFLOP_PER_IMG[True] * B * args.steps / (ms / 1e3) / 1e3}
        if "conv" in agg:
            t, f, c = agg["conv"]
            ach = f / (t / 1e3) / 1e12
            line["roofline"] = {"bound": "tensor", "kernel": "igemm_conv_kernel (conv fwd + dgrad, all shapes)",
                                "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus,
                                "peak_source": f"bf16_tflops_sustained, {which}", "traffic": None,
                                "launches": c, "share_of_step": t / ms}
        if "wgrad" in agg:
            t, f, c = agg["wgrad"]
            ach = f / (t / 1e3) / 1e12
            line["roofline_wgrad"] = {"bound": "tensor", "kernel": "wgrad_kernel", "achieved": ach, "peak": tf_sus,
                                      "unit": "TFLOP/s", "frac": ach / tf_sus, "launches": c, "share_of_step": t / ms}
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline(2, 1, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (configs[1]: 32)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--layers", default="", help="diagnostic: write a per-shape / per-kernel timing table (JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
