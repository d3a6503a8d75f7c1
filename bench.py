#!/usr/bin/env python
"""Benchmark of the tactile-gan G+D training step (BASELINE.json metric: train images/sec, UNet++ 256^2).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (the unmodified reference modules
                                                             # from baseline/_ref on the host cores)

One "step" = one full G+D iteration (reference train.py:99-168) on a per-GPU batch of 32 synthetic
256x256 pairs (configs[1]); N > 1 = one rank per GPU, NCCL gradient allreduce, weak scaling.
`value` is timed with inputs resident in HBM; `e2e` goes through TrainStep.step_from_host (pinned host
batch -> H2D -> step -> D2H of the loss scalars). `cudnn_baseline` (N=1) runs the same reference modules on the same
GPU under eager PyTorch + cuDNN; `comm` (N>1) reports the time the compute stream waited for gradient collectives.
`roofline*` are measured live in the same process with CUDA
events around every implicit-GEMM launch (tensor roofline) and every InstanceNorm tail launch (HBM roofline) of
K further steps run with serialised launches; `traffic` is the DRAM byte count of one representative launch
from the committed ncu capture.

Other workloads (not bench lines; for DESIGN.md / profiles): --gen UNet|BCDUNet, --version 1,
--workload infer (generator forward only, configs[4]).
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

# algorithmic GMAC per 256^2 image. Training = 2*(3*MAC_G + 15*MAC_D) FLOP with GP (9 without);
# version 1 drops the fifth D forward and adds 6 VGG16 forward-equivalents at 224^2; inference = 2*MAC_G.
MAC_G = {"UNet++": 137.833218048, "UNet": 15.717105664, "BCDUNet": 36.767268864}
MAC_D = 5.567061888
MAC_VGG = 13.96


def gflop_per_image(gen, workload, version, regularize=True, size=256):
    s = (size / 256.0) ** 2
    if workload == "infer":
        return 2 * MAC_G[gen] * s
    d = (15 if regularize else 9) - (0 if version == 2 else 1)
    return 2 * ((3 * MAC_G[gen] + d * MAC_D) * s + (6 * MAC_VGG if version != 2 else 0))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons every 200 ms while the timed regions run (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)

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


def _port_runner(gen, batch):
    """Fallback when baseline/_ref did not travel: the oracle port (oracle/oracle.py) of the same step."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    shapes = torch.load(os.path.join(ROOT, "tests", "golden", "state_dict_keys.pt"))
    g = torch.Generator().manual_seed(21)
    sd_g = orc.init_state_dict(shapes[gen], g)
    sd_d = orc.init_state_dict(shapes["patch"], g)
    cfg = orc.StepConfig(gen=gen, loss="ls", version=2)
    label = orc.make_real_label((batch, 1, 57, 57), True, generator=g)
    og, od = {}, {}

    def run(a, b):
        alpha = torch.rand(batch, 1, generator=g)
        orc.train_step(sd_g, sd_d, og, od, a, b, label, alpha, cfg)
    return run, "port"


def cpu_baseline(steps, warmup, batch, threads=None, gen="UNet++"):
    """The reference's own CPU path on the host cores: the UNMODIFIED reference modules from baseline/_ref driven
    through train.py:99-168 (baseline/ref_step.py; kind "reference"), or the oracle port when baseline/_ref is absent
    (kind "port"). `gen` nf=64 + PatchD, version-2 stack with GP, `batch` images of 256^2 per step, fp32."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    from baseline import ref_step
    if ref_step.available():
        rs = ref_step.RefStep(gen, 64, "cpu")
        run, kind = (lambda a, b: rs.step(a, b)), "reference"
    else:
        run, kind = _port_runner(gen, batch)
    g = torch.Generator().manual_seed(21)
    times = []
    for i in range(warmup + steps):
        a = torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1
        b = torch.rand(batch, 3, 256, 256, generator=g)
        t = time.perf_counter()
        run(a, b)
        dt = time.perf_counter() - t
        if i >= warmup:
            times.append(dt)
    per_step = statistics.median(times)
    return {"value": batch / per_step, "unit": "images/s", "cores": threads, "kind": kind,
            "sample": f"{steps} G+D steps of {gen}(nf=64)+PatchD, version-2 losses + GP, batch {batch}, 256x256, "
                      f"fp32 torch CPU after {warmup} warm-up (median step)", "s_per_step": per_step}


def cudnn_baseline(dev, batch, size, steps=3, warmup=2, gen="UNet++"):
    """The bar that matters: the reference's own modules on THIS GPU under PyTorch
    eager + cuDNN, same step / batch / size, inputs resident in HBM, CUDA events over `steps` iterations.
    Modes: stock (fp32 storage, cuDNN TF32 convolutions -- torch's default, what `python train.py` runs), strict fp32,
    and bf16 autocast. Baseline leg only: none of this repo's kernels run here and the product path never imports it."""
    from baseline import ref_step
    if not ref_step.available():
        return {"unavailable": "baseline/_ref not present on this box"}
    out = {"kind": "reference modules, torch %s eager + cuDNN %s" % (torch.__version__, torch.backends.cudnn.version()),
           "batch": batch, "size": size, "steps": steps, "warmup": warmup}
    g = torch.Generator().manual_seed(21)
    a = (torch.rand(batch, 3, size, size, generator=g) * 2 - 1).to(dev)
    b = torch.rand(batch, 3, size, size, generator=g).to(dev)
    keep = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    for mode, tf32, amp in (("bf16_autocast", True, torch.bfloat16), ("tf32_stock", True, None), ("fp32_strict", False, None)):
        try:
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            rs = ref_step.RefStep(gen, 64, dev, autocast=amp)
            for _ in range(warmup):
                rs.step(a, b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                losses = rs.step(a, b)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"images_per_s": batch / (ms / 1e3), "ms_per_step": ms,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "losses_last_step": losses}
        except Exception as e:  # e.g. out of memory at this batch: reported, never fatal for the bench line
            out[mode] = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        finally:
            rs = None
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = keep
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
    return out


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the step on the host cores, on the GPU arm's metric and
    config; each step is a bounded sample of the workload (batch 2 instead of 32 per step -- a batch-32 fp32 CPU step
    takes ~25 s and ~50 GB). Also reports BASELINE.md section 4's prescribed cfg1 (UNet, batch 4)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
    cb = cpu_baseline(steps, warmup, 2)
    cfg1 = cpu_baseline(3, 1, 4, gen="UNet")
    line = {"impl": "reference", "metric": "train images/sec (G+D step) UNet++ 256^2", "value": cb["value"],
            "unit": "images/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": cb["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "UNet++ nf=64 + PatchDiscriminator, version-2 loss stack with GP every step, "
                                   "256x256; CPU arm samples batch 2 per step (GPU arm: batch 32 per GPU)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "cfg1_unet_b4": {k: cfg1[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def layer_table(fn, path, reps=2):
    """Diagnostic (not a bench number): `reps` steps with a CUDA-event pair around every tail launch, then `reps`
    with one around every implicit-GEMM launch; per-shape TFLOP/s and per-kernel tail time go to `path` (JSON)."""
    from tactile_gan_b200 import _C
    _C.TIMING["records"].clear()
    _C.TIMING["tail_records"].clear()
    _C.TIMING["on"], _C.TIMING["tail"] = False, True
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    _C.TIMING["tail"] = False
    tail = {}
    for name, a, b in _C.TIMING["tail_records"]:
        t, c = tail.get(name, (0.0, 0))
        tail[name] = (t + a.elapsed_time(b), c + 1)
    _C.TIMING["tail_records"].clear()
    _C.TIMING["on"] = True
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    _C.TIMING["on"] = False
    gemm = {}
    for kind, flops, a, b, tag in _C.TIMING["records"]:
        if kind.startswith("tail:"):
            continue
        t, f, c = gemm.get(tag, (0.0, 0.0, 0))
        gemm[tag] = (t + a.elapsed_time(b), f + flops, c + 1)
    _C.TIMING["records"].clear()
    out = {"gemm": [{"tag": k, "ms_per_step": t / reps, "launches_per_step": c / reps,
                     "tflops": (f / (t / 1e3) / 1e12) if t > 0 else 0.0, "gflop_per_step": f / reps / 1e9}
                    for k, (t, f, c) in sorted(gemm.items(), key=lambda kv: -kv[1][0])],
           "tail": [{"kernel": k, "ms_per_step": t / reps, "launches_per_step": c / reps}
                    for k, (t, c) in sorted(tail.items(), key=lambda kv: -kv[1][0])]}
    out["gemm_ms_per_step"] = sum(r["ms_per_step"] for r in out["gemm"])
    out["tail_ms_per_step"] = sum(r["ms_per_step"] for r in out["tail"])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


def ncu_traffic():
    """DRAM bytes of one representative launch of the dominant kernel, from the committed `ncu --set full`
    summary (profiles/r02_ncu_top_kernels.json, written by tools/ncu_summary.py --json)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_top_kernels.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_ncu_top_kernels.json")
    if not os.path.exists(p):
        return None, None
    top = json.load(open(p)).get("roofline_launch")
    if not top:
        return None, None
    return top.get("dram_bytes"), top


def ncu_tensor_pipe():
    """Time-weighted tensor-pipe utilisation of all conv GEMM launches of the step (BASELINE.json's second metric), from
    the committed ncu pass profiles/r02_ncu_launch_shares.txt (sm__pipe_tensor_cycles_active over every launch)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_launch_shares.txt")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_ncu_tensor_pipe.txt")
    if not os.path.exists(p):
        return None
    for line in open(p):
        if line.startswith("# all conv GEMMs"):
            try:
                return float(line.split(":")[1].split("%")[0])
            except ValueError:
                return None
    return None


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
        os.environ.setdefault("TG_COMM_PROFILE", "1")     # event pairs around the two waits for gradient collectives
    B, S = args.batch, args.size
    torch.manual_seed(21)
    netG = create_gen(args.gen, 3, 3, 64, True).to(dev)
    init_weights(netG)
    g = torch.Generator().manual_seed(21 + rank)
    pool = 2
    host = [(torch.rand(B, 3, S, S, generator=g).mul_(2).sub_(1).pin_memory(),
             torch.rand(B, 3, S, S, generator=g).pin_memory()) for _ in range(pool)]
    devb = [(a.to(dev), b.to(dev)) for a, b in host]
    train = args.workload == "train"
    ts = None
    if train:
        netD = create_disc("patch", 3, 3, 64, return_filter=args.version == 2, activation=True).to(dev)
        init_weights(netD)
        vgg_blocks = None
        if args.version != 2:
            import warnings
            from tactile_gan_b200.util import VGGPerceptualLoss
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                vgg_blocks = VGGPerceptualLoss(resize=True).blocks
        ts = TrainStep(netG, netD, B, S, S, loss="ls", version=args.version, lambda_a=1.0, lambda_gp=0.01,
                       lambda_per=1.0, w_per=(0, .1, .3, .6), lr=1e-3, beta1=0.9, label_smoothing=True,
                       vgg_blocks=vgg_blocks)

        def step_dev(i):
            ts.step(*devb[i % pool])

        def step_host(i):
            # every step copies its batch from pinned memory (the next batch's copy is already in flight) and reads the
            # five loss scalars of a step back: with a one-step lag, so the launch queue stays fed; the last timed
            # step reads its own losses synchronously -- every step's result is on the host before the timer stops
            ts.step_from_host(*host[i % pool], prefetch=host[(i + 1) % pool], lag=i != args.steps - 1)
        h2d, d2h = 2 * B * 3 * S * S * 4, 32
    else:
        chunked = B > netG.infer_chunk(S, S)     # batch-512 sweeps: the module runs them as fixed-size chunks
        eng = None if chunked else netG._engine(B, S, S, False)
        dev_in = torch.empty(B, 3, S, S, device=dev)
        host_out = torch.empty(B, 3, S, S).pin_memory()

        def step_dev(i):
            if chunked:
                c = netG.infer_chunk(S, S)
                for j in range(0, B, c):
                    m = min(c, B - j)
                    netG._engine(m, S, S, False).forward(devb[i % pool][0][j:j + m])
            else:
                eng.forward(devb[i % pool][0])

        def step_host(i):       # test.py:202-203: out = model(real_A.to(device)).cpu(), replayed from a CUDA graph
            dev_in.copy_(host[i % pool][0], non_blocking=True)
            with torch.no_grad():
                host_out.copy_(netG(dev_in) if chunked else eng.forward_graphed(dev_in), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d = d2h = B * 3 * S * S * 4

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
        step_dev(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms_e2e = None
    if os.environ.get("TG_BENCH_E2E_FIRST"):      # diagnostic: order effects (clock / power drift) between the two timings
        for i in range(min(2, args.warmup)):
            step_host(i)
        ms_e2e = timed(step_host, args.steps)
    _C.COUNTERS["launches"] = 0
    if train and ts.comm_profile is not None:
        ts.comm_profile.clear()
    ms = timed(step_dev, args.steps)
    launches = _C.COUNTERS["launches"]
    comm = None
    if train and ts.comm_profile is not None:
        # time the compute stream spent waiting for NCCL (max over ranks): communication the step did not hide
        tot = {}
        for tag, a, b in ts.comm_profile:
            tot[tag] = tot.get(tag, 0.0) + a.elapsed_time(b)
        t = torch.tensor([tot.get("D", 0.0), tot.get("G", 0.0)], device=dev) / args.steps
        tmin = t.clone()
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(tmin, op=torch.distributed.ReduceOp.MIN)
        # a collective ends when the LAST rank has joined: the rank that arrives last waits only for the transfer (min
        # over ranks = communication the step could not hide), every other rank also waits for it (max - min = rank skew:
        # GPUs on their power caps do not run the ~20 ms between two sync points at the same speed)
        comm = {"wait_ms_per_step_min_over_ranks": {"D_allreduce": tmin[0].item(), "G_allreduce": tmin[1].item()},
                "wait_ms_per_step_max_over_ranks": {"D_allreduce": t[0].item(), "G_allreduce": t[1].item()},
                "g_buckets": [(b - a) * 4 for a, b, _ in getattr(ts, "_g_buckets", [])],
                "d_arena_bytes": ts.DA.store.grad_arena.numel() * 4,
                "note": "CUDA events on the compute stream around work.wait(); min over ranks = exposed communication, "
                        "max - min = skew between ranks"}
        ts.comm_profile = None
    losses = ts.loss_dict() if train else {}
    # Per-kernel-family rooflines: the same K steps again with a CUDA-event pair around every implicit-GEMM and
    # InstanceNorm-tail launch. The product path overlaps weight gradients (side stream) with the bandwidth-bound
    # passes, which would fold contention and double counting into per-launch durations, so this pass runs the
    # launches serialised on one stream; `value` above is the uninstrumented, overlapped path.
    side = getattr(ts.G, "wgrad_stream", None) if train else None
    if side is not None:
        ts.G.wgrad_stream = None
    _C.TIMING["records"].clear()
    _C.TIMING["on"] = True
    ms_roof = timed(step_dev, args.steps)
    _C.TIMING["on"] = False
    if side is not None:
        ts.G.wgrad_stream = side
    agg = {}
    for kind, work, a, b, _tag in _C.TIMING["records"]:
        fam = "tail" if kind.startswith("tail:") else kind
        t, f, c = agg.get(fam, (0.0, 0.0, 0))
        agg[fam] = (t + a.elapsed_time(b), f + work, c + 1)
    _C.TIMING["records"].clear()
    # end-to-end through the public call with host buffers
    if ms_e2e is None:
        for i in range(min(2, args.warmup)):
            step_host(i)
        ms_e2e = timed(step_host, args.steps)
    sampler.stop()
    if args.layers and rank == 0:
        # serialised launches (no weight-gradient side stream), so a kernel's duration is its own
        if side is not None:
            ts.G.wgrad_stream = None
        layer_table(lambda: step_dev(0), args.layers)
        if side is not None:
            ts.G.wgrad_stream = side

    if rank == 0:
        hbm, tf_burst, tf_sus, which = peaks()
        imgs = B * world * args.steps
        gfl = gflop_per_image(args.gen, args.workload, args.version, True, S)
        if train:
            what = (f"{args.gen} nf=64 + PatchDiscriminator, version-{args.version} loss stack (LSGAN + L1 + "
                    f"{'pan' if args.version == 2 else 'VGG16 perceptual'} + GP every step)")
        else:
            what = f"{args.gen} nf=64 generator forward (test.py:202-203)"
        headline = train and args.gen == "UNet++" and args.version == 2
        line = {"metric": "train images/sec (G+D step) UNet++ 256^2" if headline else
                f"{'train' if train else 'inference'} images/sec {args.gen} {S}^2 (v{args.version})",
                "value": imgs / (ms / 1e3), "unit": "images/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"{what}, batch {B}/GPU, {S}x{S}, random-init weights",
                           "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "per-step working set (>10 GB of activations) exceeds the 126 MB L2; no explicit flush",
                           "e2e_path": ("TrainStep.step_from_host: pinned batch -> H2D (next batch's copy in flight) -> step "
                                        "-> D2H of the five loss scalars, read with a one-step lag, the last timed step "
                                        "synchronously" if train else
                                        "pinned batch -> H2D -> graph-replayed forward -> D2H of the generated images")},
                "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": sampler.summary(), "losses_last_step": losses, "comm": comm,
                "step_tflops": gfl * B * args.steps / (ms / 1e3) / 1e3, "gflop_per_image": gfl}
        if "conv" in agg:
            t, f, c = agg["conv"]
            ach = f / (t / 1e3) / 1e12
            traffic, top = ncu_traffic()
            line["roofline"] = {"bound": "tensor", "kernel": "igemm_rows_kernel / igemm_halo_kernel / igemm_conv_kernel (conv forward + "
                                "input gradient, all shapes)", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s",
                                "frac": ach / tf_sus, "peak_source": f"bf16_tflops_sustained, {which}",
                                "traffic": traffic, "traffic_launch": top, "launches": c,
                                "frac_of_nominal_dense_peak": ach / 2250.0,
                                "ncu_tensor_pipe_active_pct_all_conv_gemms": ncu_tensor_pipe() if headline else None,
                                "share_of_step": t / ms_roof,
                                "measured": f"{args.steps} further steps, CUDA events around every launch, launches "
                                            f"serialised on one stream ({ms_roof / args.steps:.2f} ms/step)"}
        if "wgrad" in agg:
            t, f, c = agg["wgrad"]
            ach = f / (t / 1e3) / 1e12
            line["roofline_wgrad"] = {"bound": "tensor", "kernel": "wgrad_taps_kernel / wgrad_kernel", "achieved": ach,
                                      "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus, "launches": c,
                                      "share_of_step": t / ms_roof}
        if "tail" in agg:
            t, f, c = agg["tail"]
            ach = f / (t / 1e3) / 1e9
            line["roofline_tail"] = {"bound": "hbm", "kernel": "in_act_fwd / in_bwd_reduce / in_bwd_apply_re "
                                     "(InstanceNorm + activation forward / backward)", "achieved": ach, "peak": hbm,
                                     "unit": "GB/s", "frac": ach / hbm, "peak_source": f"hbm_gbs, {which}",
                                     "launches": c, "share_of_step": t / ms_roof}
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline(4, 1, 2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and train and not args.no_cudnn:
            # the reference's own GPU path on this box (eager + cuDNN): run after every timing of ours, our engines freed
            ts = step_dev = step_host = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            line["cudnn_baseline"] = cudnn_baseline(dev, B, S, gen=args.gen)
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
    ap.add_argument("--gen", default="UNet++", choices=["UNet++", "UNet", "BCDUNet"])
    ap.add_argument("--version", type=int, default=2, choices=[1, 2])
    ap.add_argument("--workload", default="train", choices=["train", "infer"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cudnn", action="store_true", help="skip the reference-on-cuDNN leg (same GPU, eager PyTorch)")
    ap.add_argument("--layers", default="", help="diagnostic: write a per-shape / per-kernel timing table (JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
