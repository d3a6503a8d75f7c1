"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): one rank per GPU over NCCL, each on its
shard of the global batch with its slice of the global smoothed-label / GP-alpha draws. The bucketed, overlapped
allreduce of TrainStep must reproduce the single-process global-batch gradients (InstanceNorm and the GP norm are
per sample, so only the fp32 summation order differs)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle as orc
    from tactile_gan_b200.discriminators.discriminators import create_disc
    from tactile_gan_b200.generators.generators import create_gen
    from tactile_gan_b200.step import TrainStep
    from tactile_gan_b200.util import init_weights
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TG_P2P="1")   # opt-in peer-memory all-reduce
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    solo = [dist.new_group([r]) for r in range(world)]
    per, size, nf = 2, 64, 16
    gb = per * world

    def nets():
        torch.manual_seed(5)
        g_, d_ = create_gen("UNet++", 3, 3, nf, True), create_disc("patch", 3, 3, nf, True, True)
        init_weights(g_)
        init_weights(d_)
        return g_.to(dev), d_.to(dev)

    g = torch.Generator().manual_seed(13)
    a, b = orc.synthetic_batch(g, gb, size)
    alpha = torch.rand(gb, 1, generator=g)
    sl = slice(rank * per, (rank + 1) * per)
    netG, netD = nets()
    ts = TrainStep(netG, netD, per, size, size, lr=0.0)
    label = orc.make_real_label((gb, 1, ts.h5, ts.w5), True, generator=g)   # one global draw, sliced per rank
    ts.BUCKET_BYTES = 256 << 10        # several buckets even for this small generator
    ts.set_label(label[sl])
    ts.step(a[sl].to(dev), b[sl].to(dev), regularize=True, alpha=alpha[sl])
    torch.cuda.synchronize()
    nb = len(ts._g_buckets)
    # the one-shot peer-memory all-reduce (p2p.PeerReducer / tg_allreduce_oneshot) against NCCL on the same data: both
    # staging halves, a full-size and a shorter buffer, bit-identical results on the two ranks
    peer_ok = ts.peer is not None
    peer_err, peer_same = 0.0, True
    if peer_ok:
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        for k, n in enumerate((ts.peer.numel, 4096, ts.peer.numel, 1 << 18)):
            x = torch.randn(n, device=dev, generator=gen)
            ref_sum = x.clone()
            dist.all_reduce(ref_sum)
            got = ts.peer.allreduce_(x.clone())
            torch.cuda.synchronize()
            peer_err = max(peer_err, float((got - ref_sum).abs().max() / ref_sum.abs().max()))
            other = got.clone()
            dist.broadcast(other, 0)
            peer_same = peer_same and bool(torch.equal(other, got))
    got_g = {k: v / world for k, v in ts.G.store.grads_by_name().items()}
    got_d = {k: v / world for k, v in ts.DA.store.grads_by_name().items()}
    if rank == 0:
        netG1, netD1 = nets()
        ref = TrainStep(netG1, netD1, gb, size, size, lr=0.0, process_group=solo[0])
        ref.set_label(label)
        ref.step(a.to(dev), b.to(dev), regularize=True, alpha=alpha)
        torch.cuda.synchronize()
        rg, rd = ref.G.store.grads_by_name(), ref.DA.store.grads_by_name()
        num = sum(float((got_g[k] - rg[k]).norm() ** 2) for k in rg)
        den = sum(float(rg[k].norm() ** 2) for k in rg)
        numd = sum(float((got_d[k] - rd[k]).norm() ** 2) for k in rd)
        dend = sum(float(rd[k].norm() ** 2) for k in rd)
        out.put((nb, (num / den) ** 0.5, (numd / dend) ** 0.5, peer_ok, peer_err, peer_same))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_bucketed_allreduce_matches_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    buckets, err_g, err_d, peer_ok, peer_err, peer_same = out.get()
    print(f"\nbuckets {buckets}, G {err_g:.4f}, D {err_d:.4f}, peer-memory all-reduce used: {peer_ok}, "
          f"vs NCCL {peer_err:.2e}, bit-identical on both ranks: {peer_same}")
    assert buckets >= 2
    assert peer_ok, "symmetric-memory peer all-reduce could not be set up on this 2-GPU node"
    assert peer_err < 1e-6 and peer_same
    # bf16 storage + atomics: two launches of the same step differ at this level too
    assert err_g < 5e-2 and err_d < 2e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_train_cli_matches_single_process_global_batch(tmp_path):
    """`python -m torch.distributed.run --nproc-per-node 2 -m tactile_gan_b200.train` (BASELINE configs[2] in small):
    NCCL init, sharded sampler, rank-0-only files, identical replicas (train.main raises if they diverge), and the
    weights of one epoch = one global batch agree with the single-process run on that global batch."""
    import subprocess
    common = ["--synthetic", "4", "--image_size", "64", "--nf", "8", "--total_epochs", "1", "--epoch_constant", "1",
              "--version", "2", "--threads", "0", "--no_label_smoothing", "--lambda_gp", "0"]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for name, launch, bs in (("dp", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                     "--master-addr", "127.0.0.1", "--master-port", str(35000 + os.getpid() % 2000),
                                     "-m", "tactile_gan_b200.train"], "2"),
                             ("one", [sys.executable, "-m", "tactile_gan_b200.train"], "4")):
        (tmp_path / name / "data").mkdir(parents=True)
        r = subprocess.run(launch + ["--data", str(tmp_path / name / "data"), "--batch_size", bs, "--folder_save", "run"]
                           + common, cwd=tmp_path / name, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        assert r.stdout.count("==training epoch") == 1                     # only rank 0 logs
    dp = torch.load(tmp_path / "dp" / "models" / "run" / "final_model.pth", weights_only=False)
    one = torch.load(tmp_path / "one" / "models" / "run" / "final_model.pth", weights_only=False)
    assert float(dp["optimizerG_state_dict"]["state"][0]["step"]) == 1.0
    for key in ("gen", "disc"):
        d = torch.cat([(dp[key][k] - one[key][k]).flatten().cpu() for k in dp[key] if not k.startswith("clstm")])
        # one Adam step of lr 1e-3 from identical weights: sign flips of near-zero gradients cost 2*lr, the bulk agrees
        assert (d.abs() > 0.5e-3).float().mean().item() < 0.03, key
        assert d.abs().mean().item() < 0.05e-3, key
