"""Multi-rank host logic on CPU (gloo, world_size 2): sharding the batch across ranks, summing the
gradients with all_reduce and scaling by 1/world reproduces the single-process gradient of the mean
losses -- provided each rank takes ITS SLICE of the global smoothed-label tensor and GP alpha
(SURVEY 8e). Uses the oracle as the per-rank step (test infrastructure)."""
import os
import sys
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grads(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    torch.set_num_threads(2)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "unetpp_ls.pt"), weights_only=False)
    m = fx["meta"]
    g = torch.Generator().manual_seed(7)
    gb = 2 * world
    a, b = orc.synthetic_batch(g, gb, m["size"])
    alpha = torch.rand(gb, 1, generator=g)
    label = orc.make_real_label((gb, 1, 1, 1), True, generator=g)
    sl = slice(rank * 2, rank * 2 + 2)
    sd_g = OrderedDict((k, v.clone()) for k, v in fx["init_G"].items())
    sd_d = OrderedDict((k, v.clone()) for k, v in fx["init_D"].items())
    cfg = orc.StepConfig(lr=0.0)   # lr 0: only the gradients matter
    res = orc.train_step(sd_g, sd_d, {}, {}, a[sl], b[sl], label[sl], alpha[sl], cfg)
    flat = torch.cat([res["grads_D"][k].flatten() for k in sorted(res["grads_D"])])
    dist.all_reduce(flat)            # what TrainStep._allreduce does to the flat gradient arena
    flat /= world                    # grad_scale = 1/world inside tg_adam_step
    if rank == 0:
        full = orc.train_step(OrderedDict((k, v.clone()) for k, v in fx["init_G"].items()),
                              OrderedDict((k, v.clone()) for k, v in fx["init_D"].items()), {}, {}, a, b, label,
                              alpha, cfg)
        ref = torch.cat([full["grads_D"][k].flatten() for k in sorted(full["grads_D"])])
        out.put(((flat - ref).norm() / ref.norm()).item())
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_grads, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert out.get() < 1e-4


def test_bucket_plan_covers_the_arena_in_completion_order():
    """layers.plan_buckets: contiguous, gap-free, front-to-back buckets whose 'ready' parameter is the last one
    of the bucket in backward-completion order."""
    sys.path.insert(0, ROOT)
    from tactile_gan_b200.layers import plan_buckets
    sizes = [12, 4, 40, 8, 100, 4, 4, 60]
    layout, off = [], 0
    for i, sz in zip([5, 3, 7, 0, 2, 1, 6, 4], sizes):      # (param index, offset, padded size)
        layout.append((i, off, sz))
        off += sz
    buckets = plan_buckets(layout, 50)
    assert buckets[0][0] == 0 and buckets[-1][1] == off
    for (a0, b0, _), (a1, _, _) in zip(buckets, buckets[1:]):
        assert b0 == a1
    assert [b[2] for b in buckets] == [7, 2, 4]              # ready after params 7, 2 and (tail) 4
    assert all(b - a >= 50 for a, b, _ in buckets[:-1])
    assert plan_buckets(layout, 10 ** 9) == [(0, off, 4)]


def _bucketed(rank, world, port, out):
    """What TrainStep._g_backward does on the comm stream, on CPU: all-reduce contiguous slices of the flat
    gradient arena asynchronously, bucket by bucket, and wait for all of them before the optimiser step."""
    sys.path.insert(0, ROOT)
    from tactile_gan_b200.layers import plan_buckets
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    arena = torch.randn(10_000, generator=g)
    whole = arena.clone()
    layout = [(i, i * 1000, 1000) for i in range(10)]
    works = [dist.all_reduce(arena[a:b], async_op=True) for a, b, _ in plan_buckets(layout, 2500)]
    for w in works:
        w.wait()
    dist.all_reduce(whole)
    if rank == 0:
        out.put(float((arena - whole).abs().max()))
    dist.destroy_process_group()


def test_bucketed_async_allreduce_equals_one_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucketed, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() == 0.0


def _train_host_logic(rank, world, port, out):
    """Product code of tactile_gan_b200.train's data-parallel host side (no CUDA needed): every rank draws the GLOBAL
    smoothed-label tensor from the same seed and keeps its rows; the DistributedSampler shards are disjoint and
    cover the set; the per-epoch loss reduction averages the rank means."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from tactile_gan_b200 import train as tg_train
    from torch.utils.data.distributed import DistributedSampler
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert tg_train.dp_env() == (rank, world, rank)
    per = 3
    lab = tg_train.global_real_label(world, per, 5, 5, generator=torch.Generator().manual_seed(21))
    mine = tg_train.rank_rows(lab, rank, per)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    ds = tg_train.SyntheticPairs(12, 8)
    sampler = DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=True, seed=21, drop_last=False)
    sampler.set_epoch(3)
    idx = torch.tensor(list(sampler))
    allidx = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(allidx, idx)
    mean = torch.full((8,), float(rank + 1))
    dist.all_reduce(mean)
    mean /= world
    if rank == 0:
        single = tg_train.global_real_label(world, per, 5, 5, generator=torch.Generator().manual_seed(21))
        out.put((bool(torch.equal(torch.cat(parts), single)), sorted(torch.cat(allidx).tolist()), mean[0].item()))
    dist.destroy_process_group()


def test_train_cli_data_parallel_host_logic():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = 27000 + os.getpid() % 2000
    procs = [ctx.Process(target=_train_host_logic, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same_label, covered, mean = out.get()
    assert same_label
    assert covered == list(range(12))
    assert mean == 1.5
