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
