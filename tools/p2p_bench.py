"""One-shot peer-memory all-reduce vs NCCL on the step's small gradient buffers (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tactile_gan_b200 import p2p  # noqa: E402

if __name__ == "__main__":
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    for ctas in (32, 64, 128, 256):
        p2p.CTAS = ctas
        red = p2p.PeerReducer(2 << 20, dev)
        for n in (1587072, 1 << 18, 1 << 16):
            x = torch.randn(n, device=dev)
            res = {}
            for name, fn in (("peer", lambda: red.allreduce_(x)), ("nccl", lambda: dist.all_reduce(x))):
                for _ in range(5):
                    fn()
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 50], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res[name] = t.item()
                x.normal_()
            if rank == 0:
                print(f"world {world} ctas {ctas:3d}  {n * 4 / 1e6:6.2f} MB: peer {res['peer']:7.1f} us   nccl {res['nccl']:7.1f} us", flush=True)
        del red
    dist.destroy_process_group()
