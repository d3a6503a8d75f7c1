"""One-shot all-reduce over NVLink peer memory for the small gradient buffers of the data-parallel step
(tg_allreduce_oneshot, csrc/tg_tail.cu). The staging buffers and flag pads are torch symmetric memory: device memory
whose handles are exchanged between the ranks of one node at rendezvous, so every rank can load from every other rank's
copy. Anything that cannot be set up (no peer access, symmetric memory unavailable in this build, ranks on different
nodes) raises PeerUnavailable: the caller keeps the NCCL collective for that buffer and says so once."""
import torch

from . import _C
from ._C import LL, ptr

CTAS = 128


class PeerUnavailable(RuntimeError):
    pass


class PeerReducer:
    def __init__(self, numel, device, group=None):
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            raise PeerUnavailable("torch.distributed is not initialised")
        try:
            import torch.distributed._symmetric_memory as symm
        except Exception as e:  # pragma: no cover
            raise PeerUnavailable(f"symmetric memory is not available in this torch build: {e}") from e
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise PeerUnavailable("at most 16 ranks")
        self.numel = (int(numel) + 3) // 4 * 4
        try:
            name = getattr(group, "group_name", None)
            if name is not None and hasattr(symm, "enable_symm_mem_for_group"):
                try:
                    symm.enable_symm_mem_for_group(name)
                except Exception:
                    pass
            self.stage = symm.empty(2 * self.numel, dtype=torch.float32, device=device)
            self.flags = symm.empty(CTAS * self.world, dtype=torch.int32, device=device)
            self.stage.zero_()
            self.flags.zero_()
            torch.cuda.synchronize(device)
            hs = symm.rendezvous(self.stage, group)
            hf = symm.rendezvous(self.flags, group)
            self._handles = (hs, hf)                         # keep the mappings alive
            self.bufs = torch.tensor(list(hs.buffer_ptrs), dtype=torch.int64, device=device)
            self.pads = torch.tensor(list(hf.buffer_ptrs), dtype=torch.int64, device=device)
            dist.barrier(group)                              # every rank's flags are zero before the first call
        except PeerUnavailable:
            raise
        except Exception as e:
            raise PeerUnavailable(f"symmetric-memory rendezvous failed: {type(e).__name__}: {e}") from e
        self.epoch = 0

    def allreduce_(self, buf):
        """Sum `buf` (fp32, contiguous, numel <= the size given at construction, multiple of 4) over the ranks, in place,
        on the current stream. Every rank must make the same sequence of calls."""
        n = buf.numel()
        if buf.dtype != torch.float32 or not buf.is_contiguous() or n > self.numel or n % 4:
            raise ValueError("PeerReducer.allreduce_: fp32 contiguous buffer of a multiple of 4 elements expected")
        self.epoch += 1
        # the staging halves are addressed with the size given at construction (identical on every rank)
        _C.call("allreduce_oneshot", ptr(buf), ptr(self.bufs), ptr(self.pads), self.rank, self.world, LL(n),
                LL(self.numel), self.epoch, CTAS)
        return buf
