"""Data-parallel plumbing for the -O train step: one process per GPU, views sharded across ranks,
ONE all-reduce per step over a flat fp32 gradient bucket (hash-grid table + both MLPs = 1.8 M
floats, 7.3 MB at the cfg3 sizes), NCCL over NVLink/NVSwitch on GPUs, gloo on CPU for the tests.

The reference has no working multi-GPU path (a dormant DDP wrap, nerf/utils.py:200-202); DDP would
all-reduce exactly these parameters.  Parameter .grad tensors are views into the bucket, so autograd
accumulates straight into it and no flatten / unflatten copies are needed.
"""
import torch
import torch.distributed as dist


class FlatGradBucket:
    def __init__(self, params, device=None, extra=0):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        device = device if device is not None else self.params[0].device
        # `extra` trailing floats ride along in the same collective (e.g. an inf/nan flag, a sample count)
        self.flat = torch.zeros(n + extra, dtype=torch.float32, device=device)
        self.extra = self.flat[n:]
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.numel = n

    def zero(self):
        self.flat.zero_()

    def attach(self):
        """Re-point .grad at the bucket (needed after optimizer.zero_grad(set_to_none=True))."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce(self, average=True, async_op=False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if average and not async_op:
            self.flat[:self.numel].div_(dist.get_world_size())
        return work


def shard_views(n_views, rank, world_size):
    """Contiguous, balanced split of `n_views` camera views; returns (first, count) for `rank`."""
    base, rem = divmod(n_views, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count
