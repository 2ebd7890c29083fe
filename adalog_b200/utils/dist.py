"""Sample-sharded data parallelism for the calibration sweep.

Every rank scores all candidates on its slice of the calibration samples; the per-candidate FP64
error sums are all-reduced (NCCL over NVLink on the GPU box, gloo in the CPU tests) and every rank
then takes the identical top-k.  Order statistics used for candidate seeding are taken over the
all-gathered tensor so that they equal the single-process result.
"""
import torch
import torch.distributed as dist


def active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if active() else 1


def rank():
    return dist.get_rank() if active() else 0


def all_reduce_sum(t):
    """In-place SUM all-reduce of an error-sum tensor (identity when not distributed)."""
    if active():
        t = t.contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_gather_cat(t, dim=0):
    """Concatenate equally-shaped shards along `dim` in rank order (identity when not distributed)."""
    if not active():
        return t
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=dim)
