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


def check_equal_shards(n_local, what='calibration tensor'):
    """Every rank must hold the same number of samples: the order statistics (n_glob = n_local * ranks), the 2^24 chunk
    rule and all_gather_cat assume it, and the GPU-count-invariant partial sums need shards of whole 32-token slabs."""
    if not active():
        return
    t = torch.tensor([n_local, -n_local], dtype=torch.int64)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = t.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if int(t[0]) != -int(t[1]):
        raise ValueError(f'{what}: ranks hold different numbers of samples ({-int(t[1])}..{int(t[0])}); '
                         f'shard the calibration set evenly (calib_size % world_size == 0)')


def all_reduce_sum(t):
    """In-place SUM all-reduce of an error-sum tensor (identity when not distributed)."""
    if active():
        t = t.contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_max(t):
    """In-place MAX all-reduce (identity when not distributed)."""
    if active():
        t = t.contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def all_gather_cat(t, dim=0):
    """Concatenate equally-shaped shards along `dim` in rank order (identity when not distributed)."""
    if not active():
        return t
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return torch.cat(parts, dim=dim)


# ---------------------------------------------------------------------------------------------- order statistics
def _float_keys(x):
    """float32 -> int64 keys whose integer order is the float order (negatives reversed below the positives)"""
    u = x.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return torch.where(u >= 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)


def _keys_to_float(k):
    u = torch.where(k >= 0x80000000, k & 0x7FFFFFFF, (~k) & 0xFFFFFFFF)
    u = torch.where(u >= 0x80000000, u - (1 << 32), u)               # as signed 32-bit
    return u.to(torch.int32).view(torch.float32)


def kth_values(local_sorted, k, seg=None):
    """Exact global order statistics without gathering the data.

    local_sorted [rows, n_local]: this rank's shard, sorted ascending along the last dim (float32).
    k [T] or [rows, T] int64: 0-based global ranks.  Returns [rows, T]: the k-th smallest element of the union of all
    ranks' shards -- bit for bit what sort(all_gather(x)) would hold at index k -- by bisection on the integer image
    of the float order, 64 pivots per round: count(<= pivot) is a searchsorted on every shard plus one all-reduce of
    [rows, T, 64] integers, 6 rounds.  (Every rank sorts only its own 1/R of the data; gathering and sorting everything on every rank made the
    seeding cost grow with the number of GPUs.)
    seg = (my_segment, n_segments): ranks are grouped into segments and the statistics are taken per segment (the
    reference's 2^24 chunk rule when one chunk spans several ranks); returns [n_segments, rows, T]."""
    rows, n_local = local_sorted.shape
    k = k.to(local_sorted.device)
    if k.dim() == 1:
        k = k.view(1, -1).expand(rows, -1)
    if not active():
        return torch.gather(local_sorted, 1, k)
    my_seg, n_seg = seg if seg is not None else (0, 1)
    T = k.shape[1]
    dev = local_sorted.device
    lo_loc = _float_keys(local_sorted[:, :1]).expand(rows, T)
    hi_loc = _float_keys(local_sorted[:, -1:]).expand(rows, T)
    lo = torch.full((n_seg, rows, T), 1 << 40, dtype=torch.int64, device=dev)
    hi = torch.full((n_seg, rows, T), -1, dtype=torch.int64, device=dev)
    lo[my_seg], hi[my_seg] = lo_loc, hi_loc
    both = torch.stack([-lo, hi])                             # one MAX all-reduce for (min, max)
    dist.all_reduce(both, op=dist.ReduceOp.MAX)
    lo, hi = -both[0], both[1]
    need = (k + 1).unsqueeze(0).unsqueeze(-1)
    W = 64                                                    # pivots per round: 2^32 keys shrink by 64x per round
    frac = torch.arange(1, W + 1, device=dev, dtype=torch.int64)
    for _ in range(6):
        span = hi - lo
        piv = lo.unsqueeze(-1) + (span.unsqueeze(-1) * frac) // W          # [n_seg, rows, T, W], last pivot == hi
        v = _keys_to_float(piv[my_seg].contiguous()).reshape(rows, T * W)
        cnt = torch.zeros((n_seg, rows, T, W), dtype=torch.int64, device=dev)
        cnt[my_seg] = torch.searchsorted(local_sorted, v.contiguous(), right=True).view(rows, T, W)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        first = (cnt >= need).to(torch.int64).argmax(dim=-1, keepdim=True)  # first pivot with count(<= pivot) >= k+1
        new_hi = torch.gather(piv, -1, first).squeeze(-1)
        prev = torch.gather(piv, -1, (first - 1).clamp(min=0)).squeeze(-1) + 1
        lo = torch.where(first.squeeze(-1) > 0, prev, lo)
        hi = new_hi
    out = _keys_to_float(lo.contiguous()) + 0.0          # (a selected zero is returned as +0.0)
    return out if seg is not None else out[0]
