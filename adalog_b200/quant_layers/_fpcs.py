"""Fast Progressive Combining Search driver shared by every quant layer.

One routine replaces the four near-identical loops of the reference (linear.py:483-523, :941-967,
matmul.py:243-262, conv.py:292-311).  The candidate arithmetic (linspace, -0.5, * delta, / (cnt-0.5),
gather, repeat_interleave) is kept op-for-op in FP32 so that candidate bits equal the reference's; the
scoring itself is delegated to `score(scales, aux, topk) -> indices`, i.e. to the CUDA sweeps.
"""
import torch

from .. import _const
from ..utils import dist as adist


SELECT_MIN_N = 1 << 16      # rows at least this long are selected by radix passes instead of sorted


def candidate_chunks(P, tile=128):
    """slices of at most `tile` candidates (the kernels score one TMEM-lane tile of 128 per pass)"""
    return [(p0, min(P, p0 + tile)) for p0 in range(0, P, tile)]


def refine(scales, aux, delta, score, axis, new_cnt, width, steps, floor=None):
    """Refinement steps 1..steps of FPCS starting from an already built first grid.

    scales/aux: candidate tensors with the candidate axis at `axis` (0 or -1); aux carries zero points
    (or log bases, linear.py:961); delta: grid pitch along the scale axis.
    """
    dev = scales.device
    idx = score(scales, aux, width)
    best_s = torch.gather(scales, dim=axis, index=idx)
    best_a = torch.gather(aux, dim=axis, index=idx)
    left = steps - 1
    while left > 0:
        # linspace(0, 1, new_cnt) - 0.5 as the reference computes it (FP32, same op), cached per (count, device);
        # repeat_interleave(n) as expand + reshape: same values, one copy kernel instead of five launches
        ramp_c = _const.linspace01_centered(new_cnt, dev)
        if axis == 0:
            offs = ramp_c.view(-1, *([1] * (scales.dim() - 1))) * delta
            delta = delta / (new_cnt - 0.5)
            scales = (best_s.unsqueeze(1) + offs.unsqueeze(0)).reshape(-1, *scales.shape[1:])
            aux = best_a.unsqueeze(1).expand(-1, new_cnt, *best_a.shape[1:]).reshape(-1, *best_a.shape[1:])
        else:
            offs = ramp_c[None, :] * delta
            delta = delta / (new_cnt - 0.5)
            scales = (best_s.unsqueeze(-1) + offs.unsqueeze(-2)).reshape(*scales.shape[:-1], -1)
            if floor is not None:
                scales = scales.clamp(min=floor)
            aux = best_a.unsqueeze(-1).expand(*best_a.shape, new_cnt).reshape(*best_a.shape[:-1], -1)
        idx = score(scales, aux, 1 if left == 1 else width)
        if left > 1:
            best_s = torch.gather(scales, dim=axis, index=idx)
            best_a = torch.gather(aux, dim=axis, index=idx)
        left -= 1


def search(scales, aux, score, axis, eq_n, width=16, steps=6, floor=None):
    """Full FPCS: first grid -> top `width` -> (steps-1) refinements of eq_n/width points each."""
    delta = (scales[1:2] - scales[0:1]) if axis == 0 else (scales[:, 1:2] - scales[:, 0:1])
    refine(scales, aux, delta, score, axis, int(eq_n / width), width, steps, floor)


def percentile_grid(delta_min, delta_max, n_levels, num_zp, num_scale, axis, lead_dims=0):
    """scale x zero-point grid with index p = zp_idx * num_scale + scale_idx
    (linear.py:442-451, :472-481; matmul.py:231-240; conv.py:281-290)."""
    dev = delta_min.device
    ramp = _const.linspace01(num_scale, dev)
    zp_lo = int(n_levels - num_zp / 2)
    zp_hi = int(n_levels + num_zp / 2)
    zps = _const.int_range(zp_lo, zp_hi, dev).repeat_interleave(num_scale)
    if axis == 0:
        ones = [1] * lead_dims
        scales = (delta_min + ramp.view(-1, *ones) * (delta_max - delta_min)).repeat(num_zp, *ones) / (2 * n_levels - 1)
        zps = zps.view(-1, *ones).repeat(1, *scales.shape[1:])
    else:
        scales = (delta_min + ramp[None, :] * (delta_max - delta_min)).repeat(1, num_zp) / (2 * n_levels - 1)
        zps = zps[None, :].repeat(scales.shape[0], 1)
    return scales, zps


def chunked_quantile(x2, pct):
    """quantile over the last dim of x2 = x.view(g, 1, -1), doubling the middle dim until the reduced dim fits
    torch.quantile's 2^24 limit, then averaging the chunk quantiles (linear.py:465-471, matmul.py:223-230).

    x2 is this rank's shard (all of x without data parallelism); the statistics are those of the concatenation of the
    shards in rank order.  A chunk is either a whole number of ranks (exact distributed selection per group of ranks)
    or a fraction of one rank's shard (local quantiles, the per-chunk values exchanged before the mean)."""
    g = x2.shape[0]
    R = adist.world_size()
    n_loc = x2.numel() // g
    mbs = 1
    while (n_loc * R) // mbs > (1 << 24):
        mbs *= 2
    if R == 1 or mbs >= R:
        if mbs % R:
            raise NotImplementedError(f'2^24 chunk rule with {mbs} chunks over {R} ranks')
        up, lo = quantile_pair(x2.reshape(g, mbs // R, -1), pct, -1, local=True)      # [2, g, mbs/R]
        up, lo = adist.all_gather_cat(up.contiguous(), dim=-1), adist.all_gather_cat(lo.contiguous(), dim=-1)
    else:
        if R % mbs:
            raise NotImplementedError(f'2^24 chunk rule with {mbs} chunks over {R} ranks')
        up, lo = quantile_pair(x2.reshape(g, 1, -1), pct, -1, seg=(adist.rank() // (R // mbs), mbs))   # [2, g, mbs]
    return up.mean(dim=-1), lo.mean(dim=-1)        # each [2, g]


def quantile_pair(x, pct, dim, local=False, seg=None):
    """(quantile(x, pct), quantile(x, 1 - pct)) along dim with ONE sort: torch.quantile sorts once per call and
    interpolates every requested q independently, so the values are bit-identical to the reference's two calls
    (linear.py:441-442, :459-471; matmul.py:223-230; conv.py:280-281).

    Under data parallelism (and unless local=True) `dim` is the sharded dimension: every rank sorts its own shard and
    the two neighbouring order statistics of each q are selected exactly across ranks (utils/dist.py kth_values), then
    interpolated with torch.quantile's own arithmetic (ranks = q*(n-1) in FP32, lerp): same bits as torch.quantile on
    the all-gathered tensor.  seg=(my_segment, n_segments): statistics per group of ranks, returned with a trailing
    segment axis [nq, ..., n_segments]."""
    n = pct.numel()
    q = (_const.pct_pair(float(pct[0]), float(pct[1]), x.device) if (n == 2 and not pct.is_cuda)
         else torch.cat([pct, 1 - pct]).to(x.device))
    dist_on = adist.active() and not local
    # Exact radix selection (csrc/select_kernels.cu) instead of a full sort wherever the reduced dimension is long:
    # four passes over the data, and under data parallelism four all-reduces of the digit histograms instead of
    # sort + bisection.  Short rows (weights, per-channel statistics) stay on torch.quantile.
    xs = x.movedim(dim, -1)
    use_select = (x.is_cuda and x.dtype == torch.float32 and xs.shape[-1] >= SELECT_MIN_N and q.numel() * 2 <= 8
                  and seg is None)
    if not dist_on and not use_select:
        both = torch.quantile(x, q, dim=dim)
        return both[:n], both[n:]
    lead = xs.shape[:-1]
    rows2d = xs.reshape(-1, xs.shape[-1]).contiguous()
    ranks_per_seg = (adist.world_size() // (seg[1] if seg is not None else 1)) if dist_on else 1
    n_glob = rows2d.shape[1] * ranks_per_seg
    ranks = q.to(rows2d.dtype) * (n_glob - 1)                # torch.quantile: q * last_index in the input dtype
    below = ranks.to(torch.int64)
    weights = ranks - below
    above = ranks.ceil().to(torch.int64)
    nq = q.numel()
    if use_select:
        from .. import ops
        vals = ops.select_kth(rows2d, torch.cat([below, above]), reduce_hist=adist.all_reduce_sum if dist_on else None)
        vals = vals.t()                                                 # [2nq, rows]
        res = vals[:nq].clone().lerp_(vals[nq:], weights.view(-1, 1)).reshape(nq, *lead)
        return res[:n], res[n:]
    srt, _ = rows2d.sort(dim=-1)
    vals = adist.kth_values(srt, torch.cat([below, above]), seg)       # [rows, 2nq] or [n_seg, rows, 2nq]
    if seg is not None:
        vals = vals.permute(2, 1, 0)                                    # [2nq, rows, n_seg]
        res = vals[:nq].clone().lerp_(vals[nq:], weights.view(-1, 1, 1))
        res = res.reshape(nq, *lead[:-1], seg[1]) if lead and lead[-1] == 1 else res.reshape(nq, *lead, seg[1])
    else:
        vals = vals.t()                                                 # [2nq, rows]
        res = vals[:nq].clone().lerp_(vals[nq:], weights.view(-1, 1)).reshape(nq, *lead)
    return res[:n], res[n:]
