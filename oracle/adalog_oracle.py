"""CPU restatement of AdaLog's FPCS calibration sweep and fake-quant forward.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this file, and only as the checker / the CPU arm that is timed
beside the product.  The product (adalog_b200/) never imports it and has no CPU fallback.

What it restates: the arithmetic of GoatWu/AdaLog's quantizers/ and quant_layers/ hot path
(file:line citations are into /root/reference/).  The reference is pure PyTorch, so the restatement
is written against the same torch primitives (division, round, clamp, log2, quantile, topk, F.linear,
@, F.conv2d) in the same association order; run on torch-CPU it reproduces the reference's CPU
results bit for bit (error vectors, top-k index lists incl. exact-tie order, final parameters), and
run on torch-CUDA it is the "reference on the GPU" comparator (same device-specific topk / log2).

Parity pin: tests/golden/*.pt were produced by oracle/make_golden.py from the UNMODIFIED reference
(imported from /root/reference under oracle/ref_shim.py) and tests/test_oracle_golden.py checks this
file against every recorded evaluation.  The reference itself ships no tests or golden vectors.

Structure (deliberately not the reference's class tree): stateless functions over plain tensors plus
one generic progressive-refinement driver, so the six reference module classes become six short
`search_*` recipes.
"""
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch
import torch.nn.functional as F

SHIFT_GELU = 0.16997124254703522  # linear.py:749 (|min GELU|)
R_BASE = 37.0                     # logarithm.py:71

# None: contractions run in FP32 exactly like the reference.  torch.float64: the quantisation decisions stay FP32
# (bit-identical operands) but F.linear / @ / F.conv2d and the error reduction (also of the self-error sweeps) run in
# FP64 -- the "infinitely precise reference" used by the tests to show which side of a 1e-5 disagreement is the FP32
# rounding noise.
GEMM_DTYPE = None


def _up(*ts):
    if GEMM_DTYPE is None:
        return ts
    return tuple(None if t is None else t.to(GEMM_DTYPE) for t in ts)


# ----------------------------------------------------------------------------------------------
# trace recorder: every search evaluation reports (similarity tensor, k, dim, chosen indices)
# ----------------------------------------------------------------------------------------------
@dataclass
class Trace:
    evals: List[dict] = field(default_factory=list)

    def topk(self, sims, k, dim, tag=''):
        _, idx = torch.topk(sims, k=k, dim=dim)
        self.evals.append(dict(tag=tag, sims=sims.detach().clone(), k=k, dim=dim, idx=idx.clone()))
        return idx

    def argmax(self, sims, tag=''):
        """the one selection the reference makes with Tensor.argmax instead of topk (linear.py:691); its tie order
        differs from topk's, so it is recorded (and replayed) as what it is"""
        idx = sims.argmax(dim=0, keepdim=True)
        self.evals.append(dict(tag=tag, sims=sims.detach().clone(), k=1, dim=0, idx=idx.clone(), argmax=True))
        return idx


def _sim(raw, approx):
    """linear.py:87-88 / matmul.py:78-79 / conv.py:110-111 -- the only metric in the reference."""
    return -(raw - approx) ** 2


def _round_ste(z):
    """_ste.py:5-6 evaluated in eval mode: (round(z) - z) + z in FP32."""
    return (z.round() - z) + z


def chunked_eq_n(eq_n, memory, numel):
    """linear.py:118-121, matmul.py:102-106, conv.py:150-153 (memory = total_memory // 2)."""
    peq = int((memory / 4) // numel)
    return math.ceil(eq_n * 1.0 / math.ceil(eq_n * 1.0 / peq))


# ----------------------------------------------------------------------------------------------
# quantizer forwards (inference branch)
# ----------------------------------------------------------------------------------------------
def uniform_fakequant(x, scale, zero_point, n_levels, sym=False, return_codes=False):
    """quantizers/uniform.py:25-36."""
    x_int = torch.round(x / scale)
    if sym:
        code = x_int.clamp(-n_levels, n_levels - 1)
        deq = code * scale
    else:
        zr = _round_ste(zero_point)
        code = (x_int + zr).clamp(0, 2 * n_levels - 1)
        deq = (code - zr) * scale
    return (deq, code) if return_codes else deq


def twin_uniform_fakequant(x, scale2, n_levels):
    """quantizers/uniform.py:57-68 (scale2[0] positive range, scale2[1] negative range)."""
    pos = torch.round(x / scale2[0]).clamp(0, n_levels - 1).mul(scale2[0])
    neg = torch.round(x / scale2[1]).clamp(-n_levels, 0).mul(scale2[1])
    return (pos + neg).reshape_as(x)


def log2_fakequant(x, scale, n_levels, return_codes=False):
    """quantizers/logarithm.py:25-35."""
    v = (x / scale).clamp(min=1e-15, max=1.0)
    code = torch.round(-v.log2())
    mask = code < 2 * n_levels
    code = torch.clamp(code, 0, 2 * n_levels - 1)
    deq = 2 ** (-1 * code) * scale
    deq = deq * mask
    return (deq, code) if return_codes else deq


def logsqrt2_fakequant(x, scale, n_levels, return_codes=False):
    """quantizers/logarithm.py:45-62 (inference branch)."""
    v = (x / scale).clamp(min=1e-15, max=1.0)
    code = torch.round(-v.log2() * 2)
    mask = code < 2 * n_levels
    code = torch.clamp(code, 0, 2 * n_levels - 1)
    odd = (code % 2) * (math.sqrt(2) - 1) + 1
    deq = 2 ** (-1 * torch.ceil(code / 2)) * odd * scale
    deq = deq * mask
    return (deq, code) if return_codes else deq


def adalog_tables(q, n_levels):
    """quantizers/logarithm.py:77-81: inference LUTs, built in Python float64 then stored FP32."""
    q = int(q)
    t1 = torch.zeros(2 * n_levels)
    t2 = torch.zeros(2 * n_levels)
    for i in range(2 * n_levels):
        val = round((2 ** (-((q * i) % R_BASE) / R_BASE)) * (4 * n_levels - 2)) / (4 * n_levels - 2)
        t1[i] = math.floor(i * q / R_BASE)
        t2[i] = val
    return t1, t2


def search_table(n_levels):
    """linear.py:750-752 / matmul.py:313-315: the 120-entry FP32 LUT used inside the searches."""
    table = torch.tensor([2 ** (-j / R_BASE) for j in range(120)])
    table_scale = 1. / (4 * n_levels - 2)
    return torch.round(table / table_scale) * table_scale


def adalog_fakequant(x, scale, q, n_levels, table1=None, table2=None, return_codes=False):
    """quantizers/logarithm.py:83-99 (inference branch). q: int64 tensor [1]."""
    if table1 is None:
        table1, table2 = adalog_tables(int(q.item()), n_levels)
    table1, table2 = table1.to(x.device), table2.to(x.device)
    v = (x / scale).clamp(min=1e-15, max=1.0)
    code = torch.round(-v.log2() * R_BASE / q)
    mask = code < 2 * n_levels
    code = torch.clamp(code, 0, 2 * n_levels - 1)
    deq = (2 ** (-table1[code.long()])) * table2[code.long()] * scale
    deq = deq * mask
    return (deq, code) if return_codes else deq


def shift_fakequant(base_fn, x, shift, bias_reparamed, *args, **kw):
    """quantizers/logarithm.py:105-135, uniform.py:42-50: Q(x + shift) [- shift]."""
    out = base_fn(x + shift, *args, **kw)
    return out if bias_reparamed else out - shift


# ----------------------------------------------------------------------------------------------
# small state holders
# ----------------------------------------------------------------------------------------------
@dataclass
class UQ:
    """asymmetric UniformQuantizer state."""
    n_bits: int
    scale: Optional[torch.Tensor] = None
    zero_point: Optional[torch.Tensor] = None

    @property
    def n_levels(self):
        return 2 ** (self.n_bits - 1)

    def __call__(self, x):
        if self.n_bits == 32:
            return x
        return uniform_fakequant(x, self.scale, self.zero_point, self.n_levels)


@dataclass
class LQ:
    """(Shift)AdaLogQuantizer state."""
    n_bits: int
    scale: Optional[torch.Tensor] = None
    q: torch.Tensor = field(default_factory=lambda: torch.tensor([int(R_BASE)]))
    shift: Optional[torch.Tensor] = None  # None -> plain AdaLogQuantizer
    bias_reparamed: bool = False
    table1: Optional[torch.Tensor] = None
    table2: Optional[torch.Tensor] = None

    @property
    def n_levels(self):
        return 2 ** (self.n_bits - 1)

    def update_table(self):
        self.table1, self.table2 = adalog_tables(int(self.q.item()), self.n_levels)

    def __call__(self, x):
        if self.table1 is None:
            self.update_table()
        if self.shift is None:
            return adalog_fakequant(x, self.scale, self.q, self.n_levels, self.table1, self.table2)
        return shift_fakequant(adalog_fakequant, x, self.shift, self.bias_reparamed, self.scale, self.q,
                               self.n_levels, self.table1, self.table2)


@dataclass
class FixedLogQ:
    """(Shift)Log2Quantizer / (Shift)LogSqrt2Quantizer state (logarithm.py:8-65, 105-124)."""
    kind: str
    n_bits: int
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    bias_reparamed: bool = False

    @property
    def n_levels(self):
        return 2 ** (self.n_bits - 1)

    def __call__(self, x):
        fn = log2_fakequant if self.kind == 'log2' else logsqrt2_fakequant
        if self.shift is None:
            return fn(x, self.scale, self.n_levels)
        return shift_fakequant(fn, x, self.shift, self.bias_reparamed, self.scale, self.n_levels)


@dataclass
class TQ:
    """TwinUniformQuantizer state (uniform.py:53-68): scale [2, 1]."""
    n_bits: int
    scale: Optional[torch.Tensor] = None

    @property
    def n_levels(self):
        return 2 ** (self.n_bits - 1)

    def __call__(self, x):
        return twin_uniform_fakequant(x, self.scale, self.n_levels)


def _wview(weight, n_V):
    out_f, in_f = weight.shape
    return weight.view(n_V, out_f // n_V, in_f)


def quant_weight(weight, wq: UQ, n_V):
    """linear.py:90-92."""
    return wq(_wview(weight, n_V)).view(weight.shape)


# ----------------------------------------------------------------------------------------------
# candidate seeding
# ----------------------------------------------------------------------------------------------
def _grid(delta_min, delta_max, n_levels, num_zp, num_scale, lead, dim0):
    """Shared tail of linear.py:442-451 / :472-481, matmul.py:231-240, conv.py:281-290."""
    dev = delta_min.device
    lin = torch.linspace(0, 1, steps=num_scale).to(dev)
    zp_min = int(n_levels - num_zp / 2)
    zp_max = int(n_levels + num_zp / 2)
    zps = torch.tensor(range(zp_min, zp_max)).to(dev).repeat_interleave(num_scale)
    if dim0:   # candidates along dim 0
        splits = lin.view(-1, *([1] * lead)) * (delta_max - delta_min)
        scales = (delta_min + splits).repeat(num_zp, *([1] * lead)) / (2 * n_levels - 1)
        zps = zps.view(-1, *([1] * lead)).repeat(1, *scales.shape[1:])
    else:      # candidates along the last dim
        splits = lin[None, :] * (delta_max - delta_min)
        scales = (delta_min + splits).repeat(1, num_zp) / (2 * n_levels - 1)
        zps = zps[None, :].repeat(scales.shape[0], 1)
    return scales, zps


def weight_candidates(weight, n_V, n_levels, eq_n, l=0.9, r=1.0):
    """linear.py:432-451 -> scale [eq_n,n_V,rows,1] f32, zp [eq_n,n_V,rows,1] int64."""
    num_zp = min(16, n_levels)
    num_scale = int(eq_n / num_zp)
    pct = torch.tensor([l, r])
    w3 = _wview(weight, n_V)
    up = torch.quantile(w3, pct.to(weight.device), dim=-1).unsqueeze(-1)
    lo = torch.quantile(w3, (1 - pct).to(weight.device), dim=-1).unsqueeze(-1)
    return _grid(up[0:1] - lo[0:1], up[1:] - lo[1:], n_levels, num_zp, num_scale, 3, True)


def conv_weight_candidates(weight, n_levels, eq_n, l=0.9, r=1.0):
    """conv.py:271-290 (num_zp = n_levels, not min(16, .))."""
    num_zp = n_levels
    num_scale = int(eq_n / num_zp)
    pct = torch.tensor([l, r])
    w2 = weight.view(weight.shape[0], -1)
    up = torch.quantile(w2, pct.to(weight.device), dim=-1).unsqueeze(-1)
    lo = torch.quantile(w2, (1 - pct).to(weight.device), dim=-1).unsqueeze(-1)
    return _grid(up[0:1] - lo[0:1], up[1:] - lo[1:], n_levels, num_zp, num_scale, 2, True)


def activation_candidates(x, n_levels, eq_n, channel_wise, l=0.9, r=1.0):
    """linear.py:453-481 -> scale [1|C, eq_n] (clamped at 1e-4), zp [1|C, eq_n] int64."""
    num_zp = min(16, n_levels * 2)
    num_scale = int(eq_n / num_zp)
    pct = torch.tensor([l, r])
    if channel_wise:
        up = torch.quantile(x.view(-1, x.shape[-1]), pct.to(x.device), dim=0).transpose(0, 1)
        lo = torch.quantile(x.view(-1, x.shape[-1]), (1 - pct).to(x.device), dim=0).transpose(0, 1)
    else:
        mbs = 1
        while True:
            try:
                up = torch.quantile(x.view(mbs, -1), pct.to(x.device), dim=-1).mean(dim=-1).unsqueeze(0)
                lo = torch.quantile(x.view(mbs, -1), (1 - pct).to(x.device), dim=-1).mean(dim=-1).unsqueeze(0)
                break
            except RuntimeError:
                mbs *= 2
    scales, zps = _grid(up[:, 0:1] - lo[:, 0:1], up[:, 1:] - lo[:, 1:], n_levels, num_zp, num_scale, 0, False)
    return scales.clamp(min=1e-4), zps


def matmul_candidates(x, n_levels_B, eq_n, head_channel_wise, l=0.9, r=1.0):
    """matmul.py:211-240 (num_zp uses B_quantizer.n_levels for BOTH operands, :212)."""
    num_zp = min(16, n_levels_B)
    num_scale = int(eq_n / num_zp)
    pct = torch.tensor([l, r])
    mbs = 1
    if head_channel_wise:
        x_ = x.transpose(0, 1).contiguous()
        x_ = x_.view(x_.shape[0], mbs, -1)
    else:
        x_ = x.view(1, mbs, -1)
    while True:
        try:
            up = torch.quantile(x_, pct.to(x_.device), dim=-1).mean(dim=-1, keepdim=False)
            lo = torch.quantile(x_, (1 - pct).to(x_.device), dim=-1).mean(dim=-1, keepdim=False)
            break
        except RuntimeError:
            mbs *= 2
            x_ = x_.view(x_.shape[0], mbs, -1) if head_channel_wise else x_.view(1, mbs, -1)
    dmin = (up[0] - lo[0]).view(1, 1, -1, 1, 1)
    dmax = (up[1] - lo[1]).view(1, 1, -1, 1, 1)
    return _grid(dmin, dmax, n_levels_B, num_zp, num_scale, 4, True)


def positive_percentile(t, q):
    """linear.py:763-798 for a flat tensor (dim=0): rank ceil(count*q)-1 among the positive entries."""
    pos = torch.where(t > 0, t, torch.tensor(float('nan')).to(t.device))
    srt, _ = pos.sort(dim=0)
    counts = (~torch.isnan(srt)).sum(dim=0, keepdim=True).float()
    qv = q.reshape(q.numel(), 1)
    ranks = ((counts * qv).ceil().long() - 1).clamp(min=0)
    res = torch.gather(srt.unsqueeze(0).expand(q.numel(), *srt.shape), 1, ranks).squeeze(1)
    res.masked_fill_(torch.isnan(res), 0)
    return res


def postgelu_candidates(x, shift_item, eq_n, l=0.9, r=1.0):
    """linear.py:800-814 -> (ud [1,2], scales [1,eq_n])."""
    cand = positive_percentile(x.reshape(-1), torch.tensor([l, r]).to(x.device)) + shift_item
    cand = cand.unsqueeze(0)
    lin = torch.tensor([i / (eq_n - 1) for i in range(eq_n)]).to(x.device).view(1, -1)
    return cand, cand[:, 0:1] + (cand[:, 1:] - cand[:, 0:1]) * lin


# ----------------------------------------------------------------------------------------------
# generic progressive refinement (FPCS)
# ----------------------------------------------------------------------------------------------
def fpcs(scales, aux, strategy: Callable, dim, eq_n, width=16, steps=6, clamp_min=None):
    """linear.py:483-523, matmul.py:243-262, conv.py:292-311.

    scales/aux: candidate tensors with the candidate axis at `dim` (0 or -1); aux = zero-points (or
    log bases for the post-GELU joint search, linear.py:956-967).  strategy(scales, aux, topk) -> idx.
    """
    new_cnt = int(eq_n / width)
    dev = scales.device
    if dim == 0:
        delta = scales[1:2] - scales[0:1]
    else:
        delta = scales[:, 1:2] - scales[:, 0:1]
    return _fpcs_tail(scales, aux, delta, strategy, dim, new_cnt, width, steps, clamp_min, dev)


def _fpcs_tail(scales, aux, delta, strategy, dim, new_cnt, width, steps, clamp_min, dev):
    idx = strategy(scales, aux, width)
    top_s = torch.gather(scales, dim=dim, index=idx)
    top_a = torch.gather(aux, dim=dim, index=idx)
    remain = steps - 1
    while remain > 0:
        lin = torch.linspace(0, 1, steps=new_cnt).to(dev)
        if dim == 0:
            offs = (lin.view(-1, *([1] * (scales.dim() - 1))) - 0.5) * delta
            delta = delta / (new_cnt - 0.5)
            scales = (top_s.unsqueeze(1) + offs.unsqueeze(0)).reshape(-1, *scales.shape[1:])
            aux = top_a.repeat_interleave(new_cnt, dim=0)
        else:
            offs = (lin[None, :] - 0.5) * delta
            delta = delta / (new_cnt - 0.5)
            scales = (top_s.unsqueeze(-1) + offs.unsqueeze(-2)).reshape(*scales.shape[:-1], -1)
            if clamp_min is not None:
                scales = scales.clamp(min=clamp_min)
            aux = top_a.repeat_interleave(new_cnt, dim=-1)
        idx = strategy(scales, aux, 1 if remain == 1 else width)
        if remain > 1:
            top_s = torch.gather(scales, dim=dim, index=idx)
            top_a = torch.gather(aux, dim=dim, index=idx)
        remain -= 1


# ----------------------------------------------------------------------------------------------
# Linear family
# ----------------------------------------------------------------------------------------------
class LinearSearch:
    """State + evaluations for the asymmetric linear searches (linear.py:238-545, 548-621, 724-1006)."""

    def __init__(self, weight, bias, raw_input, raw_out, w_bit, a_bit, n_V=1, eq_n=128, calib_batch_size=32,
                 search_round=3, steps=6, fpcs_on=True, memory=8 * 2 ** 30, trace: Optional[Trace] = None,
                 a_kind='uniform', a_channel_wise=False):
        self.weight, self.bias = weight, bias
        self.raw_input, self.raw_out = raw_input, raw_out
        self.n_V, self.eq_n, self.bs = n_V, eq_n, calib_batch_size
        self.rows = weight.shape[0] // n_V
        self.in_f = weight.shape[1]
        self.search_round, self.steps, self.fpcs_on = search_round, steps, fpcs_on
        self.memory = memory
        self.trace = trace or Trace()
        self.wq = UQ(w_bit)
        self.a_channel_wise = a_channel_wise
        if a_kind == 'uniform':
            self.aq = UQ(a_bit)
        elif a_kind == 'twin':
            self.aq = TQ(a_bit, scale=torch.zeros(2, 1).to(weight.device))
        else:
            self.aq = LQ(a_bit, shift=torch.tensor([SHIFT_GELU]).to(weight.device))
            self.aq.q = self.aq.q.to(weight.device)
            self.table = search_table(self.aq.n_levels)

    # -- linear.py:111-121
    def init_calib(self):
        self.calib_size = self.raw_input.shape[0]
        numel = 8 * self.raw_input[:self.bs].numel() + 16 * self.raw_out[:self.bs].numel()
        self.peq = chunked_eq_n(self.eq_n, self.memory, numel)

    def _batches(self):
        for b0 in range(0, self.calib_size, self.bs):
            yield b0, min(self.calib_size, b0 + self.bs)

    def _chunks(self):
        for p0 in range(0, self.eq_n, self.peq):
            yield p0, min(self.eq_n, p0 + self.peq)

    # -- linear.py:296-318
    def sims_w_self(self, cs, cz):
        L = 2 * self.wq.n_levels - 1
        raw = _wview(self.weight, self.n_V).unsqueeze(0)
        sims = []
        for p0, p1 in self._chunks():
            s, z = cs[p0:p1], cz[p0:p1]
            wq = ((raw / s).round_() + z).clamp(0, L)
            wd = (wq - z) * s
            sims.append(torch.mean(_sim(*_up(raw, wd)), dim=-1, keepdim=False))
        return torch.cat(sims, dim=0)

    def eval_w_self(self, cs, cz, topk=1):
        sims = self.sims_w_self(cs, cz)
        idx = self.trace.topk(sims, topk, 0, 'w_self').reshape(topk, self.n_V, -1, 1)
        if topk == 1:
            self.wq.scale = torch.gather(cs, 0, idx).squeeze(0)
            self.wq.zero_point = torch.gather(cz, 0, idx).squeeze(0).float()
        return idx.squeeze(0)

    # -- linear.py:320-353
    def sims_a_self(self, cs, cz):
        L = 2 * self.aq.n_levels - 1
        per_batch = []
        for b0, b1 in self._batches():
            x = self.raw_input[b0:b1]
            raw_x = x.unsqueeze(-1)
            parts = []
            for p0, p1 in self._chunks():
                s, z = cs[:, p0:p1], cz[:, p0:p1]
                xq = ((x.unsqueeze(-1) / s).round_() + z).clamp_(0, L)
                xd = (xq - z) * s
                sim = _sim(*_up(raw_x, xd))
                if sim.dim() > 3:
                    sim = torch.mean(sim, dim=list(range(1, sim.dim() - 2)))
                if not self.a_channel_wise:
                    sim = torch.mean(sim, dim=1, keepdim=True)
                parts.append(torch.sum(sim, dim=0, keepdim=True))
            per_batch.append(torch.cat(parts, dim=-1))
        return torch.cat(per_batch, dim=0).sum(dim=0, keepdim=False)

    def eval_a_self(self, cs, cz, topk=1):
        sims = self.sims_a_self(cs, cz)
        idx = self.trace.topk(sims, topk, -1, 'a_self')
        if topk == 1:
            self.aq.scale = torch.gather(cs, -1, idx).squeeze(-1)
            self.aq.zero_point = torch.gather(cz, -1, idx).squeeze(-1).float()
        return idx

    def _out_view(self, b0, b1):
        ro = self.raw_out[b0:b1].unsqueeze(-2)
        return ro.view(*ro.shape[:-1], self.n_V, -1)

    # -- linear.py:355-392
    def sims_w(self, cs, cz):
        L = 2 * self.wq.n_levels - 1
        per_batch = []
        for b0, b1 in self._batches():
            x = self.raw_input[b0:b1]
            ro = self._out_view(b0, b1)
            parts = []
            for p0, p1 in self._chunks():
                s, z = cs[p0:p1], cz[p0:p1]
                w = _wview(self.weight, self.n_V).unsqueeze(0)
                wq = ((w / s).round_() + z).clamp(0, L)
                wd = ((wq - z) * s).view(-1, self.in_f)
                b_sim = self.bias.repeat(p1 - p0) if self.bias is not None else None
                out = F.linear(*_up(self.aq(x), wd, b_sim))
                out = out.view(*out.shape[:-1], p1 - p0, self.n_V, -1)
                sim = _sim(_up(ro)[0], out)
                if sim.dim() > 4:
                    sim = torch.mean(sim, dim=list(range(1, sim.dim() - 3)))
                parts.append(sim.sum(dim=0, keepdim=True))
            per_batch.append(torch.cat(parts, dim=1))
        return torch.cat(per_batch, dim=0).sum(dim=0, keepdim=False)

    def eval_w(self, cs, cz, topk=1):
        sims = self.sims_w(cs, cz)
        idx = self.trace.topk(sims, topk, 0, 'w_out').reshape(topk, self.n_V, -1, 1)
        if topk == 1:
            self.wq.scale = torch.gather(cs, 0, idx).squeeze(0)
            self.wq.zero_point = torch.gather(cz, 0, idx).squeeze(0).float()
        return idx.squeeze(0)

    def _a_out_sims(self, make_xsim):
        """Shared reduction of linear.py:394-430 / :856-939: make_xsim(x4, p0, p1) -> [b,*,in,peq]."""
        per_batch = []
        for b0, b1 in self._batches():
            x = self.raw_input[b0:b1]
            ro = self.raw_out[b0:b1].unsqueeze(-2)
            parts = []
            for p0, p1 in self._chunks():
                w_sim = quant_weight(self.weight, self.wq, self.n_V)
                xs = make_xsim(x.unsqueeze(-1), p0, p1)
                xs = xs.permute(*list(range(xs.dim() - 2)), -1, -2)
                out = F.linear(*_up(xs, w_sim, self.bias))
                sim = torch.mean(_sim(_up(ro)[0], out), dim=-1)
                if sim.dim() > 2:
                    sim = torch.mean(sim, dim=list(range(1, sim.dim() - 1)))
                parts.append(torch.sum(sim, dim=0, keepdim=True))
            per_batch.append(torch.cat(parts, dim=1))
        return torch.cat(per_batch, dim=0).sum(dim=0, keepdim=True)

    # -- linear.py:394-430
    def sims_a(self, cs, cz):
        L = 2 * self.aq.n_levels - 1

        def make(x4, p0, p1):
            s, z = cs[:, p0:p1], cz[:, p0:p1]
            xq = ((x4 / s).round_() + z).clamp_(0, L)
            return (xq - z) * s
        return self._a_out_sims(make)

    def eval_a(self, cs, cz, topk=1):
        sims = self.sims_a(cs, cz)
        idx = self.trace.topk(sims, topk, -1, 'a_out')
        if topk == 1:
            self.aq.scale = torch.gather(cs, -1, idx).squeeze(-1)
            self.aq.zero_point = torch.gather(cz, -1, idx).squeeze(-1).float()
        return idx

    # -- linear.py:856-939 (log-base only when cs is None; joint scale x base otherwise)
    def _log_xsim(self, x4, s, qc):
        aq, nl = self.aq, self.aq.n_levels
        v = ((x4 + aq.shift) / s).clamp(min=1e-15, max=1.0)
        code = torch.round(-v.log2() * R_BASE / qc)
        mask = code >= 2 * nl
        code = code.clamp_(0, 2 * nl - 1)
        col = torch.remainder(code * qc, R_BASE).round_().long()
        xs = (2 ** (-1 * torch.floor(code * qc / R_BASE))) * self.table.to(x4.device)[col]
        xs[mask] = 0
        return xs * s - aq.shift

    def sims_log(self, cs, q_cands):
        """cs None: base-only search at the current scale (linear.py:856-890); else joint (:898-931)"""
        if cs is None:
            return self._a_out_sims(lambda x4, p0, p1: self._log_xsim(x4, self.aq.scale, q_cands[:, p0:p1]))
        return self._a_out_sims(lambda x4, p0, p1: self._log_xsim(x4, cs[:, p0:p1], q_cands[:, p0:p1]))

    def eval_log_base(self, q_cands=None, topk=1):
        if q_cands is None:
            q_cands = torch.tensor([i for i in range(10, 11 + self.eq_n)]).to(self.weight.device).view(1, -1)
        sims = self.sims_log(None, q_cands)
        idx = self.trace.topk(sims, topk, -1, 'log_base')
        if topk == 1:
            self.aq.q = torch.gather(q_cands, -1, idx).view(1)
            self.aq.update_table()
        return idx

    def eval_scale_logbase(self, cs, q_cands, topk=1):
        sims = self.sims_log(cs, q_cands)
        idx = self.trace.topk(sims, topk, -1, 'scale_logbase')
        if topk == 1:
            self.aq.scale = torch.gather(cs, -1, idx).squeeze(-1)
            self.aq.q = torch.gather(q_cands, -1, idx).view(1)
            self.aq.update_table()
        return idx

    # -- drivers
    def weight_fpcs(self, strategy):
        cs, cz = weight_candidates(self.weight, self.n_V, self.wq.n_levels, self.eq_n)
        fpcs(cs, cz, strategy, 0, self.eq_n, 16, self.steps)

    def activation_fpcs(self, strategy):
        cs, cz = activation_candidates(self.raw_input, self.aq.n_levels, self.eq_n, self.a_channel_wise)
        fpcs(cs, cz, strategy, -1, self.eq_n, 16, self.steps, clamp_min=1e-4)

    def search_asym(self):
        """linear.py:525-545."""
        self.init_calib()
        self.weight_fpcs(self.eval_w_self)
        self.activation_fpcs(self.eval_a_self)
        for _ in range(self.search_round):
            self.weight_fpcs(self.eval_w)
            self.activation_fpcs(self.eval_a)

    def search_channel_wise(self):
        """linear.py:585-594."""
        assert self.a_channel_wise
        self.init_calib()
        self.activation_fpcs(self.eval_a_self)

    def reparam(self, ln_weight, ln_bias):
        """linear.py:596-621: returns the rewritten (ln_weight, ln_bias); rewrites own weight/bias/raw_input,
        collapses the activation quantizer to per-tensor and runs the asymmetric search."""
        s, z = self.aq.scale, self.aq.zero_point
        ch_min = -z * s
        t_scale = torch.mean(s).view(1)
        t_zp = torch.mean(z).round().view(1)
        t_min = -t_zp * t_scale
        r = s / t_scale
        b = ch_min / r - t_min
        ln_weight = ln_weight / r
        ln_bias = ln_bias / r.view(-1) - b
        self.weight = self.weight * r.view(1, -1)
        mm = torch.mm(self.weight, b.reshape(-1, 1)).reshape(-1)
        self.bias = self.bias + mm if self.bias is not None else mm
        self.raw_input = self.raw_input / r - b
        self.a_channel_wise = False
        self.aq.scale, self.aq.zero_point = t_scale, t_zp
        self.search_asym()
        return ln_weight, ln_bias

    def postgelu_activation_fpcs(self, ud, base_num=8, scale_num=16, width=32):
        """linear.py:941-967."""
        dev = self.weight.device
        q_all = torch.tensor([i for i in range(10, 11 + self.eq_n)]).to(dev).view(1, -1)
        q_idx = self.eval_log_base(q_all, topk=base_num)
        lin = torch.tensor([i / (scale_num - 1) for i in range(scale_num)]).to(dev).view(1, -1)
        cs = ud[:, 0:1] + (ud[:, 1:] - ud[:, 0:1]) * lin
        delta = cs[:, 1:2] - cs[:, 0:1]
        cs = cs.repeat(1, base_num)
        qc = torch.gather(q_all, dim=-1, index=q_idx).repeat_interleave(scale_num, dim=-1)
        _fpcs_tail(cs, qc, delta, self.eval_scale_logbase, -1, int(self.eq_n / width), width, self.steps, None, dev)

    def eval_a_log_scale(self, cs, topk=1):
        """linear.py:816-854: scale-only search at the current base (the fpcs=False path)."""
        sims = self._a_out_sims(lambda x4, p0, p1: self._log_xsim(x4, cs[:, p0:p1], self.aq.q))
        idx = self.trace.topk(sims, topk, -1, 'log_scale')
        if topk == 1:
            self.aq.scale = torch.gather(cs, -1, idx).squeeze(-1)
            self.aq.update_table()
        return idx

    def search_postgelu(self, tmp_kind=None):
        """linear.py:969-997.  tmp_kind 'log2' / 'logsqrt2': the fixed-base quantizer swapped in after the AdaLog
        search (:990-994)."""
        self.init_calib()
        nl = self.wq.n_levels
        if self.fpcs_on:
            self.weight_fpcs(self.eval_w_self)
        else:
            self.eval_w_self(*weight_candidates(self.weight, self.n_V, nl, self.eq_n))
        ud, sc = postgelu_candidates(self.raw_input, self.aq.shift.item(), self.eq_n)
        self.aq.scale = sc[:, -2].clone()
        self.aq.update_table()
        for _ in range(self.search_round):
            if self.fpcs_on:
                self.postgelu_activation_fpcs(ud)
                self.weight_fpcs(self.eval_w)
            else:
                self.eval_log_base()
                self.eval_a_log_scale(sc)
                self.eval_w(*weight_candidates(self.weight, self.n_V, nl, self.eq_n))
        if tmp_kind is not None:
            self.aq = FixedLogQ(tmp_kind, self.aq.n_bits, scale=self.aq.scale.clone(), shift=self.aq.shift)

    # -- PostGeluTwinUniformBatchingQuantLinear, linear.py:624-721 (PTQ4ViT twin-uniform baseline)
    def twin_init_scale(self):
        """linear.py:647-662."""
        nl = self.aq.n_levels
        per_batch = []
        for b0, b1 in self._batches():
            x = self.raw_input[b0:b1]
            per_batch.append((x.abs().max() / (nl - 0.5)).view(1, 1).expand(1, self.aq.scale.shape[-1]))
        pos = torch.cat(per_batch, dim=0).amax(dim=0, keepdim=False).view(-1)
        neg = torch.tensor(SHIFT_GELU / nl, device=self.aq.scale.device).view(1).repeat(self.aq.scale.shape[-1])
        self.aq.scale = torch.stack([pos, neg]).clone()

    def sims_a_twin(self, cands):
        """linear.py:664-690: only the first cands.shape[-1] - 1 candidates are scored (:665-666)."""
        nl = self.aq.n_levels
        n = cands.shape[-1] - 1
        s_neg = self.aq.scale[1].unsqueeze(-1)
        per_batch = []
        for b0, b1 in self._batches():
            x = self.raw_input[b0:b1]
            ro = self.raw_out[b0:b1].unsqueeze(-2)
            parts = []
            for p0 in range(0, n, self.peq):
                p1 = min(n, p0 + self.peq)
                cur = cands[:, p0:p1]
                w_sim = quant_weight(self.weight, self.wq, self.n_V)
                x4 = x.unsqueeze(-1)
                x_pos = (x4 / cur).round_().clamp_(0, nl - 1) * cur
                x_neg = (x4 / s_neg).round_().clamp_(-nl, 0) * s_neg
                xs = x_pos + x_neg
                xs = xs.permute(*list(range(xs.dim() - 2)), -1, -2)
                out = F.linear(*_up(xs, w_sim, self.bias))
                sim = torch.mean(_sim(_up(ro)[0], out), dim=-1)
                if sim.dim() > 2:
                    sim = torch.mean(sim, dim=list(range(1, sim.dim() - 1)))
                parts.append(torch.sum(sim, dim=0, keepdim=True))
            per_batch.append(torch.cat(parts, dim=1))
        return torch.cat(per_batch, dim=0).sum(dim=0, keepdim=False)

    def eval_a_twin(self, cands):
        """linear.py:664-696 (argmax, not topk)."""
        sims = self.sims_a_twin(cands)
        idx = self.trace.argmax(sims, 'a_twin').reshape(1, -1)
        self.aq.scale = torch.stack([torch.gather(cands, -1, idx).squeeze(-1), self.aq.scale[1]])
        return idx.squeeze(0)

    def search_twin(self):
        """linear.py:698-721."""
        self.init_calib()
        self.twin_init_scale()
        nl = self.wq.n_levels
        if self.fpcs_on:
            self.weight_fpcs(self.eval_w_self)
        else:
            self.eval_w_self(*weight_candidates(self.weight, self.n_V, nl, self.eq_n))
        cands = torch.tensor([(2 ** i) for i in range(-5, 25)]).to(self.weight.device).view(1, -1) \
            * self.aq.scale[1].unsqueeze(-1)
        for _ in range(self.search_round):
            self.eval_a_twin(cands)
            if self.fpcs_on:
                self.weight_fpcs(self.eval_w)
            else:
                self.eval_w(*weight_candidates(self.weight, self.n_V, nl, self.eq_n))

    def reparam_bias(self):
        """linear.py:999-1006."""
        x_ = torch.full((1, self.in_f), -self.aq.shift.item()).to(self.weight.device)
        w_sim = quant_weight(self.weight, self.wq, self.n_V)
        self.bias = self.bias + (x_ @ w_sim.transpose(0, 1)).squeeze()
        self.aq.bias_reparamed = True


# ----------------------------------------------------------------------------------------------
# MatMul family
# ----------------------------------------------------------------------------------------------
class MatMulSearch:
    """matmul.py:109-283 (Q.K^T) and :286-378 (post-softmax P.V, AdaLog on A)."""

    def __init__(self, A, B, raw_out, A_bit, B_bit, num_heads, eq_n=128, calib_batch_size=32, search_round=3,
                 steps=6, head_channel_wise=True, memory=8 * 2 ** 30, trace=None, post_softmax=False,
                 quantizer='adalog'):
        self.A, self.B, self.raw_out = A, B, raw_out
        self.H, self.eq_n, self.bs = num_heads, eq_n, calib_batch_size
        self.search_round, self.steps, self.hcw = search_round, steps, head_channel_wise
        self.memory = memory
        self.trace = trace or Trace()
        self.post_softmax = post_softmax
        self.Bq = UQ(B_bit)
        self.adaptive = quantizer == 'adalog'
        if post_softmax and not self.adaptive:      # matmul.py:307-310: fixed-base baselines, scale 1, nothing to search
            self.Aq = FixedLogQ(quantizer, A_bit, scale=torch.ones(1, 1, 1, 1).to(A.device))
        elif post_softmax:
            self.Aq = LQ(A_bit, scale=torch.ones(1, 1, 1, 1).to(A.device))
            self.Aq.q = self.Aq.q.to(A.device)
            self.Aq.update_table()
            self.table = search_table(self.Aq.n_levels)
        else:
            self.Aq = UQ(A_bit)

    def init_calib(self):
        """matmul.py:95-106."""
        self.calib_size = self.A.shape[0]
        numel = (4 * self.A[:self.calib_size].numel() + 4 * self.B[:self.calib_size].numel()
                 + 8 * self.raw_out[:self.bs].numel())
        self.peq = chunked_eq_n(self.eq_n, self.memory, numel)

    def _reduce(self, make_pair, pool_heads):
        """matmul.py:137-163 / :175-201 / :325-351: make_pair(A, B, p0, p1) -> (A_sim, B_sim) broadcastable."""
        per_batch = []
        for b0 in range(0, self.calib_size, self.bs):
            b1 = min(self.calib_size, b0 + self.bs)
            A, B = self.A[b0:b1], self.B[b0:b1]
            ro = self.raw_out[b0:b1].unsqueeze(0)
            parts = []
            for p0 in range(0, self.eq_n, self.peq):
                p1 = min(self.eq_n, p0 + self.peq)
                a_s, b_s = _up(*make_pair(A, B, p0, p1))
                sim = _sim(_up(ro)[0], a_s @ b_s)
                if self.hcw and not pool_heads:
                    sim = torch.mean(sim, dim=list(range(3, sim.dim())))
                else:
                    sim = torch.mean(sim, dim=list(range(2, sim.dim())))
                parts.append(sim.sum(dim=1, keepdim=True))
            per_batch.append(torch.cat(parts, 0))
        return torch.cat(per_batch, dim=1)

    @staticmethod
    def _cand_quant(x, s, z, L):
        xq = ((x / s).round_() + z).clamp(0, L)
        return (xq - z).mul_(s)

    def sims_A(self, cs, cz):
        L = 2 * self.Aq.n_levels - 1
        return self._reduce(lambda A, B, p0, p1: (self._cand_quant(A, cs[p0:p1], cz[p0:p1], L),
                                                  self.Bq(B).unsqueeze(0)), False).sum(dim=1, keepdim=False)

    def eval_A(self, cs, cz, topk=1):
        """matmul.py:135-171."""
        sims = self.sims_A(cs, cz)
        idx = self.trace.topk(sims, topk, 0, 'mm_A').view(topk, 1, -1, 1, 1)
        if topk == 1:
            self.Aq.scale = torch.gather(cs, 0, idx).view(self.Aq.scale.shape)
            self.Aq.zero_point = torch.gather(cz, 0, idx).view(self.Aq.scale.shape).float()
        return idx

    def sims_B(self, cs, cz):
        L = 2 * self.Bq.n_levels - 1
        return self._reduce(lambda A, B, p0, p1: (self.Aq(A).unsqueeze(0),
                                                  self._cand_quant(B, cs[p0:p1], cz[p0:p1], L)),
                            False).sum(dim=1, keepdim=False)

    def eval_B(self, cs, cz, topk=1):
        """matmul.py:173-209."""
        sims = self.sims_B(cs, cz)
        idx = self.trace.topk(sims, topk, 0, 'mm_B').view(topk, 1, -1, 1, 1)
        if topk == 1:
            self.Bq.scale = torch.gather(cs, 0, idx).view(self.Bq.scale.shape)
            self.Bq.zero_point = torch.gather(cz, 0, idx).view(self.Bq.scale.shape).float()
        return idx

    def sims_A_log_base(self, q_cands):
        nl = self.Aq.n_levels

        def make(A, B, p0, p1):
            qc = q_cands[p0:p1]
            code = torch.round(-A.log2() * R_BASE / qc)
            mask = code >= 2 * nl
            code = code.clamp_(0, 2 * nl - 1)
            col = torch.remainder(code * qc, R_BASE).round_().long()
            a_s = (2 ** (-1 * torch.floor(code * qc / R_BASE))) * self.table.to(A.device)[col]
            a_s[mask] = 0
            return a_s, self.Bq(B).unsqueeze(0)
        return self._reduce(make, True).sum(dim=1, keepdim=True)

    def eval_A_log_base(self, q_cands=None, topk=1):
        """matmul.py:321-358."""
        if q_cands is None:
            q_cands = torch.tensor([i for i in range(10, 11 + self.eq_n)]).to(self.A.device).view(-1, 1, 1, 1, 1)
        sims = self.sims_A_log_base(q_cands)
        idx = self.trace.topk(sims, topk, 0, 'mm_logbase').view(topk, 1, 1, 1, 1)
        if topk == 1:
            self.Aq.q = torch.gather(q_cands, 0, idx).view(1)
            self.Aq.update_table()
        return idx

    def _fpcs(self, x, strategy):
        cs, cz = matmul_candidates(x, self.Bq.n_levels, self.eq_n, self.hcw)
        fpcs(cs, cz, strategy, 0, self.eq_n, 16, self.steps)

    def _init_from(self, quant, x):
        cs, cz = matmul_candidates(x, self.Bq.n_levels, self.eq_n, self.hcw)
        quant.scale = cs[-2].clone()
        quant.zero_point = cz[-2].clone().float()

    def search(self):
        """matmul.py:264-283 / :360-378."""
        self.init_calib()
        if not self.post_softmax:
            self._init_from(self.Aq, self.A)
        self._init_from(self.Bq, self.B)
        for _ in range(self.search_round):
            if self.post_softmax:
                if self.adaptive:
                    self.eval_A_log_base()
            else:
                self._fpcs(self.A, self.eval_A)
            self._fpcs(self.B, self.eval_B)
            if self.post_softmax and not self.adaptive:
                break                                   # matmul.py:374-375


# ----------------------------------------------------------------------------------------------
# Conv (patch-embed): weight-only output-error search, conv.py:199-334
# ----------------------------------------------------------------------------------------------
class ConvSearch:
    def __init__(self, weight, bias, raw_input, raw_out, w_bit, stride, eq_n=128, calib_batch_size=32, steps=6,
                 memory=8 * 2 ** 30, trace=None, padding=0, dilation=1, groups=1):
        self.weight, self.bias, self.raw_input, self.raw_out = weight, bias, raw_input, raw_out
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.eq_n, self.bs, self.steps, self.memory = eq_n, calib_batch_size, steps, memory
        self.trace = trace or Trace()
        self.wq = UQ(w_bit)

    def sims_w(self, cs, cz):
        L = 2 * self.wq.n_levels - 1
        oc, ic, kw, kh = self.weight.shape
        per_batch = []
        for b0 in range(0, self.calib_size, self.bs):
            b1 = min(self.calib_size, b0 + self.bs)
            x = self.raw_input[b0:b1]
            ro = self.raw_out[b0:b1].unsqueeze(1)
            parts = []
            for p0 in range(0, self.eq_n, self.peq):
                p1 = min(self.eq_n, p0 + self.peq)
                s, z = cs[p0:p1], cz[p0:p1]
                w = self.weight.view(oc, -1).unsqueeze(0)
                wq = ((w / s).round_() + z).clamp(0, L)
                wd = (wq - z).mul_(s).view(-1, ic, kw, kh)
                b_sim = self.bias.repeat(p1 - p0) if self.bias is not None else None
                out = F.conv2d(*_up(x, wd, b_sim), self.stride, self.padding, self.dilation, self.groups)
                out = torch.cat(torch.chunk(out.unsqueeze(1), chunks=p1 - p0, dim=2), dim=1)
                sim = torch.mean(_sim(_up(ro)[0], out), [3, 4])
                parts.append(torch.sum(sim, dim=0, keepdim=True))
            per_batch.append(torch.cat(parts, dim=1))
        return torch.cat(per_batch, dim=0).sum(dim=0, keepdim=False)

    def eval_w(self, cs, cz, topk=1):
        """conv.py:226-263 (a_bit >= 8 -> raw FP32 input, conv.py:55-58)."""
        sims = self.sims_w(cs, cz)
        idx = self.trace.topk(sims, topk, 0, 'conv_w').view(topk, -1, 1)
        if topk == 1:
            self.wq.scale = torch.gather(cs, 0, idx).squeeze(dim=0)
            self.wq.zero_point = torch.gather(cz, 0, idx).squeeze(dim=0).float()
        return idx

    def init_calib(self):
        self.calib_size = self.raw_input.shape[0]
        numel = 2 * self.raw_input[:self.bs].numel() + 2 * self.raw_out[:self.bs].numel()
        self.peq = chunked_eq_n(self.eq_n, self.memory, numel)

    def search(self):
        """conv.py:313-334 with a_bit >= 8: one FPCS round then break (:328-331)."""
        self.init_calib()
        cs, cz = conv_weight_candidates(self.weight, self.wq.n_levels, self.eq_n)
        self.wq.scale, self.wq.zero_point = cs[-2].clone(), cz[-2].clone().float()
        fpcs(cs, cz, self.eval_w, 0, self.eq_n, 16, self.steps)
