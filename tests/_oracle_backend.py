"""TEST-ONLY stand-in for the CUDA sweeps, built on the oracle.

The product has no CPU path.  To exercise its *host* logic (candidate seeding, FPCS refinement, selection,
reparameterisation, calibrator, wrap_net, checkpoint layout) in the CPU-only container, the tests monkeypatch the
kernel-call layer (adalog_b200.sweep.* scoring functions and adalog_b200.ops.*_fakequant) with the oracle's
similarity functions.  Nothing here is importable from the product.
"""
import torch

import adalog_oracle as O
from adalog_b200.utils import dist as adist


def _dp(sims):
    """similarities are sums over samples: under data parallelism the shards' values add (as the FP64 error sums do
    in the real sweeps)"""
    return adist.all_reduce_sum(sims.clone())


class Cfg:
    bs = 4
    memory = 8 * 2 ** 30


def _uq_from(q):
    u = O.UQ(q.n_bits)
    u.scale, u.zero_point = q.scale.detach(), q.zero_point.detach()
    return u


def _lq_from(q):
    l = O.LQ(q.n_bits, scale=q.scale.detach(), q=q.q.detach().clone(), shift=getattr(q, 'shift', None))
    if l.shift is not None:
        l.shift = l.shift.detach()
    l.table1, l.table2 = q.table1.detach(), q.table2.detach()
    l.bias_reparamed = bool(getattr(q, 'bias_reparamed', False))
    return l


def _any_q_from(q):
    """oracle state holder for whatever quantizer object the product hands to a sweep"""
    name = type(q).__name__
    if name == 'TwinUniformQuantizer':
        return O.TQ(q.n_bits, scale=q.scale.detach())
    if getattr(q, 'is_log', False) and not hasattr(q, 'table2'):
        sh = getattr(q, 'shift', None)
        return O.FixedLogQ('log2' if 'Sqrt' not in name else 'logsqrt2', q.n_bits, scale=q.scale.detach(),
                           shift=None if sh is None else sh.detach(), bias_reparamed=bool(getattr(q, 'bias_reparamed', False)))
    if getattr(q, 'is_log', False):
        return _lq_from(q)
    return _uq_from(q)


def _lin(ctx, weight3, bias, w_bit=4, a_bit=4, a_kind='uniform', cw=False):
    n_V = weight3.shape[0]
    s = O.LinearSearch(weight3.detach().reshape(-1, weight3.shape[-1]), None if bias is None else bias.detach(),
                       ctx.raw_input, ctx.raw_out, w_bit, a_bit, n_V=n_V, calib_batch_size=Cfg.bs, memory=Cfg.memory,
                       a_kind=a_kind, a_channel_wise=cw)
    s.init_calib()
    return s


def _bits(n_levels):
    return n_levels.bit_length()


def linear_err_w_self(weight3, cs, cz, n_levels):
    class _C:
        raw_input = torch.zeros(1, 1, weight3.shape[-1])
        raw_out = torch.zeros(1, 1, weight3.shape[0] * weight3.shape[1])
    s = _lin(_C, weight3, None, w_bit=_bits(n_levels))
    s.peq = s.eq_n = cs.shape[0]
    return s.sims_w_self(cs, cz)


def linear_err_a_self(ctx, cs, cz, n_levels, channel_wise):
    w3 = torch.zeros(1, ctx.raw_out.shape[-1], ctx.raw_input.shape[-1])
    s = _lin(ctx, w3, None, a_bit=_bits(n_levels), cw=channel_wise)
    return _dp(s.sims_a_self(cs, cz))


def linear_err_w(ctx, weight3, bias, aq, cs, cz, n_levels_w):
    log = getattr(aq, 'is_log', False)
    s = _lin(ctx, weight3, bias, w_bit=_bits(n_levels_w), a_bit=aq.n_bits, a_kind='adalog' if log else 'uniform')
    s.aq = _any_q_from(aq)
    return _dp(s.sims_w(cs, cz))


def linear_err_a(ctx, weight3, bias, wq, cs, cz, n_levels_a):
    s = _lin(ctx, weight3, bias, w_bit=wq.n_bits, a_bit=_bits(n_levels_a))
    s.wq = _uq_from(wq)
    return _dp(s.sims_a(cs, cz))


def linear_err_a_twin(ctx, weight3, bias, wq, s_neg, cands, n_levels):
    s = _lin(ctx, weight3, bias, w_bit=wq.n_bits, a_bit=_bits(n_levels), a_kind='twin')
    s.wq = _uq_from(wq)
    s.aq.scale = torch.stack([torch.zeros_like(s_neg), s_neg.detach()])
    pad = torch.cat([cands, cands[:, -1:]], dim=-1)          # the oracle (like the reference) drops the last column
    return _dp(s.sims_a_twin(pad))


def linear_err_log(ctx, weight3, bias, wq, aq, cs, cq):
    s = _lin(ctx, weight3, bias, w_bit=wq.n_bits, a_bit=aq.n_bits, a_kind='adalog')
    s.wq = _uq_from(wq)
    s.aq = _lq_from(aq)
    return _dp(s.sims_log(cs, cq))


def _mm(ctx, A_bit, B_bit, hcw, post_softmax=False):
    A, B = ctx.raw_input
    s = O.MatMulSearch(A, B, ctx.raw_out, A_bit, B_bit, A.shape[1], calib_batch_size=Cfg.bs, head_channel_wise=hcw,
                       memory=Cfg.memory, post_softmax=post_softmax)
    s.init_calib()
    return s


def matmul_err_A(ctx, Bq, cs, cz, n_levels_A, hcw):
    s = _mm(ctx, _bits(n_levels_A), Bq.n_bits, hcw)
    s.Bq = _uq_from(Bq)
    return _dp(s.sims_A(cs, cz))


def matmul_err_B(ctx, Aq, cs, cz, n_levels_B, hcw):
    log = getattr(Aq, 'is_log', False)
    s = _mm(ctx, Aq.n_bits, _bits(n_levels_B), hcw, post_softmax=log)
    s.Aq = _any_q_from(Aq)
    return _dp(s.sims_B(cs, cz))


def matmul_err_A_log_base(ctx, Bq, cq, n_levels_A):
    s = _mm(ctx, _bits(n_levels_A), Bq.n_bits, True, post_softmax=True)
    s.Bq = _uq_from(Bq)
    return _dp(s.sims_A_log_base(cq))


class ConvCtx:
    def __init__(self, raw_input, raw_out, kernel_size):
        self.raw_input, self.raw_out, self.kernel_size = raw_input, raw_out, kernel_size


def conv_err_w(ctx, weight2d, bias, cs, cz, n_levels_w):
    k = ctx.kernel_size
    ic = ctx.raw_input.shape[1]
    w4 = weight2d.detach().reshape(weight2d.shape[0], ic, k[0], k[1])
    s = O.ConvSearch(w4, None if bias is None else bias.detach(), ctx.raw_input, ctx.raw_out, _bits(n_levels_w), k,
                     calib_batch_size=Cfg.bs, memory=Cfg.memory)
    s.init_calib()
    return _dp(s.sims_w(cs, cz))


def uniform_fakequant(x, scale, zero_point, n_levels, sym=False, want_codes=False, want_y=True):
    y, c = O.uniform_fakequant(x, scale.detach(), None if sym else zero_point.detach(), n_levels, sym, return_codes=True)
    if want_codes:
        return (y, c.to(torch.int16)) if want_y else c.to(torch.int16)
    return y


def log_fakequant(x, scale, kind, n_levels, q=None, table1=None, table2=None, shift=None, sub_shift=False,
                  want_codes=False):
    xs = x if shift is None else x + shift.detach()
    s = scale.detach()
    if kind == 0:
        y, c = O.log2_fakequant(xs, s, n_levels, True)
    elif kind == 1:
        y, c = O.logsqrt2_fakequant(xs, s, n_levels, True)
    else:
        y, c = O.adalog_fakequant(xs, s, q, n_levels, table1, table2, True)
    if sub_shift:
        y = y - shift.detach()
    return (y, c.to(torch.int16)) if want_codes else y


def twin_fakequant(x, scale2, n_levels):
    return O.twin_uniform_fakequant(x, scale2.detach(), n_levels)


def install(monkeypatch, bs=4, memory=8 * 2 ** 30):
    """Route the product's kernel-call layer to the oracle (tests only)."""
    from adalog_b200 import ops, sweep
    Cfg.bs, Cfg.memory = bs, memory
    monkeypatch.setattr(sweep, 'require_cuda', lambda dev: None)
    for name in ('linear_err_w_self', 'linear_err_a_self', 'linear_err_w', 'linear_err_a', 'linear_err_log',
                 'linear_err_a_twin',
                 'matmul_err_A', 'matmul_err_B', 'matmul_err_A_log_base', 'conv_err_w'):
        monkeypatch.setattr(sweep, name, globals()[name])
    monkeypatch.setattr(sweep, 'ConvCtx', ConvCtx)
    for name in ('uniform_fakequant', 'log_fakequant', 'twin_fakequant'):
        monkeypatch.setattr(ops, name, globals()[name])
