"""Quantised attention matmuls with the reference's class API (quant_layers/matmul.py).

Q.K^T (`AsymmetricallyBatchingQuantMatMul`) and post-softmax P.V (`PostSoftmaxAsymmetricallyBatchingQuantMatMul`,
AdaLog on P).  Each search evaluation is one fused device sweep over all (sample, head) pairs: the candidate
operand is expanded 128x on the fly into bf16 integer tiles, multiplied on tcgen05 against the fixed quantised
operand, and reduced per (candidate, head) in the epilogue (adalog_b200/sweep.py).
"""
import torch
from torch import nn

from .. import _const, sweep
from ..quantizers.uniform import UniformQuantizer
from ..quantizers.logarithm import Log2Quantizer, LogSqrt2Quantizer, AdaLogQuantizer
from ..utils import dist as adist
from ..quantizers._ste import assign
from . import _fpcs

__all__ = ['MinMaxQuantMatMul', 'PTQSLQuantMatMul', 'PTQSLBatchingQuantMatMul', 'AsymmetricallyBatchingQuantMatMul',
           'PostSoftmaxAsymmetricallyBatchingQuantMatMul']


class MinMaxQuantMatMul(nn.Module):
    """reference: quant_layers/matmul.py:13-45"""

    def __init__(self, A_bit=8, B_bit=8, mode="raw"):
        super().__init__()
        self.mode = mode
        self.A_quantizer = UniformQuantizer(n_bits=A_bit, symmetric=True, channel_wise=False)
        self.B_quantizer = UniformQuantizer(n_bits=B_bit, symmetric=True, channel_wise=False)
        self.raw_input = None
        self.raw_out = None
        self.tmp_input = None
        self.tmp_out = None
        self.calibrated = False

    def forward(self, A, B):
        if self.mode == 'raw':
            return A @ B
        if self.mode == 'quant_forward':
            return self.quant_forward(A, B)
        raise NotImplementedError

    def quant_input_A(self, x):
        return self.A_quantizer(x)

    def quant_input_B(self, x):
        return self.B_quantizer(x)

    def quant_forward(self, A, B):
        assert self.calibrated, f"Module should be calibrated before run quant_forward for {self}"
        return self.quant_input_A(A) @ self.quant_input_B(B)


class PTQSLQuantMatMul(MinMaxQuantMatMul):
    """reference: quant_layers/matmul.py:48-79.  Q@K: A [B,H,S,C], B [B,H,C,S];  P@V: A [B,H,S,S], B [B,H,S,C]."""

    def __init__(self, A_bit=8, B_bit=8, mode="raw", search_round=1, eq_n=100, head_channel_wise=True, num_heads=12):
        super().__init__(A_bit, B_bit, mode)
        self.A_quantizer = UniformQuantizer(n_bits=A_bit, symmetric=True, channel_wise=head_channel_wise)
        self.B_quantizer = UniformQuantizer(n_bits=B_bit, symmetric=True, channel_wise=head_channel_wise)
        self.search_round = search_round
        self.eq_n = eq_n
        self.head_channel_wise = head_channel_wise
        self.num_heads = num_heads
        shape = [1, self.num_heads, 1, 1] if self.head_channel_wise else [1, 1, 1, 1]
        self.A_quantizer.scale = nn.Parameter(torch.zeros(*shape))
        self.B_quantizer.scale = nn.Parameter(torch.zeros(*shape))

    def _get_similarity(self, tensor_raw, tensor_sim):
        return -(tensor_raw - tensor_sim) ** 2


class PTQSLBatchingQuantMatMul(PTQSLQuantMatMul):
    """reference: quant_layers/matmul.py:82-106"""

    def __init__(self, A_bit=8, B_bit=8, mode="raw", calib_batch_size=32, search_round=1, eq_n=100,
                 head_channel_wise=True, num_heads=12):
        super().__init__(A_bit, B_bit, mode, search_round, eq_n, head_channel_wise, num_heads)
        self.calib_batch_size = calib_batch_size

    def _device(self):
        return self.B_quantizer.scale.device

    def _initialize_calib_parameters(self):
        """reference matmul.py:95-106; both operands and the raw output stay resident in HBM"""
        dev = self._device()
        sweep.require_cuda(dev)
        self.calib_size = self.raw_input[0].shape[0]
        self.parallel_eq_n = self.eq_n
        self.__dict__.pop('_pct_cache', None)
        self._ctx = sweep.MatMulCtx(self.raw_input[0].to(dev), self.raw_input[1].to(dev), self.raw_out.to(dev))


class AsymmetricallyBatchingQuantMatMul(PTQSLBatchingQuantMatMul):
    """reference: quant_layers/matmul.py:109-283"""

    def __init__(self, A_bit=8, B_bit=8, mode="raw", calib_batch_size=32, search_round=1, eq_n=128,
                 head_channel_wise=True, num_heads=12, fpcs=False, steps=4):
        super().__init__(A_bit, B_bit, mode, calib_batch_size, search_round, eq_n, head_channel_wise, num_heads)
        self.fpcs = fpcs
        self.steps = steps
        del self.A_quantizer, self.B_quantizer
        self.A_quantizer = UniformQuantizer(n_bits=A_bit, symmetric=False, channel_wise=head_channel_wise)
        self.B_quantizer = UniformQuantizer(n_bits=B_bit, symmetric=False, channel_wise=head_channel_wise)
        shape = [1, self.num_heads, 1, 1] if self.head_channel_wise else [1, 1, 1, 1]
        self.A_quantizer.scale = nn.Parameter(torch.zeros(*shape))
        self.B_quantizer.scale = nn.Parameter(torch.zeros(*shape))
        self.A_quantizer.zero_point = nn.Parameter(torch.zeros(*shape))
        self.B_quantizer.zero_point = nn.Parameter(torch.zeros(*shape))

    def _scored(self, fn, cs, cz):
        parts = [fn(cs[p0:p1], cz[p0:p1]) for p0, p1 in _fpcs.candidate_chunks(cs.shape[0])]
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)

    def _pick(self, sims, quantizer, cs, cz, topk):
        _, best = torch.topk(sims, k=topk, dim=0)
        best = best.view(topk, 1, -1, 1, 1)
        if topk == 1:
            assign(quantizer.scale, torch.gather(cs, dim=0, index=best).view(quantizer.scale.shape))
            assign(quantizer.zero_point, torch.gather(cz, dim=0, index=best).view(quantizer.zero_point.shape))
        return best

    def _search_best_A_scale(self, A_scale_candidates, A_zero_point_candidates, topk=1):
        """reference matmul.py:135-171"""
        sims = self._scored(lambda s, z: sweep.matmul_err_A(self._ctx, self.B_quantizer, s, z, self.A_quantizer.n_levels,
                                                            self.head_channel_wise),
                            A_scale_candidates, A_zero_point_candidates)
        return self._pick(sims, self.A_quantizer, A_scale_candidates, A_zero_point_candidates, topk)

    def _search_best_B_scale(self, B_scale_candidates, B_zero_point_candidates, topk=1):
        """reference matmul.py:173-209"""
        sims = self._scored(lambda s, z: sweep.matmul_err_B(self._ctx, self.A_quantizer, s, z, self.B_quantizer.n_levels,
                                                            self.head_channel_wise),
                            B_scale_candidates, B_zero_point_candidates)
        return self._pick(sims, self.B_quantizer, B_scale_candidates, B_zero_point_candidates, topk)

    def calculate_percentile_candidates(self, x, l=0.9, r=1.0):
        """reference matmul.py:211-240 (num_zp follows B_quantizer for both operands, :212); x: local shard"""
        nl = self.B_quantizer.n_levels
        num_zp = min(16, nl)
        num_scale = int(self.eq_n / num_zp)
        pct = torch.tensor([l, r])
        # the operands of a module do not change between its search rounds: sort each once
        key = (x.data_ptr(), tuple(x.shape), x._version, bool(self.head_channel_wise), l, r)
        cache = self.__dict__.setdefault('_pct_cache', {})
        if key not in cache:
            if len(cache) >= 2:                    # one entry per operand (A and B alternate inside a search round)
                cache.clear()
            if self.head_channel_wise:             # x: this rank's samples; chunked_quantile selects across ranks
                x_ = x.transpose(0, 1).contiguous()
                x_ = x_.view(x_.shape[0], 1, -1)
            else:
                x_ = x.reshape(1, 1, -1)
            cache[key] = _fpcs.chunked_quantile(x_, pct)
        up, lo = cache[key]
        d_min = (up[0] - lo[0]).view(1, 1, -1, 1, 1)
        d_max = (up[1] - lo[1]).view(1, 1, -1, 1, 1)
        return _fpcs.percentile_grid(d_min, d_max, nl, num_zp, num_scale, 0, 4)

    def _fpcs(self, x, fpcs_width=16, steps=6, search_strategy=None):
        """reference matmul.py:243-262"""
        cs, cz = self.calculate_percentile_candidates(x)
        _fpcs.search(cs, cz, lambda s, z, k: search_strategy(self, s, z, topk=k), 0, self.eq_n, fpcs_width, steps)

    def _operand(self, i):
        return self.raw_input[i].to(self._device())

    def _seed(self, quantizer, cs, cz):
        assign(quantizer.scale, cs[-2])
        assign(quantizer.zero_point, cz[-2])
        quantizer.inited = True

    def _finish(self):
        self.calibrated = True
        del self.raw_input, self.raw_out
        self._ctx = None
        self.__dict__.pop('_pct_cache', None)      # the sorted operands belong to this calibration only

    def hyperparameter_searching(self):
        """reference matmul.py:264-283"""
        cls = AsymmetricallyBatchingQuantMatMul
        self._initialize_calib_parameters()
        A_cs, A_cz = self.calculate_percentile_candidates(self._operand(0))
        B_cs, B_cz = self.calculate_percentile_candidates(self._operand(1))
        self._seed(self.A_quantizer, A_cs, A_cz)
        self._seed(self.B_quantizer, B_cs, B_cz)
        for _ in range(self.search_round):
            if self.fpcs:
                self._fpcs(self._operand(0), steps=self.steps, search_strategy=cls._search_best_A_scale)
                self._fpcs(self._operand(1), steps=self.steps, search_strategy=cls._search_best_B_scale)
            else:
                self._search_best_A_scale(A_cs, A_cz)
                self._search_best_B_scale(B_cs, B_cz)
        self._finish()
        return None


class PostSoftmaxAsymmetricallyBatchingQuantMatMul(AsymmetricallyBatchingQuantMatMul):
    """reference: quant_layers/matmul.py:286-378: A = softmax probabilities, quantised by AdaLog with a searched base"""

    def __init__(self, A_bit=8, B_bit=8, mode="raw", calib_batch_size=32, search_round=1, eq_n=100,
                 head_channel_wise=True, num_heads=12, fpcs=False, steps=4, quantizer='adalog'):
        super().__init__(A_bit, B_bit, mode, calib_batch_size, search_round, eq_n, head_channel_wise, num_heads, fpcs,
                         steps)
        del self.A_quantizer
        if quantizer == 'log2':
            self.A_quantizer = Log2Quantizer(n_bits=A_bit, symmetric=False, channel_wise=False)
        elif quantizer == 'logsqrt2':
            self.A_quantizer = LogSqrt2Quantizer(n_bits=A_bit, symmetric=False, channel_wise=False)
        elif quantizer == 'adalog':
            self.A_quantizer = AdaLogQuantizer(n_bits=A_bit, symmetric=False, channel_wise=False)
            self.table = torch.tensor([2 ** (-j / self.A_quantizer.r) for j in range(120)])
            self.table_scale = 1. / (4 * self.A_quantizer.n_levels - 2)
            self.table = torch.round(self.table / self.table_scale) * self.table_scale
        else:
            raise NotImplementedError(f"quantizer {quantizer} not implemented!")
        self.A_quantizer.scale = nn.Parameter(torch.ones([1, 1, 1, 1]))
        self.A_quantizer.inited = True

    def _search_best_A_log_base(self, q_candidates=None, topk=1):
        """reference matmul.py:321-358 (bases 10..137; one base for all heads)"""
        if q_candidates is None:
            q_candidates = _const.int_range(10, 11 + self.eq_n, self._device()).view(-1, 1, 1, 1, 1)
        nl = self.A_quantizer.n_levels
        parts = [sweep.matmul_err_A_log_base(self._ctx, self.B_quantizer, q_candidates[p0:p1], nl)
                 for p0, p1 in _fpcs.candidate_chunks(self.eq_n)]
        sims = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        _, best = torch.topk(sims, k=topk, dim=0)
        best = best.view(topk, 1, 1, 1, 1)
        if topk == 1:
            assign(self.A_quantizer.q, torch.gather(q_candidates, dim=0, index=best).view(*self.A_quantizer.q.shape))
            self.A_quantizer.update_table()
        return best

    def hyperparameter_searching(self):
        """reference matmul.py:360-378"""
        self._initialize_calib_parameters()
        B_cs, B_cz = self.calculate_percentile_candidates(self._operand(1))
        self._seed(self.B_quantizer, B_cs, B_cz)
        adaptive = isinstance(self.A_quantizer, AdaLogQuantizer)
        for _ in range(self.search_round):
            if adaptive:
                self._search_best_A_log_base()
            if self.fpcs:
                self._fpcs(self._operand(1), steps=self.steps,
                           search_strategy=AsymmetricallyBatchingQuantMatMul._search_best_B_scale)
            else:
                self._search_best_B_scale(B_cs, B_cz)
            if not adaptive:
                break
        self._finish()
        return None
