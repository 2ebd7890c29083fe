"""Quantised nn.Conv2d family with the reference's class API (quant_layers/conv.py).

Only the patch-embedding convolution is ever wrapped (utils/wrap_net.py:78-96): kernel == stride, no padding,
weight-only search because qconv_a_bit = 8.  The search treats it as the GEMM it is (im2col is a pure reshape for
non-overlapping patches) and runs the same tcgen05 candidate sweep as the linear layers; the unquantised FP32 input
is carried exactly as three bf16 pieces (sweep.ConvCtx).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import sweep
from ..quantizers.uniform import UniformQuantizer
from ..quantizers._ste import assign
from . import _fpcs

__all__ = ['MinMaxQuantConv2d', 'PTQSLQuantConv2d', 'PTQSLBatchingQuantConv2d', 'AsymmetricallyBatchingQuantConv2d']


class MinMaxQuantConv2d(nn.Conv2d):
    """reference: quant_layers/conv.py:11-75"""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1,
                 groups: int = 1, bias: bool = True, padding_mode: str = 'zeros', mode='raw', w_bit=8, a_bit=8):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode)
        self.mode = mode
        self.w_quantizer = UniformQuantizer(n_bits=w_bit, symmetric=True, channel_wise=False)
        self.a_quantizer = UniformQuantizer(n_bits=a_bit, symmetric=True, channel_wise=False)
        self.raw_input = None
        self.raw_out = None
        self.tmp_input = None
        self.tmp_out = None
        self.calibrated = False

    def _conv(self, x, w, b):
        return F.conv2d(x, w, b, self.stride, self.padding, self.dilation, self.groups)

    def forward(self, x):
        if self.mode == 'raw':
            return self._conv(x, self.weight, self.bias)
        if self.mode == 'quant_forward':
            return self.quant_forward(x)
        if self.mode == 'debug_only_quant_weight':
            return self.debug_only_quant_weight(x)
        if self.mode == 'debug_only_quant_act':
            return self.debug_only_quant_act(x)
        raise NotImplementedError

    def quant_weight_bias(self):
        return self.w_quantizer(self.weight), self.bias if self.bias is not None else None

    def quant_input(self, x):
        if self.a_quantizer.n_bits >= 8:
            return x
        return self.a_quantizer(x)

    def quant_forward(self, x):
        assert self.calibrated, f"Module should be calibrated before run quant_forward for {self}"
        w_sim, bias_sim = self.quant_weight_bias()
        return self._conv(self.quant_input(x), w_sim, bias_sim)

    def debug_only_quant_weight(self, x):
        w_sim, bias_sim = self.quant_weight_bias()
        return self._conv(x, w_sim, bias_sim)

    def debug_only_quant_act(self, x):
        return self._conv(self.quant_input(x), self.weight, self.bias)


class PTQSLQuantConv2d(MinMaxQuantConv2d):
    """reference: quant_layers/conv.py:78-120 (per-output-channel weight quantisation on the [oc, ic*kh*kw] view)"""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1,
                 groups: int = 1, bias: bool = True, padding_mode: str = 'zeros', mode='raw', w_bit=8, a_bit=8,
                 search_round=1, eq_n=100):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode,
                         mode, w_bit, a_bit)
        self.w_quantizer = UniformQuantizer(n_bits=w_bit, symmetric=True, channel_wise=True)
        self.a_quantizer = UniformQuantizer(n_bits=a_bit, symmetric=True, channel_wise=False)
        self.search_round = search_round
        self.eq_n = eq_n
        self.parallel_eq_n = eq_n
        self.w_quantizer.scale = nn.Parameter(torch.zeros((self.out_channels, 1)))
        self.a_quantizer.scale = nn.Parameter(torch.zeros((1, 1, 1, 1)))

    def _get_similarity(self, tensor_raw, tensor_sim):
        return -(tensor_raw - tensor_sim) ** 2

    def _weight2(self):
        return self.weight.view(self.out_channels, -1)

    def quant_weight_bias(self):
        w_sim = self.w_quantizer(self._weight2()).view(self.weight.shape)
        return w_sim, self.bias if self.bias is not None else None


class PTQSLBatchingQuantConv2d(PTQSLQuantConv2d):
    """reference: quant_layers/conv.py:123-196"""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1,
                 groups: int = 1, bias: bool = True, padding_mode: str = 'zeros', mode='raw', w_bit=8, a_bit=8,
                 calib_batch_size=32, search_round=1, eq_n=100):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode,
                         mode, w_bit, a_bit, search_round, eq_n)
        self.calib_batch_size = calib_batch_size

    def _initialize_calib_parameters(self):
        """reference conv.py:143-153"""
        dev = self.weight.device
        sweep.require_cuda(dev)
        pad = self.padding if isinstance(self.padding, tuple) else (self.padding, self.padding)
        if tuple(self.kernel_size) != tuple(self.stride) or any(pad) or self.groups != 1 or any(
                d != 1 for d in self.dilation):
            raise NotImplementedError('the calibration sweep covers patch-embedding convolutions '
                                      '(kernel == stride, no padding, groups=1), the only ones wrap_net creates')
        self.calib_size = self.raw_input.shape[0]
        self.parallel_eq_n = self.eq_n
        self._ctx = sweep.ConvCtx(self.raw_input.to(dev), self.raw_out.to(dev), tuple(self.kernel_size))


class AsymmetricallyBatchingQuantConv2d(PTQSLBatchingQuantConv2d):
    """reference: quant_layers/conv.py:199-334"""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, stride=1, padding=0, dilation=1,
                 groups: int = 1, bias: bool = True, padding_mode: str = 'zeros', mode='raw', w_bit=8, a_bit=8,
                 calib_batch_size=32, search_round=1, eq_n=100, fpcs=False, steps=4):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, padding_mode,
                         mode, w_bit, a_bit, calib_batch_size, search_round, eq_n)
        self.fpcs = fpcs
        self.steps = steps
        del self.w_quantizer
        self.w_quantizer = UniformQuantizer(n_bits=w_bit, symmetric=False, channel_wise=True)
        self.w_quantizer.scale = nn.Parameter(torch.zeros((self.out_channels, 1)))
        self.w_quantizer.zero_point = nn.Parameter(torch.zeros((self.out_channels, 1)))

    def _search_best_w_scale(self, weight_scale_candidates, weight_zero_point_candidates, topk=1):
        """reference conv.py:226-263"""
        nl = self.w_quantizer.n_levels
        parts = [sweep.conv_err_w(self._ctx, self._weight2(), self.bias, weight_scale_candidates[p0:p1],
                                  weight_zero_point_candidates[p0:p1], nl)
                 for p0, p1 in _fpcs.candidate_chunks(weight_scale_candidates.shape[0])]
        sims = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        _, best = torch.topk(sims, k=topk, dim=0)
        best = best.view(topk, -1, 1)
        if topk == 1:
            assign(self.w_quantizer.scale, torch.gather(weight_scale_candidates, dim=0, index=best).squeeze(dim=0))
            assign(self.w_quantizer.zero_point, 
                torch.gather(weight_zero_point_candidates, dim=0, index=best).squeeze(dim=0))
        return best

    def calculate_percentile_weight_candidates(self, l=0.9, r=1.0):
        """reference conv.py:271-290 (one zero point per level: num_zp = n_levels)"""
        nl = self.w_quantizer.n_levels
        num_zp = nl
        num_scale = int(self.eq_n / num_zp)
        pct = torch.tensor([l, r])
        w2 = self._weight2()
        up, lo = _fpcs.quantile_pair(w2, pct, -1, local=True)        # weights are replicated, not sharded
        up, lo = up.unsqueeze(-1), lo.unsqueeze(-1)
        return _fpcs.percentile_grid(up[0:1] - lo[0:1], up[1:] - lo[1:], nl, num_zp, num_scale, 0, 2)

    def weight_fpcs(self, fpcs_width=16, steps=4, search_strategy=None):
        """reference conv.py:292-311"""
        cs, cz = self.calculate_percentile_weight_candidates()
        _fpcs.search(cs, cz, lambda s, z, k: search_strategy(self, s, z, topk=k), 0, self.eq_n, fpcs_width, steps)

    def hyperparameter_searching(self):
        """reference conv.py:313-334.  With a_bit >= 8 the input is not quantised and the loop runs one weight search."""
        if self.a_quantizer.n_bits < 8:
            raise NotImplementedError('activation search for convolutions is dead code in the reference '
                                      '(conv.py:329 uses an undefined name); qconv_a_bit must be >= 8')
        self._initialize_calib_parameters()
        cs, cz = self.calculate_percentile_weight_candidates()
        assign(self.w_quantizer.scale, cs[-2])
        assign(self.w_quantizer.zero_point, cz[-2])
        self.w_quantizer.inited = True
        if self.fpcs:
            self.weight_fpcs(steps=self.steps, search_strategy=AsymmetricallyBatchingQuantConv2d._search_best_w_scale)
        else:
            self._search_best_w_scale(cs, cz)
        self.calibrated = True
        del self.raw_input, self.raw_out
        self._ctx = None
