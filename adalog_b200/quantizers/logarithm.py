"""Logarithmic quantizers with the reference's class API (quantizers/logarithm.py).  Inference forwards
run the sm_100a kernel adalog_log_fakequant_f32 (one pass: divide, log2, round, LUT dequant, mask, shift);
training_mode (BRECQ/STE) stays on torch ops."""
import math

import torch
import torch.nn as nn

from .. import ops
from ._ste import assign, flag, round_ste

__all__ = ['Log2Quantizer', 'LogSqrt2Quantizer', 'AdaLogQuantizer', 'ShiftLog2Quantizer', 'ShiftLogSqrt2Quantizer',
           'ShiftAdaLogQuantizer']


class Log2Quantizer(nn.Module):
    """reference: quantizers/logarithm.py:8-38"""
    _kind = ops.LOG2
    is_log = True

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        super().__init__()
        self.sym = symmetric
        self.n_bits = n_bits
        self.n_levels = 2 ** (self.n_bits - 1)
        self.inited = False
        self.drop_prob = 1.0
        self.channel_wise = channel_wise
        self.training_mode = False

    def init_training(self):
        self.training_mode = True

    def end_training(self):
        self.training_mode = False

    # ---- training (STE) branch, torch ops: logarithm.py:29-35, :49-54, :88-92
    def _ste_dequant(self, scaled_x):
        x_quant = round_ste(-scaled_x.log2())
        mask = x_quant < 2 * self.n_levels
        x_quant = torch.clamp(x_quant, 0, 2 * self.n_levels - 1)
        return 2 ** (-1 * x_quant) * self.scale * mask

    def _kernel(self, x, shift=None, sub_shift=False, want_codes=False):
        return ops.log_fakequant(x, self.scale, self._kind, self.n_levels, shift=shift, sub_shift=sub_shift,
                                 want_codes=want_codes)

    def _forward(self, x, shift=None, sub_shift=False):
        if self.n_bits == 32:
            # mirrors the reference: Shift* wrappers still add and subtract the shift around the identity
            if shift is None:
                return x
            out = x + shift
            return out - shift if sub_shift else out
        assert self.inited
        if self.training_mode:
            xs = x if shift is None else x + shift
            out = self._ste_dequant((xs / self.scale).clamp(min=1e-15, max=1.0))
            return out - shift if sub_shift else out
        return self._kernel(x, shift, sub_shift)

    def forward(self, x):
        return self._forward(x)

    def codes(self, x, shift=None):
        return self._kernel(x, shift, False, want_codes=True)[1]

    def _log_base(self):
        return 2

    def __repr__(self):
        return (f'{self.__class__.__name__}(n_bits={self.n_bits}, sym={self.sym}, channel_wise={self.channel_wise}, '
                f'log_base={self._log_base()})')


class LogSqrt2Quantizer(Log2Quantizer):
    """reference: quantizers/logarithm.py:41-65"""
    _kind = ops.LOGSQRT2

    def _ste_dequant(self, scaled_x):
        x_quant = round_ste(-scaled_x.log2() * 2)
        mask = x_quant < 2 * self.n_levels
        x_quant = torch.clamp(x_quant, 0, 2 * self.n_levels - 1)
        return 2 ** (-1 * x_quant / 2) * self.scale * mask

    def _log_base(self):
        return math.sqrt(2)


class AdaLogQuantizer(Log2Quantizer):
    """reference: quantizers/logarithm.py:68-102 -- base 2^(-q/37), LUT dequantisation"""
    _kind = ops.ADALOG

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        super().__init__(n_bits, symmetric, channel_wise)
        self.r = 37.0
        self.register_buffer('q', torch.tensor([int(self.r)]))
        self.register_buffer('table1', torch.zeros((self.n_levels * 2)))
        self.register_buffer('table2', torch.zeros((self.n_levels * 2)))
        self.update_table()

    def update_table(self):
        """logarithm.py:77-81: LUTs built in Python float64, stored FP32 (one host sync for q)."""
        q = int(self.q.item())
        n = self.n_levels
        t1 = [math.floor(i * q / self.r) for i in range(2 * n)]
        t2 = [round((2 ** (-((q * i) % self.r) / self.r)) * (4 * n - 2)) / (4 * n - 2) for i in range(2 * n)]
        assign(self.table1, torch.tensor(t1, dtype=torch.float32))
        assign(self.table2, torch.tensor(t2, dtype=torch.float32))

    def _ste_dequant(self, scaled_x):
        x_quant = round_ste(-scaled_x.log2() * self.r / self.q)
        mask = x_quant < 2 * self.n_levels
        x_quant = torch.clamp(x_quant, 0, 2 * self.n_levels - 1)
        return 2 ** (-1 * x_quant * self.q / self.r) * self.scale * mask

    def _kernel(self, x, shift=None, sub_shift=False, want_codes=False):
        return ops.log_fakequant(x, self.scale, self._kind, self.n_levels, self.q, self.table1, self.table2, shift,
                                 sub_shift, want_codes)

    def __repr__(self):
        return (f'{self.__class__.__name__}(n_bits={self.n_bits}, sym={self.sym}, channel_wise={self.channel_wise}, '
                f'q={self.q.item()})')


class _ShiftMixin:
    """reference: quantizers/logarithm.py:105-135 -- Q(x + shift) [- shift unless folded into the bias]"""

    def _init_shift(self):
        self.shift = nn.Parameter(torch.zeros((1)))
        self.register_buffer('bias_reparamed', torch.tensor(False))

    def forward(self, x):
        return self._forward(x, self.shift, not flag(self.bias_reparamed))

    def codes(self, x):
        return self._kernel(x, self.shift, False, want_codes=True)[1]


class ShiftLog2Quantizer(_ShiftMixin, Log2Quantizer):
    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        Log2Quantizer.__init__(self, n_bits, symmetric, channel_wise)
        self._init_shift()


class ShiftLogSqrt2Quantizer(_ShiftMixin, LogSqrt2Quantizer):
    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        LogSqrt2Quantizer.__init__(self, n_bits, symmetric, channel_wise)
        self._init_shift()


class ShiftAdaLogQuantizer(_ShiftMixin, AdaLogQuantizer):
    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        AdaLogQuantizer.__init__(self, n_bits, symmetric, channel_wise)
        self._init_shift()
