"""Uniform quantizers with the reference's class API (quantizers/uniform.py); the inference forward
runs the sm_100a kernel adalog_uniform_fakequant_f32, the training_mode (BRECQ/STE) branch stays on
autograd-capable torch ops as north_star prescribes."""
import torch
import torch.nn as nn

from .. import ops
from ._ste import flag, round_ste

__all__ = ['UniformQuantizer', 'ShiftUniformQuantizer', 'TwinUniformQuantizer']


class UniformQuantizer(nn.Module):
    """reference: quantizers/uniform.py:7-39"""

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        super().__init__()
        self.sym = symmetric
        self.n_bits = n_bits
        self.n_levels = 2 ** (self.n_bits - 1)
        self.channel_wise = channel_wise
        self.drop_prob = 1.0
        self.inited = False
        self.training_mode = False
        self.use_clip_forward = False

    def init_training(self):
        self.training_mode = True

    def end_training(self):
        self.training_mode = False

    def _forward_ste(self, x):
        x_int = round_ste(x / self.scale)
        if self.sym:
            return x_int.clamp(-self.n_levels, self.n_levels - 1) * self.scale
        zp = round_ste(self.zero_point)
        return ((x_int + zp).clamp(0, 2 * self.n_levels - 1) - zp) * self.scale

    def forward(self, x):
        if self.n_bits == 32:
            return x
        assert self.inited
        if self.training_mode:
            return self._forward_ste(x)
        return ops.uniform_fakequant(x, self.scale, None if self.sym else self.zero_point, self.n_levels, self.sym)

    def codes(self, x):
        """integer codes of the inference forward (int16), for parity checks"""
        return ops.uniform_fakequant(x, self.scale, None if self.sym else self.zero_point, self.n_levels, self.sym,
                                     want_codes=True, want_y=False)

    def __repr__(self):
        return f'{self.__class__.__name__}(n_bits={self.n_bits}, sym={self.sym}, channel_wise={self.channel_wise})'


class ShiftUniformQuantizer(UniformQuantizer):
    """reference: quantizers/uniform.py:42-50"""

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        super().__init__(n_bits, symmetric, channel_wise)
        self.shift = nn.Parameter(torch.zeros((1)))
        self.register_buffer('bias_reparamed', torch.tensor(False))

    def forward(self, x):
        result = UniformQuantizer.forward(self, x + self.shift)
        return result if flag(self.bias_reparamed) else result - self.shift


class TwinUniformQuantizer(UniformQuantizer):
    """reference: quantizers/uniform.py:53-68"""

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False):
        super().__init__(n_bits, symmetric, channel_wise)

    def forward(self, x):
        if self.n_bits == 32:
            return x
        assert self.inited and self.scale.shape[0] == 2
        if self.training_mode:
            x_pos = round_ste(x / (self.scale[0])).clamp(0, self.n_levels - 1).mul(self.scale[0])
            x_neg = round_ste(x / (self.scale[1])).clamp(-self.n_levels, 0).mul(self.scale[1])
            return (x_pos + x_neg).reshape_as(x)
        if self.scale.numel() != 2:
            raise NotImplementedError('TwinUniformQuantizer kernel takes one positive and one negative scale')
        return ops.twin_fakequant(x, self.scale, self.n_levels)
