"""Model surgery with the reference's API (utils/wrap_net.py): swap nn.Linear / nn.Conv2d / the two attention matmuls
for the quant modules of adalog_b200.quant_layers.  The dispatch rules are the reference's;
the implementation is table-driven instead of one long if-chain."""
from types import MethodType

import torch
from torch import nn

from ..quant_layers.conv import AsymmetricallyBatchingQuantConv2d
from ..quant_layers.linear import (AsymmetricallyBatchingQuantLinear, AsymmetricallyChannelWiseBatchingQuantLinear,
                                   PostGeluLogBasedBatchingQuantLinear, PostGeluTwinUniformBatchingQuantLinear)
from ..quant_layers.matmul import AsymmetricallyBatchingQuantMatMul, PostSoftmaxAsymmetricallyBatchingQuantMatMul
from . import models

try:  # a real timm, when present, is honoured as well
    from timm.models.vision_transformer import Attention as _TimmAttention
    from timm.models.swin_transformer import WindowAttention as _TimmWindowAttention
    _VIT_ATTN, _SWIN_ATTN = (models.Attention, _TimmAttention), (models.WindowAttention, _TimmWindowAttention)
except Exception:  # noqa: BLE001
    _VIT_ATTN, _SWIN_ATTN = (models.Attention,), (models.WindowAttention,)


class MatMul(nn.Module):
    """placeholder that gives the attention matmuls a module identity (reference wrap_net.py:14-16)"""

    def forward(self, A, B):
        return A @ B


def vit_attn_forward(self, x):
    """reference wrap_net.py:19-32: scale applied AFTER matmul1"""
    B, N, C = x.shape
    qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q, k = self.q_norm(q), self.k_norm(k)
    attn = self.matmul1(q, k.transpose(-2, -1)) * self.scale
    attn = self.attn_drop(attn.softmax(dim=-1))
    x = self.matmul2(attn, v).transpose(1, 2).reshape(B, N, C)
    return self.proj_drop(self.proj(x))


def swin_attn_forward(self, x, mask=None):
    """reference wrap_net.py:35-52: q pre-scaled BEFORE matmul1, relative position bias and shift mask added after"""
    B_, N, C = x.shape
    qkv = self.qkv(x).reshape(B_, N, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = self.matmul1(q * self.scale, k.transpose(-2, -1)) + self._get_rel_pos_bias()
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(-1, nW, self.num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, self.num_heads, N, N)
    attn = self.attn_drop(attn.softmax(dim=-1))
    x = self.matmul2(attn, v).transpose(1, 2).reshape(B_, N, C)
    return self.proj_drop(self.proj(x))


def _parent_of(lookup, name):
    head, _, leaf = name.rpartition('.')
    if head not in lookup:
        raise RuntimeError(f"father module {head} not found")
    return lookup[head], head, leaf


def _search_kwargs(cfg):
    return dict(mode='raw', calib_batch_size=cfg.calib_batch_size, search_round=cfg.search_round, eq_n=cfg.eq_n,
                fpcs=cfg.fpcs, steps=cfg.steps)


def _make_conv(module, cfg):
    new = AsymmetricallyBatchingQuantConv2d(in_channels=module.in_channels, out_channels=module.out_channels,
                                            kernel_size=module.kernel_size, stride=module.stride, w_bit=cfg.w_bit,
                                            a_bit=cfg.qconv_a_bit, **_search_kwargs(cfg))
    new.weight.data.copy_(module.weight.data)
    new.bias.data.copy_(module.bias.data)
    return new


def _make_matmul(name, parent, cfg):
    kw = dict(B_bit=cfg.a_bit, head_channel_wise=cfg.matmul_head_channel_wise, num_heads=parent.num_heads,
              **_search_kwargs(cfg))
    if 'matmul2' in name:
        return PostSoftmaxAsymmetricallyBatchingQuantMatMul(A_bit=cfg.s_bit, quantizer=cfg.post_softmax_quantizer, **kw)
    return AsymmetricallyBatchingQuantMatMul(A_bit=cfg.a_bit, **kw)


def _make_linear(name, module, parent, grandparent, cfg, reparam):
    a_bit = cfg.qhead_a_bit if 'head' in name else cfg.a_bit
    kw = dict(in_features=module.in_features, out_features=module.out_features, bias=module.bias is not None,
              w_bit=cfg.w_bit, a_bit=a_bit, n_V=3 if 'qkv' in name else 1, **_search_kwargs(cfg))
    if a_bit == cfg.w_bit and reparam and any(t in name for t in ('qkv', 'reduction', 'fc1')):
        new = AsymmetricallyChannelWiseBatchingQuantLinear(**kw)
        if 'qkv' in name:
            new.prev_layer = grandparent.norm1
        if 'fc1' in name:
            new.prev_layer = grandparent.norm2
        if 'reduction' in name:
            new.prev_layer = parent.norm
    elif 'fc2' in name and cfg.post_gelu_quantizer in ('adalog', 'log2', 'logsqrt2'):
        new = PostGeluLogBasedBatchingQuantLinear(quantizer=cfg.post_gelu_quantizer, **kw)
    elif 'fc2' in name and cfg.post_gelu_quantizer == 'ptq4vit':
        new = PostGeluTwinUniformBatchingQuantLinear(**kw)
    else:
        new = AsymmetricallyBatchingQuantLinear(**kw)
    new.weight.data.copy_(module.weight.data)
    if module.bias is not None:
        new.bias.data.copy_(module.bias.data)
    return new


def wrap_modules_in_net(model, cfg, reparam=False):
    """reference wrap_net.py:55-172"""
    for _, module in model.named_modules():
        if isinstance(module, _VIT_ATTN) or isinstance(module, _SWIN_ATTN):
            setattr(module, "matmul1", MatMul())
            setattr(module, "matmul2", MatMul())
            fwd = vit_attn_forward if isinstance(module, _VIT_ATTN) else swin_attn_forward
            module.forward = MethodType(fwd, module)
    device = next(model.parameters()).device
    lookup = {}
    for name, module in model.named_modules():
        lookup[name] = module
        if name == '':
            continue
        parent, parent_name, leaf = _parent_of(lookup, name)
        new = None
        if isinstance(module, nn.Conv2d):
            new = _make_conv(module, cfg)
        elif isinstance(module, MatMul):
            new = _make_matmul(name, parent, cfg)
        elif isinstance(module, nn.Linear):
            grandparent = lookup.get(parent_name.rpartition('.')[0])
            new = _make_linear(name, module, parent, grandparent, cfg, reparam)
        if new is not None:
            new.to(device)
            setattr(parent, leaf, new)
    return model


def wrap_reparamed_modules_in_net(model):
    """reference wrap_net.py:175-210: after calibration a channel-wise linear is an ordinary asymmetric one"""
    lookup = {}
    for name, module in model.named_modules():
        lookup[name] = module
        if name == '' or not isinstance(module, AsymmetricallyChannelWiseBatchingQuantLinear):
            continue
        parent, _, leaf = _parent_of(lookup, name)
        new = AsymmetricallyBatchingQuantLinear(
            in_features=module.in_features, out_features=module.out_features, bias=module.bias is not None,
            mode=module.mode, w_bit=module.w_quantizer.n_bits, a_bit=module.a_quantizer.n_bits,
            calib_batch_size=module.calib_batch_size, search_round=module.search_round, eq_n=module.eq_n, n_V=module.n_V,
            fpcs=module.fpcs, steps=module.steps)
        new.load_state_dict(module.state_dict())
        new.calibrated = True
        new.a_quantizer.inited = True
        new.w_quantizer.inited = True
        new.to(module.weight.device)
        setattr(parent, leaf, new)
    return model


def set_tensor_core_forward(model, enabled=True):
    """Switch the linear layers' `quant_forward` between the default (bit-identical to the reference's
    F.linear(Q_a(x), Q_w(W), b)) and the exact-integer tensor-core GEMM with a dequantising epilogue
    (quant_layers/linear.py: PTQSLQuantLinear.quant_forward).  For evaluation / serving after calibration."""
    n = 0
    for m in model.modules():
        if hasattr(m, 'w_quantizer') and hasattr(m, 'in_features'):
            m.tc_forward = bool(enabled)
            n += 1
    return n
