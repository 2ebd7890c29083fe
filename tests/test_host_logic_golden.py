"""Host-side logic of the product (candidate seeding, FPCS refinement, selection, reparam, state_dict layout)
against the reference goldens, on CPU.  The CUDA sweeps are replaced by the oracle (tests/_oracle_backend.py), so a
bit-exact match here means: given the reference's similarity values, the product walks exactly the reference's
search trajectory and stores exactly its parameters."""
import pytest
import torch

import _oracle_backend as fake
from conftest import load_golden
from adalog_b200 import quant_layers as QL

LINEAR = ['linear_asym_w4a4', 'linear_asym_w3a3_nv3', 'linear_asym_w6a6_chunked', 'linear_head_2d_w4a4',
          'linear_swin4d_w4a4', 'linear_nobias_w4a4']
CW = ['linear_cw_reparam_w4a4_nv3', 'linear_cw_reparam_w3a3']
GELU = ['linear_postgelu_w4a4', 'linear_postgelu_w3a3', 'linear_postgelu_w6a6']
MATMUL = ['matmul_qk_a4', 'matmul_qk_a3', 'matmul_qk_a6_pooled', 'matmul_pv_s4a4', 'matmul_pv_s3a3', 'matmul_pv_s6a6']
CONV = ['conv_patch_w4', 'conv_patch_w6']


class TopkTap:
    def __init__(self, argmax=False):
        self.evals = []
        self._orig = torch.topk
        self._orig_argmax = torch.Tensor.argmax
        self._argmax = argmax

    def __enter__(self):
        def tapped(inp, k, dim=-1, **kw):
            res = self._orig(inp, k=k, dim=dim, **kw)
            self.evals.append(dict(sims=inp.detach().clone(), k=k, dim=dim, idx=res[1].clone()))
            return res
        torch.topk = tapped
        if self._argmax:                     # the twin-uniform search selects with Tensor.argmax (linear.py:691)
            orig = self._orig_argmax

            def tapped_argmax(t, *a, **kw):
                res = orig(t, *a, **kw)
                self.evals.append(dict(sims=t.detach().clone(), k=1, dim=kw.get('dim', a[0] if a else None),
                                       idx=res.clone(), argmax=True))
                return res
            torch.Tensor.argmax = tapped_argmax
        return self

    def __exit__(self, *a):
        torch.topk = self._orig
        torch.Tensor.argmax = self._orig_argmax


def assert_trace(gold, got):
    assert len(gold) == len(got)
    for i, (g, o) in enumerate(zip(gold, got)):
        assert g['k'] == o['k'] and g['dim'] % g['sims'].dim() == o['dim'] % o['sims'].dim(), f'eval {i}'
        assert torch.equal(g['sims'], o['sims']), f'eval {i}: similarity bits differ'
        assert torch.equal(g['idx'], o['idx']), f'eval {i}: topk index list differs'


def assert_state(gold_state, module):
    sd = module.state_dict()
    assert set(gold_state.keys()) == set(sd.keys())
    for k, v in gold_state.items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
        assert torch.equal(sd[k], v), k


def build_linear(g, cls, **extra):
    c = g['cfg']
    m = cls(c['in_f'], c['out_f'], bias=c['bias'], w_bit=c['w_bit'], a_bit=c['a_bit'], calib_batch_size=c['bs'],
            eq_n=128, fpcs=c.get('fpcs', True), steps=6, search_round=3, n_V=c['n_V'], **extra)
    m.weight.data.copy_(g['weight'])
    if c['bias']:
        m.bias.data.copy_(g['bias'])
    return m


@pytest.mark.parametrize('name', LINEAR)
def test_asym_linear(name, monkeypatch):
    g = load_golden(name)
    fake.install(monkeypatch, g['cfg']['bs'], g['cfg']['memory'])
    m = build_linear(g, QL.AsymmetricallyBatchingQuantLinear)
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = g['x'].clone(), g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    assert m.calibrated and not hasattr(m, 'raw_input')
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['x']), g['quant_out'])


@pytest.mark.parametrize('name', CW)
def test_channel_wise_reparam(name, monkeypatch):
    g = load_golden(name)
    fake.install(monkeypatch, g['cfg']['bs'], g['cfg']['memory'])
    m = build_linear(g, QL.AsymmetricallyChannelWiseBatchingQuantLinear)
    ln = torch.nn.LayerNorm(g['cfg']['in_f'])
    ln.weight.data.copy_(g['ln_weight'])
    ln.bias.data.copy_(g['ln_bias'])
    m.prev_layer = ln
    assert 'prev_layer' not in dict(m.named_modules()) and not any('prev' in k for k in m.state_dict())
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = g['x'].clone(), g['raw_out'].clone()
        m.hyperparameter_searching()
        assert m.prev_layer is not None
        m.reparam()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    assert torch.equal(ln.weight, g['ln_weight_after']) and torch.equal(ln.bias, g['ln_bias_after'])


@pytest.mark.parametrize('name', GELU)
def test_postgelu(name, monkeypatch):
    g = load_golden(name)
    fake.install(monkeypatch, g['cfg']['bs'], g['cfg']['memory'])
    m = build_linear(g, QL.PostGeluLogBasedBatchingQuantLinear, quantizer='adalog')
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = g['x'].clone(), g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['x']), g['quant_out'])
        m.reparam_bias()
        assert_state(g['state_bias_reparamed'], m)
        assert torch.equal(m(g['x']), g['quant_out_bias_reparamed'])


@pytest.mark.parametrize('name', ['linear_postgelu_nofpcs_w4a4', 'linear_postgelu_log2_w4a4',
                                  'linear_postgelu_logsqrt2_w3a3'])
def test_postgelu_nondefault(name, monkeypatch):
    """fpcs=False driver (linear.py:985-988 with the scale-only search :816-854) and the fixed-base quantizers swapped
    in after the AdaLog search (:990-994)"""
    g = load_golden(name)
    fake.install(monkeypatch, g['cfg']['bs'], g['cfg']['memory'])
    m = build_linear(g, QL.PostGeluLogBasedBatchingQuantLinear, quantizer=g['cfg']['quantizer'])
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = g['x'].clone(), g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['x']), g['quant_out'])
        m.reparam_bias()
        assert_state(g['state_bias_reparamed'], m)
        assert torch.equal(m(g['x']), g['quant_out_bias_reparamed'])


@pytest.mark.parametrize('name', ['linear_twin_w4a4', 'linear_twin_nofpcs_w3a3'])
def test_twin_uniform(name, monkeypatch):
    """PTQ4ViT twin-uniform baseline (linear.py:624-721; post_gelu_quantizer='ptq4vit')"""
    g = load_golden(name)
    fake.install(monkeypatch, g['cfg']['bs'], g['cfg']['memory'])
    m = build_linear(g, QL.PostGeluTwinUniformBatchingQuantLinear)
    with torch.no_grad(), TopkTap(argmax=True) as tap:
        m.raw_input, m.raw_out = g['x'].clone(), g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert [bool(e.get('argmax')) for e in g['evals']] == [bool(e.get('argmax')) for e in tap.evals]
    assert_state(g['state'], m)
    assert m.calibrated and not hasattr(m, 'raw_input')
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['x']), g['quant_out'])


@pytest.mark.parametrize('name', ['matmul_pv_log2_s4a4', 'matmul_pv_logsqrt2_s4a4', 'matmul_pv_logsqrt2_s6a6'])
def test_matmul_post_softmax_fixed_base(name, monkeypatch):
    """post_softmax_quantizer 'log2' / 'logsqrt2' (matmul.py:307-310, :367-375)"""
    g = load_golden(name)
    c = g['cfg']
    fake.install(monkeypatch, c['bs'], c['memory'])
    m = QL.PostSoftmaxAsymmetricallyBatchingQuantMatMul(
        A_bit=c['A_bit'], B_bit=c['B_bit'], calib_batch_size=c['bs'], search_round=3, eq_n=128,
        head_channel_wise=c['hcw'], num_heads=c['H'], fpcs=True, steps=6, quantizer=c['quantizer'])
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = [g['A'].clone(), g['B'].clone()], g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['A'], g['B']), g['quant_out'])


@pytest.mark.parametrize('name', MATMUL)
def test_matmul(name, monkeypatch):
    g = load_golden(name)
    c = g['cfg']
    fake.install(monkeypatch, c['bs'], c['memory'])
    kw = dict(A_bit=c['A_bit'], B_bit=c['B_bit'], calib_batch_size=c['bs'], search_round=3, eq_n=128,
              head_channel_wise=c['hcw'], num_heads=c['H'], fpcs=True, steps=6)
    if 'pv' in name:
        m = QL.PostSoftmaxAsymmetricallyBatchingQuantMatMul(quantizer='adalog', **kw)
    else:
        m = QL.AsymmetricallyBatchingQuantMatMul(**kw)
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = [g['A'].clone(), g['B'].clone()], g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['A'], g['B']), g['quant_out'])


@pytest.mark.parametrize('name', CONV)
def test_conv(name, monkeypatch):
    g = load_golden(name)
    c = g['cfg']
    fake.install(monkeypatch, c['bs'], c['memory'])
    m = QL.AsymmetricallyBatchingQuantConv2d(c['ic'], c['oc'], c['k'], stride=c['k'], w_bit=c['w_bit'], a_bit=8,
                                             calib_batch_size=c['bs'], search_round=3, eq_n=128, fpcs=True, steps=6)
    m.weight.data.copy_(g['weight'])
    m.bias.data.copy_(g['bias'])
    with torch.no_grad(), TopkTap() as tap:
        m.raw_input, m.raw_out = g['x'].clone(), g['raw_out'].clone()
        m.hyperparameter_searching()
    assert_trace(g['evals'], tap.evals)
    assert_state(g['state'], m)
    m.mode = 'quant_forward'
    with torch.no_grad():
        assert torch.equal(m(g['x']), g['quant_out'])


def test_no_cpu_fallback():
    """without the test backend the product refuses to run on CPU tensors"""
    from adalog_b200._lib import AdalogError
    from adalog_b200.quantizers import UniformQuantizer
    q = UniformQuantizer(4)
    q.scale = torch.nn.Parameter(torch.tensor([0.1]))
    q.zero_point = torch.nn.Parameter(torch.tensor([3.0]))
    q.inited = True
    with pytest.raises(AdalogError):
        q(torch.randn(8))
    m = QL.AsymmetricallyBatchingQuantLinear(8, 8, w_bit=4, a_bit=4, eq_n=128, fpcs=True, steps=6)
    m.raw_input, m.raw_out = torch.randn(2, 3, 8), torch.randn(2, 3, 8)
    with pytest.raises(EnvironmentError):
        m.hyperparameter_searching()
