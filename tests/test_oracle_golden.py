"""Pins oracle/adalog_oracle.py to the unmodified reference: every recorded search evaluation
(similarity tensor, k, dim, index list incl. exact-tie order) and the final parameters must be
bit-identical on torch-CPU.  Goldens come from oracle/make_golden.py (reference imported from
/root/reference in the build container)."""
import pytest
import torch

import adalog_oracle as O
from conftest import load_golden

LINEAR = ['linear_asym_w4a4', 'linear_asym_w3a3_nv3', 'linear_asym_w6a6_chunked', 'linear_head_2d_w4a4',
          'linear_swin4d_w4a4', 'linear_nobias_w4a4']
CW = ['linear_cw_reparam_w4a4_nv3', 'linear_cw_reparam_w3a3']
GELU = ['linear_postgelu_w4a4', 'linear_postgelu_w3a3', 'linear_postgelu_w6a6']
MATMUL = ['matmul_qk_a4', 'matmul_qk_a3', 'matmul_qk_a6_pooled', 'matmul_pv_s4a4', 'matmul_pv_s3a3', 'matmul_pv_s6a6']
CONV = ['conv_patch_w4', 'conv_patch_w6']


def assert_trace(gold_evals, trace):
    assert len(gold_evals) == len(trace.evals)
    for i, (g, o) in enumerate(zip(gold_evals, trace.evals)):
        assert g['k'] == o['k'] and g['dim'] == o['dim'], f'eval {i}'
        assert g['sims'].shape == o['sims'].shape, f'eval {i} ({o["tag"]})'
        assert torch.equal(g['sims'], o['sims']), f'eval {i} ({o["tag"]}): similarity bits differ'
        assert torch.equal(g['idx'], o['idx']), f'eval {i} ({o["tag"]}): topk index list differs'


def make_linear(g, **kw):
    c = g['cfg']
    return O.LinearSearch(g['weight'].clone(), None if g['bias'] is None else g['bias'].clone(), g['x'].clone(),
                          g['raw_out'].clone(), c['w_bit'], c['a_bit'], n_V=c['n_V'], calib_batch_size=c['bs'],
                          memory=c['memory'], **kw)


@pytest.mark.parametrize('name', LINEAR)
def test_linear_asym(name):
    g = load_golden(name)
    s = make_linear(g)
    s.search_asym()
    if 'chunked' in name:
        assert s.peq == 64
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(st['w_quantizer.scale'], s.wq.scale)
    assert torch.equal(st['w_quantizer.zero_point'], s.wq.zero_point)
    assert torch.equal(st['a_quantizer.scale'], s.aq.scale)
    assert torch.equal(st['a_quantizer.zero_point'], s.aq.zero_point)
    out = torch.nn.functional.linear(s.aq(g['x']), O.quant_weight(s.weight, s.wq, s.n_V), s.bias)
    assert torch.equal(out, g['quant_out'])


@pytest.mark.parametrize('name', CW)
def test_linear_channel_wise_reparam(name):
    g = load_golden(name)
    s = make_linear(g, a_channel_wise=True)
    s.search_channel_wise()
    lw, lb = s.reparam(g['ln_weight'].clone(), g['ln_bias'].clone())
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(lw, g['ln_weight_after']) and torch.equal(lb, g['ln_bias_after'])
    assert torch.equal(st['weight'], s.weight) and torch.equal(st['bias'], s.bias)
    assert torch.equal(st['w_quantizer.scale'], s.wq.scale)
    assert torch.equal(st['w_quantizer.zero_point'], s.wq.zero_point)
    assert torch.equal(st['a_quantizer.scale'], s.aq.scale)
    assert torch.equal(st['a_quantizer.zero_point'], s.aq.zero_point)


@pytest.mark.parametrize('name', GELU)
def test_linear_postgelu(name):
    g = load_golden(name)
    s = make_linear(g, a_kind='adalog')
    s.search_postgelu()
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(st['w_quantizer.scale'], s.wq.scale)
    assert torch.equal(st['w_quantizer.zero_point'], s.wq.zero_point)
    assert torch.equal(st['a_quantizer.scale'], s.aq.scale)
    assert torch.equal(st['a_quantizer.q'], s.aq.q)
    assert torch.equal(st['a_quantizer.table1'], s.aq.table1) and torch.equal(st['a_quantizer.table2'], s.aq.table2)
    out = torch.nn.functional.linear(s.aq(g['x']), O.quant_weight(s.weight, s.wq, s.n_V), s.bias)
    assert torch.equal(out, g['quant_out'])
    s.reparam_bias()
    assert torch.equal(g['state_bias_reparamed']['bias'], s.bias)
    out = torch.nn.functional.linear(s.aq(g['x']), O.quant_weight(s.weight, s.wq, s.n_V), s.bias)
    assert torch.equal(out, g['quant_out_bias_reparamed'])


@pytest.mark.parametrize('name', MATMUL)
def test_matmul(name):
    g = load_golden(name)
    c = g['cfg']
    ps = 'pv' in name
    s = O.MatMulSearch(g['A'].clone(), g['B'].clone(), g['raw_out'].clone(), c['A_bit'], c['B_bit'], c['H'],
                       calib_batch_size=c['bs'], head_channel_wise=c['hcw'], memory=c['memory'], post_softmax=ps)
    s.search()
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(st['B_quantizer.scale'], s.Bq.scale) and torch.equal(st['B_quantizer.zero_point'], s.Bq.zero_point)
    if ps:
        assert torch.equal(st['A_quantizer.q'], s.Aq.q)
        assert torch.equal(st['A_quantizer.table1'], s.Aq.table1) and torch.equal(st['A_quantizer.table2'], s.Aq.table2)
    else:
        assert torch.equal(st['A_quantizer.scale'], s.Aq.scale)
        assert torch.equal(st['A_quantizer.zero_point'], s.Aq.zero_point)
    assert torch.equal(s.Aq(g['A']) @ s.Bq(g['B']), g['quant_out'])


@pytest.mark.parametrize('name', CONV)
def test_conv(name):
    g = load_golden(name)
    c = g['cfg']
    s = O.ConvSearch(g['weight'].clone(), g['bias'].clone(), g['x'].clone(), g['raw_out'].clone(), c['w_bit'], c['k'],
                     calib_batch_size=c['bs'], memory=c['memory'])
    s.search()
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(st['w_quantizer.scale'], s.wq.scale) and torch.equal(st['w_quantizer.zero_point'], s.wq.zero_point)
    oc = s.weight.shape[0]
    w = s.wq(s.weight.view(oc, -1)).view(s.weight.shape)
    assert torch.equal(torch.nn.functional.conv2d(g['x'], w, s.bias, c['k']), g['quant_out'])


GELU_R2 = [('linear_postgelu_nofpcs_w4a4', None), ('linear_postgelu_log2_w4a4', 'log2'),
           ('linear_postgelu_logsqrt2_w3a3', 'logsqrt2')]


@pytest.mark.parametrize('name,tmp_kind', GELU_R2)
def test_linear_postgelu_nondefault(name, tmp_kind):
    """linear.py:985-988 (fpcs=False: base search, scale search :816-854, plain weight search) and :990-994 (fixed-base
    quantizer swapped in after the AdaLog search)"""
    g = load_golden(name)
    s = make_linear(g, a_kind='adalog', fpcs_on=g['cfg']['fpcs'])
    s.search_postgelu(tmp_kind=tmp_kind)
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(st['w_quantizer.scale'], s.wq.scale)
    assert torch.equal(st['w_quantizer.zero_point'], s.wq.zero_point)
    assert torch.equal(st['a_quantizer.scale'], s.aq.scale)
    if tmp_kind is None:
        assert torch.equal(st['a_quantizer.q'], s.aq.q)
    else:
        assert 'a_quantizer.q' not in st
    out = torch.nn.functional.linear(s.aq(g['x']), O.quant_weight(s.weight, s.wq, s.n_V), s.bias)
    assert torch.equal(out, g['quant_out'])
    s.reparam_bias()
    assert torch.equal(g['state_bias_reparamed']['bias'], s.bias)
    out = torch.nn.functional.linear(s.aq(g['x']), O.quant_weight(s.weight, s.wq, s.n_V), s.bias)
    assert torch.equal(out, g['quant_out_bias_reparamed'])


@pytest.mark.parametrize('name', ['linear_twin_w4a4', 'linear_twin_nofpcs_w3a3'])
def test_linear_twin_uniform(name):
    """linear.py:624-721: PTQ4ViT twin-uniform baseline (29 power-of-two scales, argmax selection)"""
    g = load_golden(name)
    s = make_linear(g, a_kind='twin', fpcs_on=g['cfg']['fpcs'])
    s.search_twin()
    assert_trace(g['evals'], s.trace)
    assert [bool(e.get('argmax')) for e in g['evals']] == [bool(e.get('argmax')) for e in s.trace.evals]
    st = g['state']
    assert torch.equal(st['w_quantizer.scale'], s.wq.scale)
    assert torch.equal(st['w_quantizer.zero_point'], s.wq.zero_point)
    assert torch.equal(st['a_quantizer.scale'], s.aq.scale)
    out = torch.nn.functional.linear(s.aq(g['x']), O.quant_weight(s.weight, s.wq, s.n_V), s.bias)
    assert torch.equal(out, g['quant_out'])


@pytest.mark.parametrize('name', ['matmul_pv_log2_s4a4', 'matmul_pv_logsqrt2_s4a4', 'matmul_pv_logsqrt2_s6a6'])
def test_matmul_post_softmax_fixed_base(name):
    """matmul.py:307-310, :367-375: Log2 / LogSqrt2 on the softmax operand, one FPCS pass over B"""
    g = load_golden(name)
    c = g['cfg']
    s = O.MatMulSearch(g['A'].clone(), g['B'].clone(), g['raw_out'].clone(), c['A_bit'], c['B_bit'], c['H'],
                       calib_batch_size=c['bs'], head_channel_wise=c['hcw'], memory=c['memory'], post_softmax=True,
                       quantizer=c['quantizer'])
    s.search()
    assert_trace(g['evals'], s.trace)
    st = g['state']
    assert torch.equal(st['B_quantizer.scale'], s.Bq.scale) and torch.equal(st['B_quantizer.zero_point'], s.Bq.zero_point)
    assert torch.equal(s.Aq(g['A']) @ s.Bq(g['B']), g['quant_out'])


def test_quantizer_forwards():
    g = load_golden('quantizers')
    for c in g['cases']:
        nl = 2 ** (c['bits'] - 1)
        x = c['x']
        if c['q'] == 'uniform':
            y = O.uniform_fakequant(x, c['scale'], c['zero_point'], nl)
        elif c['q'] == 'uniform_sym':
            y = O.uniform_fakequant(x, c['scale'], None, nl, sym=True)
        elif c['q'] == 'adalog':
            t1, t2 = O.adalog_tables(c['qv'], nl)
            assert torch.equal(t1, c['table1']) and torch.equal(t2, c['table2'])
            y = O.adalog_fakequant(x, c['scale'], torch.tensor([c['qv']]), nl)
        elif c['q'] == 'shift_adalog':
            y = O.shift_fakequant(O.adalog_fakequant, x, c['shift'], c['reparamed'], c['scale'],
                                  torch.tensor([c['qv']]), nl)
        elif c['q'] == 'log2':
            y = O.log2_fakequant(x, c['scale'], nl)
        elif c['q'] == 'logsqrt2':
            y = O.logsqrt2_fakequant(x, c['scale'], nl)
        elif c['q'] == 'shift_log2':
            y = O.shift_fakequant(O.log2_fakequant, x, c['shift'], False, c['scale'], nl)
        elif c['q'] == 'shift_logsqrt2':
            y = O.shift_fakequant(O.logsqrt2_fakequant, x, c['shift'], False, c['scale'], nl)
        elif c['q'] == 'twin':
            y = O.twin_uniform_fakequant(x, c['scale'], nl)
        else:
            raise AssertionError(c['q'])
        assert torch.equal(y, c['y']), (c['q'], c['bits'], c.get('tag'), c.get('qv'))
