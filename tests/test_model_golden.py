"""Model-level host logic on CPU: the product's wrap_net + QuantCalibrator on a tiny ViT and a tiny Swin must walk the
reference's whole calibration (570 / 1140 evaluations) bit for bit when the sweeps are scored by the oracle backend,
and end with an identical checkpoint (state_dict keys, shapes, dtypes, bits) and identical fake-quant logits.
Goldens: the UNMODIFIED reference utils/wrap_net.py + utils/calibrator.py (oracle/make_golden.py models)."""
import hashlib
import types

import pytest
import torch

import _oracle_backend as fake
from conftest import load_golden
from adalog_b200.utils import models as zoo
from adalog_b200.utils.calibrator import QuantCalibrator
from adalog_b200.utils.wrap_net import wrap_modules_in_net, wrap_reparamed_modules_in_net
from adalog_b200.configs import __path__ as cfg_path  # noqa: F401
import importlib


def digest(t):
    return hashlib.sha1(t.detach().contiguous().numpy().tobytes()).hexdigest()


class TopkTap:
    def __init__(self):
        self.evals = []
        self._orig = torch.topk

    def __enter__(self):
        def tapped(inp, k, dim=-1, **kw):
            res = self._orig(inp, k=k, dim=dim, **kw)
            self.evals.append(dict(digest=digest(inp), shape=tuple(inp.shape), k=k, idx=res[1].clone()))
            return res
        torch.topk = tapped
        return self

    def __exit__(self, *a):
        torch.topk = self._orig


def load_cfg(bits, n_img, bs):
    cfg = importlib.import_module(f'adalog_b200.configs.{bits}bit').Config()
    cfg.calib_size, cfg.calib_batch_size = n_img, bs
    return cfg


@pytest.mark.parametrize('name', ['model_vit_test_w4a4', 'model_swin_test_w4a4'])
def test_calibrate_model(name, monkeypatch):
    g = load_golden(name)
    fake.install(monkeypatch, g['bs'], g['memory'])
    cfg = load_cfg(g['bits'], g['n_img'], g['bs'])
    model = zoo.create_model(g['model']).eval()
    model.load_state_dict(g['init_state'])
    images = g['images']
    loader = [(images[i:i + g['bs']], torch.zeros(g['bs'], dtype=torch.long)) for i in range(0, g['n_img'], g['bs'])]
    with torch.no_grad():
        assert torch.equal(model(images), g['fp_logits'])
    model = wrap_modules_in_net(model, cfg, reparam=True)
    assert [n for n, m in model.named_modules() if hasattr(m, 'calibrated')] == g['order']
    calib = QuantCalibrator(model, loader)
    calib.progress = False
    with TopkTap() as tap:
        calib.batching_quant_calib()
    assert len(tap.evals) == len(g['evals'])
    for i, (a, b) in enumerate(zip(tap.evals, g['evals'])):
        assert a['shape'] == b['shape'] and a['k'] == b['k'], i
        assert a['digest'] == b['digest'], f'eval {i}: similarity bits differ'
        assert torch.equal(a['idx'], b['idx'].to(torch.int64)), f'eval {i}: top-k differs'
    model = wrap_reparamed_modules_in_net(model)
    sd = model.state_dict()
    assert set(sd) == set(g['state_calib'])
    for k, v in g['state_calib'].items():
        assert sd[k].dtype == v.dtype and sd[k].shape == v.shape and torch.equal(sd[k], v), k
    with torch.no_grad():
        assert torch.equal(model(images), g['logits'])
    for _, m in model.named_modules():
        if hasattr(m, 'mode') and hasattr(m, 'reparam_bias'):
            m.reparam_bias()
    sd = model.state_dict()
    for k, v in g['state_final'].items():
        assert torch.equal(sd[k], v), k
    with torch.no_grad():
        assert torch.equal(model(images), g['logits_final'])
