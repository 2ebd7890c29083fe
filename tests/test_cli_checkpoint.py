"""CLI surface and checkpoint interchange: the reference's flags exist with the same defaults, and a checkpoint written
by the UNMODIFIED reference (golden `state_final`) loads into the product's freshly wrapped model through the CLI's
load_model() and reproduces the reference's fake-quant logits bit for bit (CPU, oracle backend for the forwards)."""
import importlib
import os

import pytest
import torch

import _oracle_backend as fake
import test_quant as cli
from conftest import load_golden
from adalog_b200.utils import models as zoo
from adalog_b200.utils.wrap_net import wrap_modules_in_net

REFERENCE_FLAGS = {  # reference test_quant.py:47-80
    '--model': 'deit_small', '--config': './configs/vit_config.py', '--dataset': '/dataset/imagenet/',
    '--val-batch-size': 200, '--num-workers': 8, '--device': 'cuda', '--calibrate': False,
    '--load-calibrate-checkpoint': None, '--test-calibrate-checkpoint': False, '--optimize': False,
    '--load-optimize-checkpoint': None, '--test-optimize-checkpoint': False, '--print-freq': 10, '--seed': 5,
}
SUPPRESSED = ['--calib-size', '--calib-batch-size', '--w_bit', '--a_bit', '--s_bit']


def test_flags_match_reference():
    p = cli.get_args_parser()
    opts = {o: a for a in p._actions for o in a.option_strings}
    for flag, default in REFERENCE_FLAGS.items():
        assert flag in opts, flag
        assert opts[flag].default == default, flag
    args = p.parse_args([])
    for flag in SUPPRESSED:
        assert flag in opts and not hasattr(args, flag.lstrip('-').replace('-', '_'))
    with pytest.raises(SystemExit):
        p.parse_args(['--calibrate', '--load-calibrate-checkpoint', 'x'])      # mutually exclusive, as in the reference


@pytest.mark.parametrize('name', ['model_vit_test_w4a4', 'model_swin_test_w4a4'])
def test_reference_checkpoint_loads(name, monkeypatch, tmp_path):
    g = load_golden(name)
    fake.install(monkeypatch, g['bs'], g['memory'])
    cfg = importlib.import_module(f'adalog_b200.configs.{g["bits"]}bit').Config()
    ckpt = tmp_path / 'ref.pth'
    torch.save(g['state_final'], ckpt)
    model = zoo.create_model(g['model']).eval()
    model = wrap_modules_in_net(model, cfg, reparam=False)
    model = cli.load_model(model, str(ckpt), torch.device('cpu'))
    with torch.no_grad():
        assert torch.equal(model(g['images']), g['logits_final'])
    # and back: our state_dict has the reference's keys, shapes and dtypes
    sd = model.state_dict()
    assert set(sd) == set(g['state_final'])
    for k, v in g['state_final'].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k


def test_checkpoint_name_layout(tmp_path, monkeypatch):
    class A:
        model = 'deit_small'
    cfg = importlib.import_module('adalog_b200.configs.3bit').Config()
    cfg.calib_size = 128
    path = cli.save_model(torch.nn.Linear(2, 2), A, cfg, str(tmp_path))
    assert os.path.basename(path) == 'deit_small_w3_a3_s3_calibsize_128.pth'     # reference test_quant.py:98-100
