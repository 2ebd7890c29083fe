#!/usr/bin/env python
"""Command-line entry with the reference's flags and calibrate-checkpoint layout (reference test_quant.py).

  python test_quant.py --model deit_small --config adalog_b200/configs/3bit.py --calibrate --dataset synthetic

Differences forced by the environment (no timm, no ImageNet, no network): models come from the timm-free zoo with
random-init weights (or `./checkpoints/vit_raw/<name>.bin` when present), and `--dataset synthetic` (default when the
given path does not exist) uses seeded N(0,1) images; "validation" then reports top-1 AGREEMENT of the fake-quant model
with the FP32 model on a synthetic batch instead of ImageNet accuracy.  --optimize (BRECQ) is out of scope.
"""
import argparse
import copy
import importlib.util
import logging
import os
import sys
import time
from datetime import datetime

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from adalog_b200.utils import models as zoo  # noqa: E402
from adalog_b200.utils.calibrator import QuantCalibrator  # noqa: E402
from adalog_b200.utils.wrap_net import wrap_modules_in_net, wrap_reparamed_modules_in_net  # noqa: E402

MODEL_ZOO = {  # reference test_quant.py:162-176
    'vit_tiny': 'vit_tiny_patch16_224', 'vit_small': 'vit_small_patch16_224', 'vit_base': 'vit_base_patch16_224',
    'vit_large': 'vit_large_patch16_224', 'deit_tiny': 'deit_tiny_patch16_224', 'deit_small': 'deit_small_patch16_224',
    'deit_base': 'deit_base_patch16_224', 'swin_tiny': 'swin_tiny_patch4_window7_224',
    'swin_small': 'swin_small_patch4_window7_224', 'swin_base': 'swin_base_patch4_window7_224',
    'swin_base_384': 'swin_base_patch4_window12_384',
}


def get_args_parser():
    """reference test_quant.py:45-81 (same flags, same defaults)"""
    parser = argparse.ArgumentParser(add_help=False)
    parser.add_argument("--model", default="deit_small", choices=sorted(MODEL_ZOO), help="model")
    parser.add_argument('--config', type=str, default="./configs/vit_config.py",
                        help="File path to import Config class from")
    parser.add_argument('--dataset', default="/dataset/imagenet/", help='path to dataset, or "synthetic"')
    parser.add_argument("--calib-size", default=argparse.SUPPRESS, type=int, help="size of calibration set")
    parser.add_argument("--calib-batch-size", default=argparse.SUPPRESS, type=int, help="batchsize of calibration set")
    parser.add_argument("--val-batch-size", default=200, type=int, help="batchsize of validation set")
    parser.add_argument("--num-workers", default=8, type=int, help="number of data loading workers (default: 8)")
    parser.add_argument("--device", default="cuda", type=str, help="device")
    mode = parser.add_mutually_exclusive_group()
    mode.add_argument('--calibrate', action='store_true', help="Calibrate the model")
    mode.add_argument('--load-calibrate-checkpoint', type=str, default=None, help="Path to the calibrated checkpoint.")
    parser.add_argument('--test-calibrate-checkpoint', action='store_true', help='validate the calibrated checkpoint.')
    opt = parser.add_mutually_exclusive_group()
    opt.add_argument('--optimize', action='store_true', help="Optimize the model (BRECQ; not part of adalog_b200)")
    opt.add_argument('--load-optimize-checkpoint', type=str, default=None, help="Path to the optimized checkpoint.")
    parser.add_argument('--test-optimize-checkpoint', action='store_true', help='validate the optimized checkpoint.')
    parser.add_argument("--print-freq", default=10, type=int, help="print frequency")
    parser.add_argument("--seed", default=5, type=int, help="seed")
    parser.add_argument('--w_bit', type=int, default=argparse.SUPPRESS, help='bit-precision of weights')
    parser.add_argument('--a_bit', type=int, default=argparse.SUPPRESS, help='bit-precision of activation')
    parser.add_argument('--s_bit', type=int, default=argparse.SUPPRESS, help='bit-precision of post softmax activation')
    return parser


def seed_all(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def load_config(path):
    """reference test_quant.py:139-147: import `Config` from a file path"""
    spec = importlib.util.spec_from_file_location(os.path.splitext(os.path.basename(path))[0], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Config()


def make_root():
    """reference test_quant.py:20-29"""
    while True:
        try:
            root = './checkpoints/quant_result/{}'.format(datetime.now().strftime("%Y%m%d_%H%M"))
            os.makedirs(root)
            return root
        except FileExistsError:
            time.sleep(10)


def save_model(model, args, cfg, root_path, mode='calibrate'):
    """reference test_quant.py:95-106: same file name and state_dict layout"""
    assert mode in ['calibrate', 'optimize']
    size = cfg.calib_size if mode == 'calibrate' else cfg.optim_size
    tag = 'calibsize' if mode == 'calibrate' else 'optimsize'
    path = os.path.join(root_path, '{}_w{}_a{}_s{}_{}_{}.pth'.format(args.model, cfg.w_bit, cfg.a_bit, cfg.s_bit, tag, size))
    logging.info(f"Saving checkpoint to {path}")
    torch.save(model.state_dict(), path)
    return path


def load_model(model, ckpt_path, device):
    """reference test_quant.py:109-127"""
    for name, module in model.named_modules():
        if hasattr(module, 'mode'):
            module.calibrated = True
            module.mode = 'quant_forward'
        if isinstance(module, nn.Linear) and 'reduction' in name:
            module.bias = nn.Parameter(torch.zeros(module.out_features))
        for attr in ('a_quantizer', 'w_quantizer', 'A_quantizer', 'B_quantizer'):
            if hasattr(module, attr):
                getattr(module, attr).inited = True
    result = model.load_state_dict(torch.load(ckpt_path, map_location='cpu'), strict=False)
    logging.info(str(result))
    model.to(device)
    model.eval()
    return model


def finish_training(model):
    """reference test_quant.py:130-133"""
    for _, module in model.named_modules():
        if hasattr(module, 'mode') and hasattr(module, 'reparam_bias'):
            module.reparam_bias()


def synthetic_loader(n, batch_size, img, seed, device):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, img, img, generator=g)
    return [(x[i:i + batch_size].to(device), torch.zeros(min(batch_size, n - i), dtype=torch.long))
            for i in range(0, n, batch_size)]


@torch.no_grad()
def validate_agreement(loader, model, full_model):
    """stand-in for utils/test_utils.py:validate when no labelled data exists: top-1 agreement with the FP32 model"""
    agree = total = 0
    t0 = time.time()
    for x, _ in loader:
        agree += (model(x).argmax(-1) == full_model(x).argmax(-1)).sum().item()
        total += x.shape[0]
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dt = time.time() - t0
    logging.info(f' * top-1 agreement with FP32 model {100.0 * agree / total:.2f}% on {total} images ({total / dt:.0f} img/s incl. FP32 model)')
    return agree / total


def main(args):
    root_path = make_root()
    logging.basicConfig(level=logging.INFO, format='%(message)s',
                        handlers=[logging.FileHandler('{}/output.log'.format(root_path)), logging.StreamHandler()])
    logging.info(str(args))
    cfg = load_config(args.config)
    for k in ('calib_size', 'calib_batch_size', 'w_bit', 'a_bit', 's_bit'):
        if hasattr(args, k):
            setattr(cfg, k, getattr(args, k))
    for name, value in vars(cfg).items():
        logging.info(f"{name}: {value}")
    if args.optimize or args.load_optimize_checkpoint or args.test_optimize_checkpoint:
        raise NotImplementedError('BRECQ block reconstruction stays on the reference PyTorch path (out of scope)')
    device = torch.device(args.device)
    seed_all(args.seed)
    logging.info('Building model ...')
    name = MODEL_ZOO[args.model]
    raw = './checkpoints/vit_raw/{}.bin'.format(name)
    model = zoo.create_model(name, checkpoint_path=raw if os.path.exists(raw) else None)
    full_model = copy.deepcopy(model).to(device).eval()
    model.to(device).eval()
    img = model.default_cfg['input_size'][-1]
    if args.dataset != 'synthetic' and not os.path.isdir(args.dataset):
        logging.info(f"dataset path {args.dataset} not found: using synthetic images")
    val_loader = synthetic_loader(args.val_batch_size, args.val_batch_size, img, args.seed + 1, device)
    reparam = args.load_calibrate_checkpoint is None
    logging.info('Wraping quantiztion modules (reparam: {}) ...'.format(reparam))
    model = wrap_modules_in_net(model, cfg, reparam=reparam).to(device).eval()
    if args.load_calibrate_checkpoint:
        model = load_model(model, args.load_calibrate_checkpoint, device)
        if args.test_calibrate_checkpoint:
            validate_agreement(val_loader, model, full_model)
        return model
    t0 = time.time()
    calib_loader = synthetic_loader(cfg.calib_size, cfg.calib_batch_size, img, args.seed, device)
    QuantCalibrator(model, calib_loader).batching_quant_calib()
    model = wrap_reparamed_modules_in_net(model).to(device)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    logging.info("calibration finished in {:.1f} s".format(time.time() - t0))
    finish_training(model)
    save_model(model, args, cfg, root_path, mode='calibrate')
    logging.info('Validating after calibration ...')
    validate_agreement(val_loader, model, full_model)
    return model


if __name__ == "__main__":
    parser = argparse.ArgumentParser(parents=[get_args_parser()])
    main(parser.parse_args())
