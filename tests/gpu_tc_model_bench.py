"""Whole-model timing of the default vs the opt-in tensor-core inference forward, rep by rep (not a pytest file).
python tests/gpu_tc_model_bench.py [model bits images]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from adalog_b200.utils.wrap_net import set_tensor_core_forward, wrap_reparamed_modules_in_net  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'deit_small'
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_img = int(sys.argv[3]) if len(sys.argv) > 3 else 128
cfg = importlib.import_module(f'adalog_b200.configs.{bits}bit').Config()
cfg.calib_size, cfg.calib_batch_size = 32, 32
dev = torch.device('cuda', 0)
model = bench.build_wrapped(bench.MODEL_ALIASES[name], cfg, dev)
images = torch.randn(n_img, 3, 224, 224, generator=torch.Generator().manual_seed(5)).to(dev)
bench.calibrate(model, images[:32], 32)
model = wrap_reparamed_modules_in_net(model)


def reps(bs, n=4):
    out = []
    with torch.no_grad():
        for _ in range(n):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(0, n_img, bs):
                model(images[i:i + bs])
            e1.record()
            torch.cuda.synchronize()
            out.append(round(e0.elapsed_time(e1), 2))
    return out


for tc in (False, True, False, True):
    set_tensor_core_forward(model, tc)
    print(f'tensor_core={tc}: batch 32 ms per {n_img} images {reps(32)}   batch {n_img}: {reps(n_img)}', flush=True)
