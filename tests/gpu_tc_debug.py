"""GPU helper (not a pytest file): per-module difference between the default and the tensor-core inference forward in
a calibrated model."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from adalog_b200.utils.wrap_net import wrap_reparamed_modules_in_net  # noqa: E402

name, bits, n_img = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = importlib.import_module(f'adalog_b200.configs.{bits}bit').Config()
cfg.calib_size, cfg.calib_batch_size = n_img, 32
dev = torch.device('cuda', 0)
model = bench.build_wrapped(bench.MODEL_ALIASES[name], cfg, dev)
images = torch.randn(n_img, 3, 224, 224, generator=torch.Generator().manual_seed(5)).to(dev)
bench.calibrate(model, images, 32)
model = wrap_reparamed_modules_in_net(model)
rows = []


def hook(nm):
    def f(m, inp, out):
        x = inp[0]
        with torch.no_grad():
            m.tc_forward = False
            y0 = m.quant_forward(x)
            m.tc_forward = True
            y1 = m.quant_forward(x)
            m.tc_forward = False
        rows.append((nm, type(m).__name__, tuple(x.shape), float((y1 - y0).abs().max() / y0.abs().max()),
                     type(m.a_quantizer).__name__, bool(getattr(m.a_quantizer, 'bias_reparamed', False))))
    return f


hs = [m.register_forward_hook(hook(n)) for n, m in model.named_modules() if hasattr(m, 'w_quantizer') and hasattr(m, 'in_features')]
with torch.no_grad():
    model(images[:32])
for r in rows[:12] + rows[-3:]:
    print(r)

# sensitivity of the network itself: the DEFAULT forward on inputs perturbed at the 1e-7 level vs unperturbed
for h in hs:
    h.remove()
from adalog_b200.utils.wrap_net import set_tensor_core_forward  # noqa: E402
with torch.no_grad():
    ref = model(images[:n_img])
    pert = model(images[:n_img] * (1 + 1e-7 * torch.randn_like(images[:n_img])))
    set_tensor_core_forward(model, True)
    tc = model(images[:n_img])
    set_tensor_core_forward(model, False)
agree = lambda a, b: float((a.argmax(-1) == b.argmax(-1)).float().mean())
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
print(f'default vs default on 1e-7-perturbed images: top-1 agreement {agree(pert, ref):.3f}, max rel logit diff {rel(pert, ref):.3f}')
print(f'tensor-core vs default, same images:          top-1 agreement {agree(tc, ref):.3f}, max rel logit diff {rel(tc, ref):.3f}')
