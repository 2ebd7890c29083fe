"""GPU helper (not a pytest file): achieved HBM GB/s of the memory-bound fake-quant kernels vs MEASURED_PEAKS.json.
Algorithmic bytes = 4 B read + 4 B write per element; CUDA events, 3 warm-ups, tensors > L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from adalog_b200 import ops  # noqa: E402
import adalog_oracle as O  # noqa: E402

DEV = 'cuda'
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {}
from adalog_b200._lib import call  # noqa: E402
P = ops._p
ST = ops._stream


def abi_uniform(xt, yt, sc, zr, inner, ngroups, nl):
    call('adalog_uniform_fakequant_f32', P(xt), P(yt), None, xt.numel(), P(sc), P(zr), inner, ngroups, nl, 0, ST())


def abi_log(xt, yt, sc, qq, t1, t2, sh, sub, nl):
    call('adalog_log_fakequant_f32', P(xt), P(yt), None, xt.numel(), P(sc), 2, nl, P(qq), P(t1), P(t2), P(sh), sub, ST())


# kernel-level numbers: the C-ABI entry point on preallocated buffers (what the roofline is about); the Python
# wrappers add three tiny torch kernels for round_ste(zero_point) per call, reported separately below
x = torch.randn(128, 197, 3072, device=DEV)            # 310 MB: post-GELU-sized activation, > L2
n = x.numel()
yb = torch.empty_like(x)
s1, z1 = torch.tensor([0.05], device=DEV), torch.tensor([7.0], device=DEV)
ms = timed(lambda: abi_uniform(x, yb, s1, z1, n, 1, 8))
res['uniform_fakequant per-tensor'] = 8 * n / ms / 1e6
ms = timed(lambda: ops.uniform_fakequant(x, s1, z1, 8))
res['uniform_fakequant per-tensor (python wrapper)'] = 8 * n / ms / 1e6
w = torch.randn(3, 4096, 4096, device=DEV)             # 201 MB weight-like, per-row scale
wb = torch.empty_like(w)
sw = (torch.rand(3 * 4096, device=DEV) * 0.1 + 0.01).contiguous()
zw = torch.randint(0, 16, (3 * 4096,), device=DEV).float()
ms = timed(lambda: abi_uniform(w, wb, sw, zw, 4096, 3 * 4096, 8))
res['uniform_fakequant per-row'] = 8 * w.numel() / ms / 1e6
xg = torch.nn.functional.gelu(x)
t1, t2 = O.adalog_tables(29, 8)
t1, t2 = t1.to(DEV).float().contiguous(), t2.to(DEV).float().contiguous()
q = torch.tensor([29], device=DEV)
sc, sh = torch.tensor([2.7], device=DEV), torch.tensor([O.SHIFT_GELU], device=DEV)
ms = timed(lambda: abi_log(xg, yb, sc, q, t1, t2, sh, 1, 8))
res['shift_adalog_fakequant'] = 8 * n / ms / 1e6
p = torch.softmax(torch.randn(128, 12, 197, 197, device=DEV), -1)   # 238 MB
pb = torch.empty_like(p)
one = torch.ones(1, device=DEV)
ms = timed(lambda: abi_log(p, pb, one, q, t1, t2, None, 0, 8))
res['adalog_fakequant post-softmax'] = 8 * p.numel() / ms / 1e6
# calibration: a plain elementwise copy kernel (torch) on the same tensor, same accounting (4 B read + 4 B write)
yc = torch.empty_like(x)
ms = timed(lambda: yc.copy_(x))
res['torch copy_ kernel (calibration)'] = 8 * n / ms / 1e6
ms = timed(lambda: torch.mul(x, 1.5, out=yc))
res['torch mul kernel (calibration)'] = 8 * n / ms / 1e6
# the reference's eager chain for comparison (same device)
ms = timed(lambda: O.uniform_fakequant(x, s1, z1, 8), 3)
res['reference eager uniform chain (torch)'] = 8 * n / ms / 1e6
ms = timed(lambda: O.adalog_fakequant(p, one.view(1, 1, 1, 1), q, 8, t1, t2), 3)
res['reference eager adalog chain (torch)'] = 8 * p.numel() / ms / 1e6
# self-error sweep of the activations (linear.py:320-345): 4 B per element read ONCE for all 128 candidates -- by bytes it
# is far from the HBM roof because 128 candidates x ~7 FP32 operations are evaluated per element (FMA-pipe bound)
xa = torch.randn(128 * 197, 384, device=DEV) * 2
acs = (torch.rand(1, 128, device=DEV) * 0.3 + 0.05)
acz = torch.randint(0, 8, (1, 128), device=DEV).float()
ms = timed(lambda: ops.sweep_err_a_self(xa, acs, acz, 4, False))
res['sweep_err_a_self (128 candidates per element)'] = 4 * xa.numel() / ms / 1e6
res_extra = {'sweep_err_a_self_ms_25216x384': ms,
             'sweep_err_a_self_candidate_elements_per_s': 128 * xa.numel() / (ms / 1e3)}
# exact radix selection (K11): 4 passes over the tensor, 4 B per element and pass
ranks = torch.tensor([int(0.1 * n), int(0.9 * n), n - 1], device=DEV)
ms = timed(lambda: ops.select_kth(x.view(1, -1), ranks), 3)
res['select_kth (4 radix passes, 16 B / element)'] = 16 * n / ms / 1e6
res_extra['select_kth_ms_77M'] = ms
ms = timed(lambda: x.view(-1).sort(), 3)
res_extra['torch_sort_ms_77M'] = ms
for k, v in res.items():
    print(f'{k:42s} {v:8.0f} GB/s algorithmic  = {100 * v / peak:5.1f}% of measured HBM peak ({peak:.0f} GB/s)')
print(res_extra)
json.dump(dict(peak_gbs=peak, achieved_gbs=res, extra=res_extra), open(os.path.join(ROOT, 'gpurun_out', 'hbm_probe.json'), 'w'), indent=1)
