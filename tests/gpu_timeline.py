"""GPU helper (not a pytest file): kernel timeline (start, duration, stream) of one chunked sweep of a DeiT-B
fc2-shaped layer under torch.profiler -- shows whether the generator of chunk i+1 really runs under the GEMM of chunk i.
  python tests/gpu_timeline.py [log|a]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from adalog_b200 import sweep  # noqa: E402
from adalog_b200.quantizers import ShiftAdaLogQuantizer, UniformQuantizer  # noqa: E402
import adalog_oracle as O  # noqa: E402  (candidate seeding only)

DEV = torch.device('cuda', 0)
Bn, T, D, Do = 128, 197, 3072, 768
torch.manual_seed(0)
x0 = torch.randn(Bn, T, D, device=DEV)
W = torch.randn(Do, D, device=DEV) * 0.02
b = torch.zeros(Do, device=DEV)
which = sys.argv[1] if len(sys.argv) > 1 else 'log'
wq = UniformQuantizer(4)
cs, cz = O.weight_candidates(W, 1, 8, 128)
wq.scale, wq.zero_point = cs[64].clone(), cz[64].clone().float()
if which == 'log':
    x = torch.nn.functional.gelu(x0)
    ctx = sweep.LinearCtx(x, torch.nn.functional.linear(x, W, b), Do)
    lq = ShiftAdaLogQuantizer(4).to(DEV)
    lq.scale = torch.nn.Parameter(torch.tensor([3.0], device=DEV))
    lq.shift.data.fill_(O.SHIFT_GELU)
    lq.inited = True
    qc = torch.arange(10, 138, device=DEV).view(1, -1)
    sc = torch.linspace(2.0, 4.0, 128, device=DEV).view(1, -1)
    fn = lambda: sweep.linear_err_log(ctx, W.view(1, Do, D), b, wq, lq, sc, qc)
else:
    ctx = sweep.LinearCtx(x0, torch.nn.functional.linear(x0, W, b), Do)
    acs, acz = O.activation_candidates(x0, 8, 128, False)
    fn = lambda: sweep.linear_err_a(ctx, W.view(1, Do, D), b, wq, acs, acz, 8)
fn(); fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    nm = e.name.replace('void adalog::', '').split('(')[0][:40]
    print(f'{(e.time_range.start - t0) / 1e3:9.3f} ms  +{(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms  {nm}')
print(f'span {(ev[-1].time_range.end - t0) / 1e3:.3f} ms')
