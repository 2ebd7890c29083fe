"""GPU helper (not a pytest file): kernel-time table of one calibration of a two-block DeiT-S (W3A3, 128 images) under
torch.profiler -- which kernels (ours and torch's) the step's GPU time goes to, and how much of the wall the GPU is busy.
  python tests/gpu_torchprof.py [bits] [images]"""
import importlib
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 128
name = sys.argv[3] if len(sys.argv) > 3 else 'deit_small_depth2_patch16_224'
cfg = importlib.import_module(f'adalog_b200.configs.{bits}bit').Config()
cfg.calib_size, cfg.calib_batch_size = n_img, 32
dev = torch.device('cuda', 0)
images = torch.randn(n_img, 3, 224, 224, generator=torch.Generator().manual_seed(5)).to(dev)
for rep in range(2):
    model = bench.build_wrapped(name, cfg, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rep == 1:
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            bench.calibrate(model, images, 32)
            torch.cuda.synchronize()
    else:
        bench.calibrate(model, images, 32)
        torch.cuda.synchronize()
    print(f'rep {rep}: wall {time.perf_counter() - t0:.3f} s')
from collections import defaultdict
agg = defaultdict(lambda: [0.0, 0])
lo, hi = None, None
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        dur = e.time_range.end - e.time_range.start
        agg[e.name][0] += dur
        agg[e.name][1] += 1
        lo = e.time_range.start if lo is None else min(lo, e.time_range.start)
        hi = e.time_range.end if hi is None else max(hi, e.time_range.end)
# GPU busy time = union of all kernel / memcpy intervals (streams overlap); the rest of the span is launch latency and
# host glue during which no kernel runs
iv = sorted((e.time_range.start, e.time_range.end) for e in prof.events()
            if e.device_type == torch.autograd.DeviceType.CUDA)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s_, e_ in iv[1:]:
    if s_ > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
print(f'GPU busy (union of kernel intervals) {busy / 1e6:.3f} s of a span of {(hi - lo) / 1e6:.3f} s = '
      f'{100 * busy / (hi - lo):.1f}%; {len(iv)} device activities')
rows = sorted(((t, n, k) for k, (t, n) in agg.items()), reverse=True)
tot = sum(r[0] for r in rows)
print(f'sum of GPU kernel+memcpy durations {tot / 1e6:.3f} s over a GPU span of {(hi - lo) / 1e6:.3f} s '
      f'(streams overlap, so the sum can exceed the busy time)')
for t, n, k in rows[:32]:
    print(f'{t / 1e3:10.2f} ms {100 * t / tot:5.1f}%  n={n:6d}  {k[:100]}')
