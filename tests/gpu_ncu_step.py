"""ncu launch-list target (not a pytest file): one FPCS calibration of a one-block DeiT-Tiny (312 evaluations) on 32
images -- the same code path as bench.py's step, small enough to sit under `ncu --metrics gpu__time_duration.sum`.
Also prints the live CUDA-event share of the candidate GEMM for comparison with the ncu share."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from adalog_b200 import ops  # noqa: E402

cfg = importlib.import_module('adalog_b200.configs.4bit').Config()
cfg.calib_size, cfg.calib_batch_size = 32, 32
dev = torch.device('cuda', 0)
images = torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(5)).to(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for rep in range(reps):
    model = bench.build_wrapped('deit_tiny_depth1_patch16_224', cfg, dev)
    ops.profile_reset(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bench.calibrate(model, images, 32)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    flops, ms, n = ops.profile_gemm_summary()
    print(f'rep {rep}: wall {wall * 1e3:.1f} ms, cand_gemm_err: {n} launches, {ms:.1f} ms '
          f'({100 * ms / (wall * 1e3):.1f}% of wall), {flops / ms / 1e9:.0f} TFLOP/s')
