"""Timing of the opt-in tensor-core inference forward of one linear layer (adalog_gemm_dequant) against the default
composition (fake-quant kernels + FP32 F.linear); not a pytest file.  python tests/gpu_tc_forward_bench.py [dim bits]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import adalog_oracle as O  # noqa: E402
from adalog_b200 import sweep  # noqa: E402
from adalog_b200.quantizers import UniformQuantizer  # noqa: E402

DEV = 'cuda'
D = int(sys.argv[1]) if len(sys.argv) > 1 else 384
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nl = 2 ** (bits - 1)


def uq(s, z):
    q = UniformQuantizer(bits)
    q.scale, q.zero_point, q.inited = s, z, True
    return q


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for tokens in (32 * 197, 128 * 197):
    for name, in_f, out_f in (('qkv', D, 3 * D), ('proj', D, D), ('fc1', D, 4 * D)):
        torch.manual_seed(1)
        x = torch.randn(tokens, in_f, device=DEV)
        W = torch.nn.init.trunc_normal_(torch.empty(out_f, in_f, device=DEV), std=.02)
        b = torch.zeros(out_f, device=DEV)
        wcs, wcz = O.weight_candidates(W, 1, nl, 128)
        wq = uq(wcs[64].clone(), wcz[64].clone().float())
        acs, acz = O.activation_candidates(x.view(1, tokens, in_f), nl, 128, False)
        aq = uq(acs[:, 64].clone(), acz[:, 64].clone().float())
        cache = {}
        W3 = W.view(1, out_f, in_f)
        t_tc = timeit(lambda: sweep.linear_quant_forward(x, W3, b, wq, aq, cache))
        t_def = timeit(lambda: torch.nn.functional.linear(aq(x), wq(W3).view(out_f, in_f), b))
        ops_ = 2.0 * tokens * in_f * out_f
        print(f'tokens {tokens:6d} {name:5s} K={in_f:5d} N={out_f:5d}: tensor-core {t_tc:7.3f} ms ({ops_ / t_tc / 1e9:6.0f} Tops/s)  '
              f'default {t_def:7.3f} ms', flush=True)
