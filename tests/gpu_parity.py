"""Free-running parity of a whole calibration against "the reference on the same GPU" (test infrastructure; also
imported by bench.py's checker leg -- never by the product).

The reference side is the product's host logic (bit-exact against the unmodified reference on CPU,
tests/test_model_golden.py) scored by the oracle (tests/_oracle_backend.py, pinned to the reference by
tests/test_oracle_golden.py) on CUDA: torch-CUDA FP32 contractions, torch.topk on the same device.  Three runs of the
same seeded model and images:

  R  oracle-scored, free running.  Every evaluation is scored twice -- FP32 like the reference, and with the
     contractions in FP64 (adalog_oracle.GEMM_DTYPE) -- so each evaluation carries the reference's OWN rounding noise
     |ref32 - ref64| next to its similarities and selection.
  F  CUDA sweeps, teacher-forced along R's trajectory: every evaluation compared on identical candidates.
     * max relative difference of the per-candidate scores (bar 1e-5; 3e-5 for the split-3 patch embedding);
     * where F's own top-k differs from R's: the score gap (in R's FP32 scores) between the worst candidate F took
       and the k-th best of R, against the evaluation's reference noise 2 max_p |ref32 - ref64| (two scores, each off
       by up to the noise); a flip inside that margin is one the reference itself would make under a different FP32
       summation order;
     * exact ties of R (equal FP32 bits between candidates) must be exact ties in F.
     Forced along R's selections the final checkpoint and logits must be bit-identical to R's.
  P  CUDA sweeps, free running: quantizer tensors bit-identical to R / total, top-1 agreement on a probe batch.
"""
import importlib

import torch

import adalog_oracle as O
import _oracle_backend as fake
from gpu_util import ForcedTopk, rel_diff
from adalog_b200.utils import models as zoo
from adalog_b200.utils.calibrator import QuantCalibrator
from adalog_b200.utils.wrap_net import wrap_modules_in_net, wrap_reparamed_modules_in_net

GEMM_SWEEPS = ('linear_err_w_self', 'linear_err_a_self', 'linear_err_w', 'linear_err_a', 'linear_err_log', 'linear_err_a_twin', 'matmul_err_A', 'matmul_err_B',
               'matmul_err_A_log_base', 'conv_err_w')


class _Patch:
    """minimal monkeypatch stand-in (bench.py has no pytest fixture)"""

    def __init__(self):
        self.saved = []

    def setattr(self, obj, name, value):
        self.saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)

    def undo(self):
        for obj, name, old in reversed(self.saved):
            setattr(obj, name, old)
        self.saved = []


class NoiseTap:
    """records every torch.topk of run R together with the FP64 re-evaluation of the same candidates"""

    def __init__(self):
        self.evals = []
        self.pending64 = None
        self._orig = torch.topk

    def wrap(self, fn):
        def both(*a, **kw):
            s32 = fn(*a, **kw)
            tf32 = torch.backends.cudnn.allow_tf32
            O.GEMM_DTYPE = torch.float64
            try:
                self.pending64 = fn(*a, **kw).double()
            finally:
                O.GEMM_DTYPE = None
                torch.backends.cudnn.allow_tf32 = tf32
            return s32
        return both

    def __enter__(self):
        def tapped(inp, k, dim=-1, **kw):
            res = self._orig(inp, k=k, dim=dim, **kw)
            s64 = self.pending64 if (self.pending64 is not None and self.pending64.shape == inp.shape) else None
            self.pending64 = None
            self.evals.append(dict(sims=inp.detach().clone(), k=k, dim=dim, idx=res[1].clone(), sims64=s64))
            return res
        torch.topk = tapped
        return self

    def __exit__(self, *a):
        torch.topk = self._orig


def _build(model_name, bits, images, bs, init_state, dev):
    cfg = importlib.import_module(f'adalog_b200.configs.{bits}bit').Config()
    cfg.calib_size, cfg.calib_batch_size = images.shape[0], bs
    model = zoo.create_model(model_name).eval()
    model.load_state_dict(init_state)
    model = wrap_modules_in_net(model.to(dev), cfg, reparam=True).to(dev).eval()
    loader = [(images[i:i + bs], None) for i in range(0, images.shape[0], bs)]
    return model, loader


def _calibrate(model, loader):
    cal = QuantCalibrator(model, loader)
    cal.progress = False
    cal.batching_quant_calib()
    model = wrap_reparamed_modules_in_net(model)
    for _, m in model.named_modules():
        if hasattr(m, 'mode') and hasattr(m, 'reparam_bias'):
            m.reparam_bias()
    return model


def _kth_gap(sims_ref, idx_ref, idx_got, dim):
    """max over slices of (k-th best reference score - worst reference score among the product's picks) >= 0"""
    s = sims_ref.double()
    d = dim % s.dim()
    shape = list(s.shape)
    k = idx_ref.numel() // max(1, s.numel() // shape[d])
    shape[d] = k
    ref = torch.gather(s, d, idx_ref.reshape(shape)).min(dim=d).values
    got = torch.gather(s, d, idx_got.reshape(shape)).min(dim=d).values
    return (ref - got).clamp_min(0)                                  # [slices]


def _ties(sims, dim):
    """boolean mask of candidates whose FP32 score equals the score of the next-ranked candidate (exact ties)"""
    srt, order = sims.sort(dim=dim)
    n = sims.shape[dim]
    a = srt.narrow(dim, 0, n - 1)
    b = srt.narrow(dim, 1, n - 1)
    return order, (a == b)


def run_parity(model_name, bits, n_img, bs=32, seed=5, dev='cuda', probe_extra=96, conv_evals=6, log=print):
    torch.manual_seed(seed)
    init_state = {k: v.clone() for k, v in zoo.create_model(model_name).state_dict().items()}
    images = torch.randn(n_img, 3, 224, 224, device=dev)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False        # the reference's convolution in true FP32 (see test_gpu_gemm)

    # ---- R: oracle-scored, free running, each evaluation also in FP64
    mp = _Patch()
    tap = NoiseTap()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    try:
        fake.install(mp, bs, 24 * 2 ** 30)
        from adalog_b200 import sweep
        for name in GEMM_SWEEPS:
            mp.setattr(sweep, name, tap.wrap(getattr(sweep, name)))
        ref_model, loader = _build(model_name, bits, images, bs, init_state, dev)
        t0.record()
        with tap:
            ref_model = _calibrate(ref_model, loader)
        t1.record()
        torch.manual_seed(seed + 1)
        probe = torch.cat([images, torch.randn(probe_extra, 3, 224, 224, device=dev)])
        with torch.no_grad():
            ref_logits = ref_model(probe)
    finally:
        mp.undo()
    torch.cuda.synchronize()
    ref_state = {k: v.clone() for k, v in ref_model.state_dict().items()}
    ref_seconds = t0.elapsed_time(t1) / 1e3

    # ---- F: CUDA sweeps, teacher-forced along R
    model, loader = _build(model_name, bits, images, bs, init_state, dev)
    with ForcedTopk(tap.evals) as forced:
        model = _calibrate(model, loader)
    assert len(forced.got) == len(tap.evals)
    worst, worst_conv, flips, flips_outside, worst_ratio = 0.0, 0.0, 0, 0, 0.0
    ties_ref, ties_kept, noise_evals = 0, 0, 0
    for i, (g, o) in enumerate(zip(forced.got, tap.evals)):
        sr = o['sims'].reshape(g['sims'].shape)
        rd = rel_diff(g['sims'], sr)
        if i < conv_evals:
            worst_conv = max(worst_conv, rd)
        else:
            worst = max(worst, rd)
        dim = g['dim']
        order, tied = _ties(sr, dim)
        if bool(tied.any()):
            gs = torch.gather(g['sims'], dim % sr.dim(), order)
            n = sr.shape[dim]
            kept = (gs.narrow(dim, 0, n - 1) == gs.narrow(dim, 1, n - 1)) & tied
            ties_ref += int(tied.sum())
            ties_kept += int(kept.sum())
        if not torch.equal(g['idx'], o['idx'].reshape(g['idx'].shape)):
            gap = _kth_gap(sr, o['idx'].reshape(g['idx'].shape), g['idx'], dim)
            if float(gap.max()) > 0:                      # a different SET of candidates, not just tie order
                flips += 1
                if o['sims64'] is not None:
                    noise_evals += 1
                    noise = 2 * (sr.double() - o['sims64'].reshape(sr.shape)).abs().amax(dim=dim % sr.dim())
                    ratio = float((gap / noise.reshape(gap.shape).clamp_min(1e-300)).max())
                else:
                    ratio = float('inf')
                worst_ratio = max(worst_ratio, ratio)
                if ratio > 1.0:
                    flips_outside += 1
    sd = model.state_dict()
    forced_identical = all(torch.equal(sd[k], v) for k, v in ref_state.items())
    with torch.no_grad():
        forced_logits_identical = bool(torch.equal(model(probe), ref_logits))

    # ---- P: CUDA sweeps, free running
    model, loader = _build(model_name, bits, images, bs, init_state, dev)
    p0 = torch.cuda.Event(enable_timing=True)
    p1 = torch.cuda.Event(enable_timing=True)
    p0.record()
    model = _calibrate(model, loader)
    p1.record()
    sd = model.state_dict()
    qkeys = [k for k in ref_state if 'quantizer' in k]
    same = sum(int(torch.equal(sd[k], ref_state[k])) for k in qkeys)
    with torch.no_grad():
        logits = model(probe)
    torch.cuda.synchronize()
    agree = float((logits.argmax(-1) == ref_logits.argmax(-1)).float().mean())
    torch.backends.cudnn.allow_tf32 = tf32
    rec = dict(model=model_name, bits=bits, images=n_img, evaluations=len(tap.evals),
               forced_max_rel_diff=worst, forced_max_rel_diff_patch_embed=worst_conv,
               forced_checkpoint_bit_identical=bool(forced_identical), forced_logits_bit_identical=forced_logits_identical,
               topk_sets_differing=flips, topk_sets_differing_outside_reference_noise=flips_outside,
               worst_gap_over_reference_noise=worst_ratio, exact_ties_in_reference=ties_ref, exact_ties_preserved=ties_kept,
               free_running_quantizer_tensors_identical=same, quantizer_tensors_total=len(qkeys),
               free_running_top1_agreement=agree, free_running_logits_rel_diff=float((logits - ref_logits).norm() / ref_logits.norm()),
               probe_images=int(probe.shape[0]), reference_on_gpu_seconds=ref_seconds,
               product_seconds=p0.elapsed_time(p1) / 1e3,
               reference_on_gpu_note='oracle on CUDA, every evaluation scored in FP32 and again in FP64')
    log('[parity] ' + ', '.join(f'{k}={v}' for k, v in rec.items()))
    return rec
