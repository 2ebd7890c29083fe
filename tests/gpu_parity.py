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
     * max relative difference of the per-candidate scores against the reference's FP32 scores (north_star's bar:
       1e-5).  An evaluation above the bar is excused only if the product is CLOSER to the FP64 evaluation than the
       reference's FP32 scores are, i.e. the excess is the reference's own rounding; `over_bar_unexcused` counts the
       rest and must be 0.  Per sweep type the record also holds max |product - fp64| beside max |ref32 - fp64|;
     * where F's own top-k differs from R's: the score gap (in R's FP32 scores) between the worst candidate F took
       and the k-th best of R, against the evaluation's reference noise 2 max_p |ref32 - ref64| (two scores, each off
       by up to the noise); a flip inside that margin is one the reference itself would make under a different FP32
       summation order;
     * exact ties of R -- candidates with equal scores in FP32 AND in the FP64 re-evaluation, i.e. ties of the
       algorithm rather than of the FP32 resolution -- must be exact ties in F.
     Forced along R's selections the final checkpoint and logits must be bit-identical to R's.
  P  CUDA sweeps, free running: quantizer tensors bit-identical to R / total, top-1 agreement on a probe batch.
  R64 oracle-scored with FP64 contractions, free running: the same two numbers for the reference against ITSELF under
     a different (here: exact) summation -- the yardstick for P.  FPCS refines candidates until their scores differ
     by less than the FP32 noise, so late selections of the reference are decided by its own rounding; two legitimate
     evaluations of the reference therefore end in different parameters, and on random-init weights the logits of
     two such W3/W4 calibrations decorrelate visibly.  P is held to agree with R at least as well as R64 does.
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

    def wrap(self, fn, kind=''):
        def both(*a, **kw):
            self.kind = kind
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
            self.evals.append(dict(sims=inp.detach().clone(), k=k, dim=dim, idx=res[1].clone(), sims64=s64,
                                   kind=getattr(self, 'kind', '?')))
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


def oracle_calibrate(model_name, bits, images, bs, init_state, dev, gemm_dtype=None, memory=24 * 2 ** 30):
    """free-running calibration scored by the oracle on `dev` (FP32 like the reference, or FP64 contractions)"""
    mp = _Patch()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    O.GEMM_DTYPE = gemm_dtype
    try:
        fake.install(mp, bs, memory)
        model, loader = _build(model_name, bits, images, bs, init_state, dev)
        model = _calibrate(model, loader)
        mp.undo()
        return model
    finally:
        O.GEMM_DTYPE = None
        torch.backends.cudnn.allow_tf32 = tf32
        mp.undo()


def run_parity(model_name, bits, n_img, bs=32, seed=5, dev='cuda', probe_extra=96, conv_evals=6, log=print,
               images=None, init_state=None, with_ref64=True):
    torch.manual_seed(seed)
    if init_state is None:
        init_state = {k: v.clone() for k, v in zoo.create_model(model_name).state_dict().items()}
    if images is None:
        images = torch.randn(n_img, 3, 224, 224, device=dev)
    images = images.to(dev)
    n_img = images.shape[0]
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
            mp.setattr(sweep, name, tap.wrap(getattr(sweep, name), name))
        ref_model, loader = _build(model_name, bits, images, bs, init_state, dev)
        t0.record()
        with tap:
            ref_model = _calibrate(ref_model, loader)
        t1.record()
        torch.manual_seed(seed + 1)
        probe = torch.cat([images, torch.randn(probe_extra, *images.shape[1:], device=dev)])
        mp.undo()
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
    over_bar, over_bar_unexcused = 0, 0
    ties_ref, ties_kept, noise_evals = 0, 0, 0
    by_kind = {}
    for i, (g, o) in enumerate(zip(forced.got, tap.evals)):
        sr = o['sims'].reshape(g['sims'].shape)
        rd = rel_diff(g['sims'], sr)
        bk = by_kind.setdefault(o['kind'], dict(evals=0, max_rel_diff_vs_ref32=0.0, product_vs_fp64=0.0, ref32_vs_fp64=0.0,
                                                sets_differing=0, outside_noise=0, worst_gap_over_noise=0.0))
        bk['evals'] += 1
        bk['max_rel_diff_vs_ref32'] = max(bk['max_rel_diff_vs_ref32'], rd)
        e_prod = e_ref = None
        if o['sims64'] is not None:
            s64 = o['sims64'].reshape(sr.shape)
            e_prod, e_ref = rel_diff(g['sims'], s64), rel_diff(sr, s64)
            bk['product_vs_fp64'] = max(bk['product_vs_fp64'], e_prod)
            bk['ref32_vs_fp64'] = max(bk['ref32_vs_fp64'], e_ref)
        if rd > 1e-5:
            over_bar += 1
            if e_prod is None or e_prod > e_ref:
                over_bar_unexcused += 1
        if i < conv_evals:
            worst_conv = max(worst_conv, rd)
        else:
            worst = max(worst, rd)
        dim = g['dim']
        order, tied = _ties(sr, dim)
        if bool(tied.any()):
            # STRUCTURAL ties only: two candidates the reference cannot tell apart in FP64 either (identical fake-
            # quantised tensors, e.g. equal scale and an unclamped zero-point shift).  Candidates of the late FPCS steps
            # lie so close that their FP32 scores coincide by resolution alone; those are not ties of the algorithm.
            d = dim % sr.dim()
            n = sr.shape[d]
            if o['sims64'] is not None:
                s64o = torch.gather(o['sims64'].reshape(sr.shape), d, order)
                tied = tied & (s64o.narrow(d, 0, n - 1) == s64o.narrow(d, 1, n - 1))
            gs = torch.gather(g['sims'], d, order)
            kept = (gs.narrow(d, 0, n - 1) == gs.narrow(d, 1, n - 1)) & tied
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
                bk['sets_differing'] += 1
                bk['worst_gap_over_noise'] = max(bk['worst_gap_over_noise'], ratio)
                if ratio > 1.0:
                    flips_outside += 1
                    bk['outside_noise'] += 1
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

    # ---- R64: the reference against itself under FP64 contractions, free running
    r64 = {}
    if with_ref64:
        m64 = oracle_calibrate(model_name, bits, images, bs, init_state, dev, torch.float64)
        sd64 = m64.state_dict()
        with torch.no_grad():
            l64 = m64(probe)
        r64 = dict(reference_fp64_quantizer_tensors_identical=sum(int(torch.equal(sd64[k], ref_state[k])) for k in qkeys),
                   reference_fp64_top1_agreement=float((l64.argmax(-1) == ref_logits.argmax(-1)).float().mean()),
                   reference_fp64_logits_rel_diff=float((l64 - ref_logits).norm() / ref_logits.norm()))
        del m64
    rec = dict(model=model_name, bits=bits, images=n_img, evaluations=len(tap.evals),
               forced_max_rel_diff=worst, forced_max_rel_diff_patch_embed=worst_conv,
               evaluations_over_1e5=over_bar, evaluations_over_1e5_unexcused=over_bar_unexcused,
               forced_checkpoint_bit_identical=bool(forced_identical), forced_logits_bit_identical=forced_logits_identical,
               topk_sets_differing=flips, topk_sets_differing_outside_reference_noise=flips_outside,
               worst_gap_over_reference_noise=worst_ratio, exact_ties_in_reference=ties_ref, exact_ties_preserved=ties_kept,
               free_running_quantizer_tensors_identical=same, quantizer_tensors_total=len(qkeys),
               free_running_top1_agreement=agree, free_running_logits_rel_diff=float((logits - ref_logits).norm() / ref_logits.norm()),
               probe_images=int(probe.shape[0]), reference_on_gpu_seconds=ref_seconds,
               product_seconds=p0.elapsed_time(p1) / 1e3,
               reference_on_gpu_note='oracle on CUDA, every evaluation scored in FP32 and again in FP64',
               by_sweep=by_kind, **r64)
    log('[parity] ' + ', '.join(f'{k}={v}' for k, v in rec.items()))
    return rec
