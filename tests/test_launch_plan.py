"""CPU: invariants of the static launch plans (pure host logic, no kernel call).  A wrong plan would not crash -- the
kernel validates and the partition is computed from (UG, cpg, S) on the device -- but it could leave SMs idle or break
the 1024-column limit of the column-scale staging, so the planners are pinned here."""
import math

import pytest
from hypothesis import given, settings, strategies as st

from adalog_b200 import ops, sweep


@settings(max_examples=300, deadline=None)
@given(nu=st.integers(1, 30000), NT=st.integers(1, 99), ka=st.sampled_from([64, 128, 384, 768, 1536, 3072]),
       N=st.sampled_from([64, 197, 384, 768, 1152, 3072, 25216]), i8=st.booleans(), cs=st.booleans())
def test_linear_launch_plan_invariants(nu, NT, ka, N, i8, cs):
    BN = ops.pick_bn(N)
    NT = (N + BN - 1) // BN
    esz = 1 if i8 else 2
    max_tps = (1024 // BN) if cs else 1 << 30
    groups, upc, cpg, S, sf = sweep._launch_plan(nu, nu, NT, N * ka * esz, 128 * ka * esz,
                                                 max((ka // 64) * BN * (1 if i8 else 2), 8 * BN), max_tps)
    assert groups == 1 and 1 <= upc <= nu and cpg == math.ceil(nu / upc) and 1 <= S <= NT
    assert math.ceil(NT / S) <= max_tps                       # a CTA never covers more than 1024 staged columns
    # every unit and every N tile is covered exactly once by the device-side partition
    units = [(ci * nu) // cpg for ci in range(cpg + 1)]
    assert units[0] == 0 and units[-1] == nu and all(b >= a for a, b in zip(units, units[1:]))
    assert max(b - a for a, b in zip(units, units[1:])) <= upc
    tiles = [(y * NT) // S for y in range(S + 1)]
    assert tiles[0] == 0 and tiles[-1] == NT and all(b > a for a, b in zip(tiles, tiles[1:]))
    # L2 rule: when 148 units of candidate rows cannot sit in L2 the split-fast order is used
    assert sf == (NT > 1 and 128 * ka * esz * min(sweep.NUM_SMS, nu) > sweep.L2_CAND_BYTES)


@pytest.mark.parametrize('K,N,i8,log,nl,ok', [
    (64, 197, False, False, 4, True), (197, 64, False, True, 4, True), (197, 197, False, False, 8, True),
    (32, 49, False, False, 32, True), (49, 32, False, True, 32, True), (144, 144, False, False, 8, True),
    (257, 64, False, False, 8, False),        # K > 4 blocks of 64
    (64, 300, False, False, 8, False),        # more than one N tile
    (197, 64, True, True, 4, False),          # AdaLog candidates are bf16
    (197, 64, False, True, 64, False),        # 7-bit AdaLog is not bf16-exact
])
def test_fused_plan_eligibility(K, N, i8, log, nl, ok):
    plan = ops.fused_plan(K, N, i8, log, nl)
    assert (plan is not None) == ok
    if ok:
        KB, BN = plan
        assert KB == math.ceil(K / (128 if i8 else 64)) <= 4 and N <= BN <= 256 and BN % 16 == 0


@settings(max_examples=200, deadline=None)
@given(groups=st.integers(1, 20000), ug=st.integers(1, 400), K=st.integers(8, 256), N=st.integers(8, 256))
def test_fused_units_per_cta(groups, ug, K, N):
    upc = sweep._fused_upc(groups, ug, K, N)
    assert 1 <= upc <= ug
    cpg = math.ceil(ug / upc)
    assert cpg <= 16                                           # a CTA amortises its fixed-operand load over >= ug/16 units


def _lin_args(K, N, U, log, i8, nl):
    """adalog_lin_fused_args with dummy (non-null, 16-byte aligned) pointers: _grid / _passes only validate and plan"""
    import ctypes
    from adalog_b200 import _lib
    a = _lib.LinFusedArgs()
    dummy = 1 << 20
    a.x, a.ldx, a.Bm, a.b_rows = dummy, K, dummy, N
    a.K, a.N, a.U, a.P, a.n_levels = K, N, U, 128, nl
    a.gen, a.dtype = (1 if log else 0), (_lib_i8() if i8 else 0)
    a.cs, a.cz, a.cq, a.shift, a.mtab = dummy, dummy, dummy, None, dummy
    a.y, a.ldy, a.rs, a.ccs, a.ccb = dummy, N, dummy, dummy, dummy
    return a, ctypes


def _lib_i8():
    from adalog_b200 import ops
    return ops.I8


@pytest.mark.parametrize('K,N,U,log,i8,nl,passes', [
    (384, 1152, 25216, False, True, 4, 1),      # DeiT-S qkv int8: all 3 K blocks resident
    (384, 384, 25216, False, True, 4, 1),
    (768, 3072, 25216, False, True, 8, 1),      # DeiT-B fc1 int8: 6 K blocks resident
    (1024, 4096, 6272, False, True, 32, 1),     # Swin-B stage-4 fc1: 8 int8 K blocks + one spare stage still fit: resident
    (768, 768, 900, False, False, 128, 2),      # 8-bit uniform (bf16 operands), 12 K blocks: streamed, two passes
    (1536, 384, 25216, True, False, 4, 1),      # DeiT-S fc2 AdaLog: streamed, one 384-column pass
    (3072, 768, 25216, True, False, 8, 2),      # DeiT-B fc2 AdaLog: two 512-column TMEM passes -> caller keeps the workspace path
    (100, 52, 777, False, True, 8, 1),          # ragged K and N
])
def test_lin_fused_schedule(K, N, U, log, i8, nl, passes):
    """host-side schedule of the fused linear sweep (no kernel call): grid = one persistent CTA per SM (or per unit),
    and how often a unit's operand is generated -- the quantity sweep._lin_fused_passes() gates the fused path on"""
    from adalog_b200 import _lib
    lib = _lib.load()
    a, ctypes = _lin_args(K, N, U, log, i8, nl)
    assert lib.adalog_lin_fused_cand_gemm_err_grid(ctypes.byref(a)) == min(U, sweep.NUM_SMS)
    assert lib.adalog_lin_fused_cand_gemm_err_passes(ctypes.byref(a)) == passes


def test_lin_fused_rejects_bad_arguments():
    from adalog_b200 import _lib
    lib = _lib.load()
    a, ctypes = _lin_args(384, 1150, 100, False, True, 4)              # N not a multiple of 4
    assert lib.adalog_lin_fused_cand_gemm_err_grid(ctypes.byref(a)) == -2
    a, ctypes = _lin_args(1536, 384, 100, True, True, 4)               # AdaLog candidates are bf16 operands
    assert lib.adalog_lin_fused_cand_gemm_err_grid(ctypes.byref(a)) == -2
    a, ctypes = _lin_args(384, 384, 100, False, True, 4)
    a.y = None
    assert lib.adalog_lin_fused_cand_gemm_err_passes(ctypes.byref(a)) == -1
    assert b'null pointer' in lib.adalog_last_error()


def test_fixed_operand_cache_keys(monkeypatch):
    """host logic of sweep._cached_fixed (no kernel): one entry per role on the calibration context, rebuilt exactly when
    a key tensor is written through quantizers/_ste.assign (version bump) or replaced, dropped by invalidate_caches"""
    import torch
    from adalog_b200.quantizers._ste import assign, invalidate_caches

    class Ctx:
        pass

    ctx, calls = Ctx(), []
    scale, zp = torch.nn.Parameter(torch.ones(4)), torch.nn.Parameter(torch.zeros(4))

    def get():
        return sweep._cached_fixed(ctx, 'weight', (8, True) + sweep._pkey(scale, zp),
                                   lambda: calls.append(1) or float(scale.sum()))
    assert get() == 4.0 and get() == 4.0 and len(calls) == 1                 # hit
    assign(scale, torch.full((4,), 2.0))
    assert get() == 8.0 and len(calls) == 2                                  # version bump -> rebuilt
    assert get() == 8.0 and len(calls) == 2
    with torch.no_grad():
        scale.data.fill_(3.0)                                                # a write through .data is NOT seen ...
    assert get() == 8.0 and len(calls) == 2

    class Mod(torch.nn.Module):
        pass
    m = Mod()
    m.__dict__['_ctx'] = ctx
    invalidate_caches(m)                                                     # ... until the caller invalidates
    assert get() == 12.0 and len(calls) == 3
    assert sweep._cached_fixed(None, 'weight', (), lambda: 7) == 7            # no context: no caching
    monkeypatch.setattr(sweep, 'FIXED_CACHE', False)
    assert get() == 12.0 and len(calls) == 4                                 # switched off: always rebuilt
