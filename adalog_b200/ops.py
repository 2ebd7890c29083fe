"""Thin torch-tensor wrappers over the C ABI (include/adalog_b200.h).

Every function takes CUDA float32 tensors, enqueues on torch's current stream and returns tensors.
No CPU path exists: CPU tensors raise AdalogError.
"""
import ctypes
import math
import threading

import torch

from . import _lib
from ._lib import AdalogError, FusedArgs, GemmErrArgs, LinFusedArgs
from ._lib import call as _lib_call

P_TILE = 128   # ADALOG_P
BK = 64        # ADALOG_BK

# bench.py sets PROFILE['on']: every candidate-GEMM launch is then bracketed by CUDA events on its stream and its
# algorithmic FLOPs recorded, which is how roofline.achieved is measured live inside the timed region
PROFILE = {'on': False, 'gemm': [], 'fused': [], 'lin': []}


def profile_reset(on):
    PROFILE['on'] = bool(on)
    PROFILE['gemm'] = []
    PROFILE['fused'] = []
    PROFILE['lin'] = []


def profile_gemm_summary(split=False):
    """(flops, milliseconds, launches) of the candidate-GEMM launches recorded since profile_reset(True);
    split=True: {'bf16': (...), 'i8': (...)} by operand type (int8 MMAs run at twice the bf16 rate)."""
    torch.cuda.synchronize()
    def tot(rows):
        return sum(r[2] for r in rows), sum(r[0].elapsed_time(r[1]) for r in rows), len(rows)
    if split:
        return {'bf16': tot([r for r in PROFILE['gemm'] if not r[3]]), 'i8': tot([r for r in PROFILE['gemm'] if r[3]])}
    return tot(PROFILE['gemm'])


_tls = threading.local()


def _cuda(*ts):
    """All operands of one kernel call must be CUDA tensors on ONE device; that device becomes the target of the call:
    `call` below switches to it for the launch (TMA descriptors, kernel attributes and the launch itself are per-device
    state) and `_stream()` hands the C ABI torch's current stream OF THAT DEVICE -- a model on cuda:1 while the current
    device is 0 launches on cuda:1, not on device 0's stream with device-1 pointers."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise AdalogError('adalog_b200 kernels need CUDA tensors (there is no CPU fallback)')
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise AdalogError(f'operands of one adalog_b200 call sit on different devices ({dev} and {t.device})')
    _tls.dev = dev
    return dev


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream(getattr(_tls, 'dev', None)).cuda_stream)


def call(name, *args):
    dev = getattr(_tls, 'dev', None)
    if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
        return _lib_call(name, *args)
    with torch.cuda.device(dev):
        return _lib_call(name, *args)


def _f32(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


BF16, I8 = 0, 1   # ADALOG_BF16 / ADALOG_I8


def kpad(K, i8=False):
    """operand row pitch in ELEMENTS: whole 128-byte swizzle rows (64 bf16 or 128 int8)"""
    blk = 2 * BK if i8 else BK
    return ((K + blk - 1) // blk) * blk


def _prod(xs):
    r = 1
    for v in xs:
        r *= int(v)
    return r


def group_layout(x_shape, s_shape):
    """How a broadcast scale maps onto flat x: group(i) = (i // inner) % ngroups."""
    s_shape = list(s_shape)
    while len(s_shape) > len(x_shape) and s_shape[0] == 1:
        s_shape = s_shape[1:]
    if _prod(s_shape) == 1:
        return 1, 1
    if len(s_shape) > len(x_shape):
        raise AdalogError(f'scale shape {s_shape} does not broadcast onto {list(x_shape)}')
    s = [1] * (len(x_shape) - len(s_shape)) + s_shape
    nz = [i for i, d in enumerate(s) if d != 1]
    d0, d1 = nz[0], nz[-1]
    for i in range(d0, d1 + 1):
        if s[i] != x_shape[i]:
            raise AdalogError(f'scale shape {s_shape} does not broadcast onto {list(x_shape)}')
    return _prod(x_shape[d1 + 1:]), _prod(x_shape[d0:d1 + 1])


# ------------------------------------------------------------------------------------------ forwards
def _round_ste_cached(zero_point, sc):
    """round_ste(zero_point) (_ste.py:5-6: (z.round() - z) + z in FP32) as a flat tensor matching `sc`; remembered on
    the zero-point tensor itself until it is modified, so an inference forward is ONE kernel launch, not four."""
    key = (zero_point.data_ptr(), zero_point._version, sc.numel())
    hit = getattr(zero_point, '_adalog_zr', None)
    if hit is not None and hit[0] == key:
        return hit[1]
    z = _f32(zero_point.detach()).reshape(-1)
    zr = (z.round() - z) + z
    if zr.numel() != sc.numel():
        zr = zr.expand_as(sc).contiguous()
    try:
        zero_point._adalog_zr = (key, zr)
    except AttributeError:
        pass
    return zr


def uniform_fakequant(x, scale, zero_point, n_levels, sym=False, want_codes=False, want_y=True):
    """quantizers/uniform.py:25-36 on device."""
    _cuda(x, scale, zero_point)
    xc = _f32(x)
    inner, ngroups = group_layout(xc.shape, scale.shape)
    sc = _f32(scale.detach()).reshape(-1)
    zr = None if sym else _round_ste_cached(zero_point, sc)
    y = torch.empty_like(xc) if want_y else None
    codes = torch.empty(xc.shape, dtype=torch.int16, device=xc.device) if want_codes else None
    call('adalog_uniform_fakequant_f32', _p(xc), _p(y), _p(codes), xc.numel(), _p(sc), _p(zr), inner, ngroups,
         int(n_levels), int(bool(sym)), _stream())
    if want_codes:
        return (y, codes) if want_y else codes
    return y


LOG2, LOGSQRT2, ADALOG = 0, 1, 2


def log_fakequant(x, scale, kind, n_levels, q=None, table1=None, table2=None, shift=None, sub_shift=False,
                  want_codes=False):
    """quantizers/logarithm.py forwards (inference branch) on device; per-tensor scale."""
    _cuda(x, scale, q, table1, table2, shift)
    xc = _f32(x)
    if scale.numel() != 1:
        raise AdalogError('log-family quantizers take a per-tensor scale')
    sc = _f32(scale.detach()).reshape(-1)
    y = torch.empty_like(xc)
    codes = torch.empty(xc.shape, dtype=torch.int16, device=xc.device) if want_codes else None
    qq = q.detach().to(torch.int64).contiguous() if q is not None else None
    t1 = _f32(table1) if table1 is not None else None
    t2 = _f32(table2) if table2 is not None else None
    sh = _f32(shift.detach()).reshape(-1) if shift is not None else None
    call('adalog_log_fakequant_f32', _p(xc), _p(y), _p(codes), xc.numel(), _p(sc), int(kind), int(n_levels), _p(qq),
         _p(t1), _p(t2), _p(sh), int(bool(sub_shift)), _stream())
    return (y, codes) if want_codes else y


def twin_fakequant(x, scale2, n_levels):
    _cuda(x, scale2)
    xc = _f32(x)
    y = torch.empty_like(xc)
    call('adalog_twin_fakequant_f32', _p(xc), _p(y), xc.numel(), _p(_f32(scale2.detach()).reshape(-1)), int(n_levels),
         _stream())
    return y


# ------------------------------------------------------------------------------------------ self sweeps
def sweep_err_w_self(W2d, cs, cz, n_levels):
    """W2d [R,K]; cs/cz [P,R] -> FP64 [P,R] = sum_k (w - dequant_p(w))^2   (linear.py:296-309)."""
    _cuda(W2d, cs, cz)
    W2d, cs, cz = _f32(W2d), _f32(cs), _f32(cz)
    R, K = W2d.shape
    P = cs.shape[0]
    out = torch.empty(P, R, dtype=torch.float64, device=W2d.device)
    call('adalog_sweep_err_w_self', _p(W2d), R, K, _p(cs), _p(cz), P, int(n_levels), _p(out), _stream())
    return out


def sweep_err_a_self(x, cs, cz, n_levels, per_channel):
    """x [..., C]; cs/cz [C|1, P] -> FP64 [C|1, P] = sum over all rows of (x - dequant_p(x))^2."""
    _cuda(x, cs, cz)
    xc, cs, cz = _f32(x), _f32(cs), _f32(cz)
    C = xc.shape[-1]
    G, P = cs.shape
    n = xc.numel()
    cw = C if per_channel else 32
    rows = (n + cw - 1) // cw
    col_tiles = (cw + 31) // 32
    nsplit = int(max(1, min(rows, (148 * 8 + col_tiles - 1) // col_tiles)))
    partial = torch.empty(nsplit, C if per_channel else 1, P, dtype=torch.float64, device=xc.device)
    call('adalog_sweep_err_a_self', _p(xc), n, C, int(bool(per_channel)), _p(cs), _p(cz), P, int(n_levels),
         _p(partial), nsplit, _stream())
    return partial.sum(dim=0)


# ------------------------------------------------------------------------------------------ order statistics
def select_kth(x2d, ranks, positive_only=False, reduce_hist=None, rows_total=None, row0=0):
    """Exact order statistics by radix selection (adalog_select_*): x2d [rows, n] FP32 with contiguous rows, ranks [T]
    int64 on the device (T <= 8, the same for every row) -> [rows_total, T] float32 = sort(row)[ranks], bit for bit.
    reduce_hist: callable(hist_tensor) -> hist_tensor, e.g. an all-reduce(SUM) over the ranks of a data-parallel job
    (each holding a shard of every row, or local rows [row0, row0 + rows) of rows_total): the statistics are then those
    of the union of the shards."""
    _cuda(x2d, ranks)
    assert x2d.dtype == torch.float32 and x2d.dim() == 2 and x2d.stride(1) == 1
    rows, n = x2d.shape
    R = int(rows_total if rows_total is not None else rows)
    T = int(ranks.numel())
    lib = _lib.load()
    nbytes = lib.adalog_select_workspace_bytes(R, T)
    if nbytes < 0:
        raise AdalogError('select_kth: at most 8 target ranks per call')
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x2d.device)
    rk = ranks.to(torch.int64).contiguous()
    call('adalog_select_init', _p(ws), R, T, _p(rk), _stream())
    hist = None
    if reduce_hist is not None:
        off = R * T * 16
        hist = ws[off:off + R * T * 256 * 4].view(torch.int32)
    for p in range(4):
        call('adalog_select_hist', _p(x2d), rows, n, int(x2d.stride(0)), _p(ws), R, int(row0), T, p,
             int(bool(positive_only)), _stream())
        if reduce_hist is not None:
            red = reduce_hist(hist)
            if red.data_ptr() != hist.data_ptr():
                hist.copy_(red)
        call('adalog_select_scan', _p(ws), R, T, p, _stream())
    out = torch.empty(R, T, dtype=torch.float32, device=x2d.device)
    call('adalog_select_finish', _p(ws), R, T, _p(out), _stream())
    return out


# ------------------------------------------------------------------------------------------ generators
def _rows2d(x):
    if x.dim() != 2 or x.stride(1) != 1:
        raise AdalogError('generator inputs are 2-D with unit K stride')
    return x


def gen_uniform_fixed(x2d, scale_g, zp_g, g_div, g_mod, n_levels, want_rowsum=False, i8=False):
    _cuda(x2d, scale_g, zp_g)
    x2d = _rows2d(_f32(x2d))
    R, K = x2d.shape
    kp = kpad(K, i8)
    out = torch.empty(R, kp, dtype=torch.int8 if i8 else torch.bfloat16, device=x2d.device)
    rowsum = torch.empty(R, dtype=torch.float32, device=x2d.device) if want_rowsum else None
    call('adalog_gen_uniform_fixed', _p(x2d), R, K, x2d.stride(0), _p(_f32(scale_g).reshape(-1)),
         _p(_f32(zp_g).reshape(-1)), int(g_div), int(g_mod), int(n_levels), _p(out), kp, _p(rowsum),
         I8 if i8 else BF16, _stream())
    return out, rowsum


def gen_uniform_cand(x2d, u0, nu, cs, cz, P, pstride, gstride, g_div, g_mod, n_levels, out, krep=1, rowsum=None,
                     i8=False):
    """rows [u0, u0+nu) of x2d -> out[(u*128+p), krep*kpad]  (bf16, or int8 when i8)."""
    _cuda(x2d, cs, cz, out, rowsum)
    K = x2d.shape[1]
    xs = x2d[u0:u0 + nu]
    call('adalog_gen_uniform_cand', _p(xs), nu, K, x2d.stride(0), _p(cs), _p(cz), int(P), int(pstride), int(gstride),
         int(g_div), int(g_mod), int(u0), int(n_levels), _p(out), kpad(K, i8), int(krep), _p(rowsum),
         I8 if i8 else BF16, _stream())


def gen_log_cand(x2d, u0, nu, cs, cq, P, shift, mtab, n_levels, out):
    _cuda(x2d, cs, cq, shift, mtab, out)
    K = x2d.shape[1]
    xs = x2d[u0:u0 + nu]
    call('adalog_gen_log_cand', _p(xs), nu, K, x2d.stride(0), _p(cs), _p(cq), int(P), _p(shift), _p(mtab),
         int(n_levels), _p(out), kpad(K), _stream())


def gen_log_fixed(x2d, scale, q, shift, table1, m2, n_levels):
    _cuda(x2d, scale, q, table1, m2)
    x2d = _rows2d(_f32(x2d))
    R, K = x2d.shape
    kp = kpad(K)
    out = torch.empty(R, kp, dtype=torch.bfloat16, device=x2d.device)
    call('adalog_gen_log_fixed', _p(x2d), R, K, x2d.stride(0), _p(_f32(scale.detach()).reshape(-1)),
         _p(q.detach().to(torch.int64).contiguous()), _p(_f32(shift.detach()).reshape(-1)) if shift is not None else None,
         _p(_f32(table1)), _p(_f32(m2)), int(n_levels), _p(out), kp, _stream())
    return out


def gen_split3(x2d):
    _cuda(x2d)
    x2d = _rows2d(_f32(x2d))
    R, K = x2d.shape
    kp = kpad(K)
    out = torch.empty(R, 3 * kp, dtype=torch.bfloat16, device=x2d.device)
    call('adalog_gen_split3', _p(x2d), R, K, x2d.stride(0), _p(out), kp, _stream())
    return out


# ------------------------------------------------------------------------------------------ candidate GEMM
def pick_bn(N, mult=16):
    """N-tile width: as few tiles as possible, a multiple of `mult` (16 = the UMMA granule; 32 on the W-side / conv
    sweeps, whose epilogue promotes one FP32 partial per absolute 32-token slab to FP64 -- see gemm_err.cu SLAB64)"""
    nt = (N + 255) // 256
    bn = ((math.ceil(N / nt) + mult - 1) // mult) * mult
    return min(256, max(mult, bn))


def cand_gemm_err(A, a_rows, Bm, ka, N, U, UG, brpg, g_base, u_base, y, y_off, ldy, rs, rb, rs_div, rs_mod, cs, cb,
                  upc, S, BN=None, k_true=None, i8=False, split_fast=False):
    """One launch of adalog_cand_gemm_err.  Returns FP64 partial [S, gridX, 128].  ka: operand pitch in elements.
    split_fast only changes which CTAs are co-resident (L2 reuse), never a result bit."""
    BN = BN or pick_bn(N, 32 if rb is not None else 16)
    _cuda(A, Bm, y, rs, rb, cs, cb)
    a = GemmErrArgs()
    a.A, a.Bm = A.data_ptr(), Bm.data_ptr()
    a.a_rows, a.b_rows = int(a_rows), int(Bm.shape[0])
    a.KB = ka // (2 * BK if i8 else BK)
    a.dtype = I8 if i8 else BF16
    a.order = 1 if split_fast else 0
    a.N, a.BN, a.U, a.UG, a.upc, a.S = int(N), int(BN), int(U), int(UG), int(upc), int(S)
    a.brpg, a.g_base, a.u_base = int(brpg), int(g_base), int(u_base)
    a.y, a.ldy = y.data_ptr() + 4 * int(y_off), int(ldy)
    a.rs, a.rb = rs.data_ptr(), (rb.data_ptr() if rb is not None else None)
    a.rs_div, a.rs_mod = int(rs_div), int(rs_mod)
    a.cs, a.cb = (cs.data_ptr() if cs is not None else None), (cb.data_ptr() if cb is not None else None)
    gx = call('adalog_cand_gemm_err_grid', ctypes.byref(a))
    partial = torch.empty(S, gx, P_TILE, dtype=torch.float64, device=Bm.device)
    a.partial = partial.data_ptr()
    if PROFILE['on']:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call('adalog_cand_gemm_err', ctypes.byref(a), _stream())
        e1.record()
        PROFILE['gemm'].append((e0, e1, 2.0 * P_TILE * U * N * (k_true if k_true else ka), bool(i8)))
    else:
        call('adalog_cand_gemm_err', ctypes.byref(a), _stream())
    return partial


def gemm_dequant(A, m_rows, Bm, ka, N, rs, cs, cb, upc, S, i8=False, split_fast=False, out=None):
    """out[m, n] = rs * cs[n] * (A[m, :] . Bm[n, :]) + cb[n] through the tcgen05 pipeline (adalog_gemm_dequant): the
    fake-quant inference forward of a linear layer on the integer parts of Q_a(x) [m_rows, ka] and Q_w(W) [N, ka]."""
    _cuda(A, Bm, rs, cs, cb)
    BN = pick_bn(N)
    U = (m_rows + P_TILE - 1) // P_TILE
    a = GemmErrArgs()
    a.A, a.Bm = A.data_ptr(), Bm.data_ptr()
    a.a_rows, a.b_rows = int(m_rows), int(Bm.shape[0])
    a.KB = ka // (2 * BK if i8 else BK)
    a.dtype = I8 if i8 else BF16
    a.order = 1 if split_fast else 0
    a.N, a.BN, a.U, a.UG, a.upc, a.S = int(N), int(BN), int(U), int(U), int(upc), int(S)
    a.brpg, a.g_base, a.u_base = 0, 0, 0
    a.rs, a.rs_div, a.rs_mod = rs.data_ptr(), 1 << 62, 1
    a.cs, a.cb = cs.data_ptr(), cb.data_ptr()
    if out is None:
        out = torch.empty(m_rows, N, dtype=torch.float32, device=A.device)
    call('adalog_gemm_dequant', ctypes.byref(a), _p(out), int(out.stride(0)), int(m_rows), _stream())
    return out


SMEM_LIMIT = 227 * 1024 - 16384   # dynamic shared memory the fused kernel may ask for (its static part is ~15 KB)


def fused_plan(K, N, i8, log, n_levels):
    """(KB, BN) when the fused generator + GEMM kernel can take the shape (one N tile, K <= 4 blocks, operands fit in
    shared memory), else None (the caller then uses the generator -> workspace -> GEMM path)."""
    el = 2 * BK if i8 else BK
    KB = (K + el - 1) // el
    if N > 256 or KB > 4:
        return None
    BN = pick_bn(N)
    smem = 1024 + KB * 16384 + KB * BN * 128 + (P_TILE * (2 * n_levels + 1) * 4 if log else 0)
    if smem > SMEM_LIMIT or (log and (i8 or 2 * n_levels > 64)):
        return None
    return KB, BN


def fused_cand_gemm_err(x2d, K, Bm, N, U, UG, brpg, y, ldy, rs, rs_div, rs_mod, upc, n_levels, P, cs=None, cz=None,
                        pstride=0, gstride=0, g_div=1, g_mod=1, cq=None, mtab=None, i8=False):
    """One launch of adalog_fused_cand_gemm_err over all U units.  Returns FP64 partial [grid, 128]."""
    log = cq is not None
    KB, BN = fused_plan(K, N, i8, log, n_levels)
    _cuda(x2d, Bm, y, rs)
    assert x2d.dtype == torch.float32 and x2d.stride(1) == 1
    a = FusedArgs()
    a.x, a.ldx = x2d.data_ptr(), int(x2d.stride(0))
    a.Bm, a.b_rows = Bm.data_ptr(), int(Bm.shape[0])
    a.K, a.KB, a.N, a.BN = int(K), int(KB), int(N), int(BN)
    a.U, a.UG, a.upc, a.P, a.n_levels = int(U), int(UG), int(upc), int(P), int(n_levels)
    a.gen, a.dtype = (1 if log else 0), (I8 if i8 else BF16)
    # 14 worker warps: E reduce the error (~20 warp-instructions per column and unit), 14 - E generate candidates
    # (~36 per K element for a uniform quantizer, ~64 for AdaLog); take the better balanced of E = 4 / 8
    ck = (64 if log else 36) * K
    a.epi_warps = min((4, 8), key=lambda e: max(20.0 * N / e, ck / (14 - e)))
    a.brpg, a.g_base, a.u_base = int(brpg), 0, 0
    if log:
        a.cq, a.mtab = cq.data_ptr(), mtab.data_ptr()
    else:
        a.cs, a.cz = cs.data_ptr(), cz.data_ptr()
        a.pstride, a.gstride = int(pstride), int(gstride)
    a.g_div, a.g_mod = int(g_div), int(g_mod)
    a.y, a.ldy = y.data_ptr(), int(ldy)
    a.rs, a.rs_div, a.rs_mod = rs.data_ptr(), int(rs_div), int(rs_mod)
    grid = call('adalog_fused_cand_gemm_err_grid', ctypes.byref(a))
    partial = torch.empty(grid, P_TILE, dtype=torch.float64, device=Bm.device)
    a.partial = partial.data_ptr()
    if PROFILE['on']:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call('adalog_fused_cand_gemm_err', ctypes.byref(a), _stream())
        e1.record()
        PROFILE['fused'].append((e0, e1, 2.0 * P_TILE * U * N * K))
    else:
        call('adalog_fused_cand_gemm_err', ctypes.byref(a), _stream())
    return partial


def lin_fused_cand_gemm_err(x2d, Bm, N, y, rs, ccs, ccb, n_levels, P, cs, cz=None, cq=None, shift=None, mtab=None,
                            i8=False, max_passes=None):
    """One launch of adalog_lin_fused_cand_gemm_err over all tokens of a linear activation sweep.  Returns the FP64
    partial [grid, 128], or None when no schedule fits in shared memory or the schedule would generate each unit's
    operand more than `max_passes` times (caller: two-kernel path)."""
    _cuda(x2d, Bm, y, rs, ccs, ccb, cs, cz, cq, shift, mtab)
    assert x2d.dtype == torch.float32 and x2d.stride(1) == 1 and y.stride(1) == 1
    log = cq is not None
    a = LinFusedArgs()
    a.x, a.ldx = x2d.data_ptr(), int(x2d.stride(0))
    a.Bm, a.b_rows = Bm.data_ptr(), int(Bm.shape[0])
    a.K, a.N, a.U = int(x2d.shape[1]), int(N), int(x2d.shape[0])
    a.P, a.n_levels = int(P), int(n_levels)
    a.gen, a.dtype = (1 if log else 0), (I8 if i8 else BF16)
    a.cs = cs.data_ptr()
    if log:
        a.cq, a.mtab = cq.data_ptr(), mtab.data_ptr()
        a.shift = shift.data_ptr() if shift is not None else None
    else:
        a.cz = cz.data_ptr()
    a.y, a.ldy = y.data_ptr(), int(y.stride(0))
    a.rs, a.ccs, a.ccb = rs.data_ptr(), ccs.data_ptr(), ccb.data_ptr()
    lib = _lib.load()
    grid = lib.adalog_lin_fused_cand_gemm_err_grid(ctypes.byref(a))
    if grid == -3 or grid == -2:
        return None
    if max_passes is not None and grid > 0 and lib.adalog_lin_fused_cand_gemm_err_passes(ctypes.byref(a)) > max_passes:
        return None
    if grid < 0:
        raise AdalogError(f'adalog_lin_fused_cand_gemm_err_grid failed ({grid}): {lib.adalog_last_error().decode()}')
    partial = torch.empty(grid, P_TILE, dtype=torch.float64, device=Bm.device)
    a.partial = partial.data_ptr()
    if PROFILE['on']:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call('adalog_lin_fused_cand_gemm_err', ctypes.byref(a), _stream())
        e1.record()
        PROFILE['lin'].append((e0, e1, 2.0 * P_TILE * a.U * N * a.K, bool(i8)))
    else:
        call('adalog_lin_fused_cand_gemm_err', ctypes.byref(a), _stream())
    return partial


def profile_lin_summary():
    """{'bf16': (ops, ms, launches), 'i8': (...)} of the fused linear activation sweeps since profile_reset(True)"""
    torch.cuda.synchronize()
    def tot(rows):
        return sum(r[2] for r in rows), sum(r[0].elapsed_time(r[1]) for r in rows), len(rows)
    return {'bf16': tot([r for r in PROFILE['lin'] if not r[3]]), 'i8': tot([r for r in PROFILE['lin'] if r[3]])}


def profile_fused_summary():
    """(flops, milliseconds, launches) of the fused generator + GEMM launches recorded since profile_reset(True)"""
    torch.cuda.synchronize()
    return (sum(f for _, _, f in PROFILE['fused']), sum(e0.elapsed_time(e1) for e0, e1, _ in PROFILE['fused']),
            len(PROFILE['fused']))


def debug_gemm_tile(A, Bm):
    """D[128, N] = A[128, ka] @ Bm[N, ka]^T through the tcgen05 pipeline (test hook); bf16 or int8 operands."""
    _cuda(A, Bm)
    assert A.dtype == Bm.dtype and A.dtype in (torch.bfloat16, torch.int8) and A.shape[0] == 128
    i8 = A.dtype == torch.int8
    ka = A.shape[1]
    N = Bm.shape[0]
    D = torch.zeros(128, N, dtype=torch.float32, device=A.device)
    zeros = torch.zeros(max(N, 1), dtype=torch.float32, device=A.device)
    ones = torch.ones(P_TILE, dtype=torch.float32, device=A.device)
    partial = torch.empty(64 * P_TILE, dtype=torch.float64, device=A.device)
    call('adalog_debug_gemm_tile', _p(A.contiguous()), _p(Bm.contiguous()), ka // (2 * BK if i8 else BK), N, _p(D),
         _p(zeros), _p(ones), _p(partial), I8 if i8 else BF16, _stream())
    return D


def version():
    return _lib.load().adalog_version()
