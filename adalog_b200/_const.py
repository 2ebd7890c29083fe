"""Device-resident copies of the small host-computed constant tensors of the search (linspace ramps, zero-point
ranges, base grids, percentile pairs, LUT numerators).

The reference builds each of them on the CPU and moves it with `.cuda()` inside the search loops (linear.py:442-451,
:497-502, :861-863; matmul.py:231-240, :323-324).  Bits must stay the CPU-computed ones, so they are still built on the
CPU -- but once per (value, device): a host-to-device copy from pageable memory first waits for the stream to drain, so
one such copy per evaluation serialised the CPU's launch work with the GPU's execution (measured: 15% of a
calibration step with the GPU idle, 0.45 ms x 3150 evaluations).  Callers must not modify the returned tensors in place.
"""
import functools

import torch


def _key(device):
    device = torch.device(device)
    return device.type, device.index


@functools.lru_cache(maxsize=None)
def _linspace01(n, dtype_key):
    return torch.linspace(0, 1, steps=n).to(torch.device(*dtype_key) if dtype_key[1] is not None else dtype_key[0])


def linspace01(n, device):
    """torch.linspace(0, 1, steps=n) computed on the CPU (FP32), resident on `device`"""
    return _linspace01(int(n), _key(device))


@functools.lru_cache(maxsize=None)
def _linspace01_centered(n, dtype_key):
    return _linspace01(n, dtype_key) - 0.5       # on the device, as the reference does (linear.py:497-502)


def linspace01_centered(n, device):
    """linspace01(n) - 0.5 (the refinement offsets of FPCS before scaling by the grid pitch), cached per device"""
    return _linspace01_centered(int(n), _key(device))


@functools.lru_cache(maxsize=None)
def _int_range(lo, hi, dk):
    return torch.tensor(range(lo, hi)).to(torch.device(*dk) if dk[1] is not None else dk[0])


def int_range(lo, hi, device):
    """torch.tensor(range(lo, hi)) (int64)"""
    return _int_range(int(lo), int(hi), _key(device))


@functools.lru_cache(maxsize=None)
def _ramp(n, dk):
    return torch.tensor([i / (n - 1) for i in range(n)]).to(torch.device(*dk) if dk[1] is not None else dk[0])


def ramp(n, device):
    """torch.tensor([i / (n - 1) for i in range(n)]): Python float64 quotients rounded to FP32 (linear.py:811-813, :950)"""
    return _ramp(int(n), _key(device))


@functools.lru_cache(maxsize=None)
def _floats(values, dk):
    return torch.tensor(list(values)).to(torch.device(*dk) if dk[1] is not None else dk[0])


def floats(values, device):
    """torch.tensor([...]) of Python numbers (default dtype inference), e.g. the percentile pair [l, r]"""
    return _floats(tuple(values), _key(device))


@functools.lru_cache(maxsize=None)
def _pct_pair(l, r, dk):
    pct = torch.tensor([l, r])
    return torch.cat([pct, 1 - pct]).to(torch.device(*dk) if dk[1] is not None else dk[0])


def pct_pair(l, r, device):
    """cat([pct, 1 - pct]) for pct = tensor([l, r]) in FP32 (1 - 0.9 is 0.10000002384185791)"""
    return _pct_pair(float(l), float(r), _key(device))


@functools.lru_cache(maxsize=None)
def _search_table_ints(n_levels, dk):
    table = torch.tensor([2 ** (-j / 37.0) for j in range(120)])
    table_scale = 1. / (4 * n_levels - 2)
    return torch.round(table / table_scale)[:37].contiguous().to(torch.device(*dk) if dk[1] is not None else dk[0])


def search_table_ints(n_levels, device):
    """integer numerators of the 37 live entries of the search LUT (linear.py:750-752 / matmul.py:313-315)"""
    return _search_table_ints(int(n_levels), _key(device))
