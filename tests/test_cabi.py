"""The C-ABI library loads and exports every symbol include/adalog_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

from adalog_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'adalog_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(adalog_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_header_symbols():
    syms = declared_symbols()
    assert len(syms) >= 14
    assert os.path.exists(_lib.LIB_PATH), 'run `python __graft_entry__.py` (build()) first'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/adalog_b200.h but not exported'


def test_binding_covers_header():
    syms = set(declared_symbols())
    bound = set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS)
    assert syms == bound, (syms - bound, bound - syms)
    lib = _lib.load()
    assert lib.adalog_version() >= 100


def test_struct_layout_matches_header():
    # adalog_gemm_err_args: 2 ptr + 2 i64 + 10 i32 + 3 i64 + ptr + i64 + 2 ptr + 2 i64 + 2 ptr + ptr
    assert ctypes.sizeof(_lib.GemmErrArgs) == 8 * 4 + 4 * 10 + 8 * 3 + 8 * 2 + 8 * 2 + 8 * 2 + 8 * 2 + 8


def test_integration_stub_struct_matches_binding():
    """the ctypes stub shown to reference maintainers in INTEGRATION.md lists the fields of adalog_gemm_err_args in the
    same order as the tested binding"""
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = text[text.index('class _GemmErrArgs'):text.index('def _check')]
    names = re.findall(r"\('(\w+)',\s*ctypes\.", block)
    assert names == [f[0] for f in _lib.GemmErrArgs._fields_]


def test_integration_lin_fused_stub_matches_binding():
    """same for the adalog_lin_fused_args stub"""
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    block = text[text.index('class _LinFusedArgs'):text.index('a = _LinFusedArgs(')]
    names = re.findall(r"\('(\w+)',\s*ctypes\.", block)
    assert names == [f[0] for f in _lib.LinFusedArgs._fields_]

