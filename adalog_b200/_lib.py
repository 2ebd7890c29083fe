"""ctypes binding of libadalog_b200.so (include/adalog_b200.h).

There is no CPU fallback: if the shared library is missing or a tensor is not on a CUDA device the
call raises.  Build the library with `python __graft_entry__.py` (or `make -C adalog_b200/csrc`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libadalog_b200.so')

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_vp = ctypes.c_void_p


class GemmErrArgs(ctypes.Structure):
    """mirror of adalog_gemm_err_args"""
    _fields_ = [
        ('A', c_vp), ('Bm', c_vp), ('a_rows', c_i64), ('b_rows', c_i64),
        ('KB', ctypes.c_int32), ('N', ctypes.c_int32), ('BN', ctypes.c_int32), ('U', ctypes.c_int32),
        ('UG', ctypes.c_int32), ('upc', ctypes.c_int32), ('S', ctypes.c_int32), ('dtype', ctypes.c_int32),
        ('order', ctypes.c_int32), ('reserved', ctypes.c_int32),
        ('brpg', c_i64), ('g_base', c_i64), ('u_base', c_i64),
        ('y', c_vp), ('ldy', c_i64),
        ('rs', c_vp), ('rb', c_vp), ('rs_div', c_i64), ('rs_mod', c_i64),
        ('cs', c_vp), ('cb', c_vp),
        ('partial', c_vp),
    ]


class FusedArgs(ctypes.Structure):
    """mirror of adalog_fused_args"""
    _fields_ = [
        ('x', c_vp), ('ldx', c_i64), ('Bm', c_vp), ('b_rows', c_i64),
        ('K', ctypes.c_int32), ('KB', ctypes.c_int32), ('N', ctypes.c_int32), ('BN', ctypes.c_int32),
        ('U', ctypes.c_int32), ('UG', ctypes.c_int32), ('upc', ctypes.c_int32),
        ('P', ctypes.c_int32), ('n_levels', ctypes.c_int32),
        ('gen', ctypes.c_int32), ('dtype', ctypes.c_int32), ('epi_warps', ctypes.c_int32),
        ('brpg', c_i64), ('g_base', c_i64), ('u_base', c_i64),
        ('cs', c_vp), ('cz', c_vp), ('pstride', c_i64), ('gstride', c_i64), ('g_div', c_i64), ('g_mod', c_i64),
        ('cq', c_vp), ('mtab', c_vp),
        ('y', c_vp), ('ldy', c_i64),
        ('rs', c_vp), ('rs_div', c_i64), ('rs_mod', c_i64),
        ('partial', c_vp),
    ]


class LinFusedArgs(ctypes.Structure):
    """mirror of adalog_lin_fused_args"""
    _fields_ = [
        ('x', c_vp), ('ldx', c_i64), ('Bm', c_vp), ('b_rows', c_i64),
        ('K', ctypes.c_int32), ('N', ctypes.c_int32), ('U', ctypes.c_int32),
        ('P', ctypes.c_int32), ('n_levels', ctypes.c_int32),
        ('gen', ctypes.c_int32), ('dtype', ctypes.c_int32), ('reserved', ctypes.c_int32),
        ('cs', c_vp), ('cz', c_vp), ('cq', c_vp), ('shift', c_vp), ('mtab', c_vp),
        ('y', c_vp), ('ldy', c_i64),
        ('rs', c_vp), ('ccs', c_vp), ('ccb', c_vp),
        ('partial', c_vp),
    ]


# name -> argtypes; every function returns int except adalog_last_error
SIGNATURES = {
    'adalog_version': [],
    'adalog_uniform_fakequant_f32': [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp],
    'adalog_log_fakequant_f32': [c_vp, c_vp, c_vp, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp],
    'adalog_twin_fakequant_f32': [c_vp, c_vp, c_i64, c_vp, c_int, c_vp],
    'adalog_sweep_err_w_self': [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    'adalog_sweep_err_a_self': [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_vp],
    'adalog_gen_uniform_fixed': [c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_int, c_vp, c_int, c_vp],
    'adalog_gen_uniform_cand': [c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_int,
                                c_vp, c_int, c_int, c_vp, c_int, c_vp],
    'adalog_gen_log_cand': [c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_vp],
    'adalog_gen_log_fixed': [c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp],
    'adalog_gen_split3': [c_vp, c_i64, c_int, c_i64, c_vp, c_int, c_vp],
    'adalog_select_init': [c_vp, c_i64, c_int, c_vp, c_vp],
    'adalog_select_hist': [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_vp],
    'adalog_select_scan': [c_vp, c_i64, c_int, c_int, c_vp],
    'adalog_select_finish': [c_vp, c_i64, c_int, c_vp, c_vp],
    'adalog_cand_gemm_err_grid': [ctypes.POINTER(GemmErrArgs)],
    'adalog_cand_gemm_err': [ctypes.POINTER(GemmErrArgs), c_vp],
    'adalog_fused_cand_gemm_err_grid': [ctypes.POINTER(FusedArgs)],
    'adalog_fused_cand_gemm_err': [ctypes.POINTER(FusedArgs), c_vp],
    'adalog_lin_fused_cand_gemm_err_grid': [ctypes.POINTER(LinFusedArgs)],
    'adalog_lin_fused_cand_gemm_err_passes': [ctypes.POINTER(LinFusedArgs)],
    'adalog_lin_fused_cand_gemm_err': [ctypes.POINTER(LinFusedArgs), c_vp],
    'adalog_gemm_dequant': [ctypes.POINTER(GemmErrArgs), c_vp, c_i64, c_i64, c_vp],
    'adalog_debug_gemm_tile': [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp],
}

# exported functions that do not return an int status (bound separately in load())
OTHER_SYMBOLS = {'adalog_last_error', 'adalog_select_workspace_bytes', 'adalog_select_hist_ptr'}

_lib = None
# kernels launched through this binding (bench.py reports it as gpu_launches)
LAUNCHES = {'count': 0}
_NO_LAUNCH = {'adalog_version', 'adalog_cand_gemm_err_grid', 'adalog_fused_cand_gemm_err_grid',
              'adalog_lin_fused_cand_gemm_err_grid', 'adalog_lin_fused_cand_gemm_err_passes'}


class AdalogError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdalogError(f'{LIB_PATH} not found: build it with `python __graft_entry__.py` '
                          f'(adalog_b200 has no CPU fallback)')
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    lib.adalog_last_error.argtypes = []
    lib.adalog_last_error.restype = ctypes.c_char_p
    lib.adalog_select_workspace_bytes.argtypes = [c_i64, c_int]
    lib.adalog_select_workspace_bytes.restype = ctypes.c_int64
    lib.adalog_select_hist_ptr.argtypes = [c_vp, c_i64, c_int]
    lib.adalog_select_hist_ptr.restype = ctypes.c_void_p
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if name not in _NO_LAUNCH:
        LAUNCHES['count'] += 1
    if rc < 0:
        raise AdalogError(f'{name} failed ({rc}): {lib.adalog_last_error().decode()}')
    return rc
