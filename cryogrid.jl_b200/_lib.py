"""ctypes binding of libcgb200.so (include/cgb200.h). Fails loudly when the CUDA library is missing:
there is no CPU fallback in the product path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CGB_LIB_PATH") or os.path.join(HERE, "lib", "libcgb200.so")   # override: A/B builds only

ABI_VERSION = 1
OK = 0
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_ALLOC = -1, -2, -3, -4, -5
PHYSICS_HEAT, PHYSICS_HEAT_SALT, PHYSICS_HEAT_T = 0, 1, 2
STEPPER_EULER, STEPPER_LITE_IMPLICIT = 0, 1
LIMITER_MAXDELTA, LIMITER_CFL = 0, 1
BC_DIRICHLET, BC_NEUMANN = 0, 1
TOP, BOTTOM = 0, 1
FC_FREEWATER, FC_DALLAMICO, FC_PAINTERKARRA, FC_DALLAMICOSALT = 0, 1, 2, 3
SOLVER_LUT, SOLVER_NEWTON = 0, 1
EXTRAP_FLAT, EXTRAP_LINE = 0, 1
PER_GLOBAL, PER_LAYER, PER_COLUMN, PER_COLUMN_LAYER, PER_CELL, PER_CELL_COLUMN = 0, 1, 2, 3, 4, 5
STATUS_MAXITERS, STATUS_NAN, STATUS_FORCING_RANGE, STATUS_BAD_DT = 1, 2, 4, 8


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("n_cells", C.c_int32), ("n_layers", C.c_int32),
        ("n_columns", C.c_int64),
        ("physics", C.c_int32), ("stepper", C.c_int32), ("limiter", C.c_int32), ("top_kind", C.c_int32),
        ("bot_kind", C.c_int32), ("use_nfactor", C.c_int32),
        ("maxdelta", C.c_double), ("courant", C.c_double), ("dt0", C.c_double), ("dtmin", C.c_double),
        ("dtmax", C.c_double),
        ("lite_miniters", C.c_int32), ("lite_maxiters", C.c_int32), ("lite_tol", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_columns", C.c_int64), ("n_cells", C.c_int64), ("column_steps", C.c_int64), ("cell_steps", C.c_int64),
        ("lite_iterations", C.c_int64), ("kernel_launches", C.c_int64), ("min_steps", C.c_int64),
        ("max_steps", C.c_int64), ("min_dt", C.c_double), ("max_dt", C.c_double), ("status_or", C.c_int32),
        ("_pad", C.c_int32),
    ]


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)

# every symbol include/cgb200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cgb_abi_version": (C.c_int, []),
    "cgb_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "cgb_destroy": (None, [_P]),
    "cgb_last_error": (C.c_char_p, [_P]),
    "cgb_set_grid": (C.c_int, [_P, _D]),
    "cgb_set_layers": (C.c_int, [_P, _I32]),
    "cgb_set_param": (C.c_int, [_P, C.c_char_p, _D, C.c_int64, C.c_int32]),
    "cgb_set_freezecurve": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "cgb_set_sfcc_table": (C.c_int, [_P, C.c_int32, _D, _D, C.c_int64, C.c_int32]),
    "cgb_build_sfcc_tables": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_int32, _I32]),
    "cgb_get_sfcc_table": (C.c_int, [_P, C.c_int32, _D, _D, C.c_int64, _I64]),
    "cgb_set_sfcc_table_map": (C.c_int, [_P, _I32]),
    "cgb_set_forcing_table": (C.c_int, [_P, C.c_int32, _D, _D, C.c_int64]),
    "cgb_set_forcing_periodic": (C.c_int, [_P, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]),
    "cgb_set_forcing_constant": (C.c_int, [_P, C.c_int32, C.c_double]),
    "cgb_set_state": (C.c_int, [_P, _D, C.c_double]),
    "cgb_restore_state": (C.c_int, [_P, _D, _D, _D, _I64]),
    "cgb_init_interp": (C.c_int, [_P, _D, _D, C.c_int64, C.c_int32, C.c_double]),
    "cgb_init_steadystate": (C.c_int, [_P, _D, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_double]),
    "cgb_get_state": (C.c_int, [_P, _D, _D, _D]),
    "cgb_get_status": (C.c_int, [_P, _I32, _I64]),
    "cgb_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "cgb_rhs": (C.c_int, [_P, C.c_double, _D]),
    "cgb_advance_to": (C.c_int, [_P, C.c_double]),
    "cgb_synchronize": (C.c_int, [_P]),
    "cgb_get_diag": (C.c_int, [_P, C.c_char_p, C.c_double, _D]),
    "cgb_solve_saveat": (C.c_int, [_P, _D, C.c_int64, C.c_char_p, _D]),
    "cgb_solve_saveat_select": (C.c_int, [_P, _D, C.c_int64, C.c_char_p, _I32, C.c_int32, _D, _D]),
    "cgb_solve_saveat_permafrost": (C.c_int, [_P, _D, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _D, _D, _D, _D, _D]),
    "cgb_ensemble_partial": (C.c_int, [_P, C.c_char_p, C.c_double, _D, _D]),
    "cgb_device_state": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "cgb_ensemble_partial_device": (C.c_int, [_P, C.c_char_p, C.c_double, _P]),
    "cgb_profile_enable": (C.c_int, [_P, C.c_int32]),
    "cgb_profile_read": (C.c_int, [_P, _I64, C.POINTER(C.c_double)]),
    "cgb_selftest_rcp": (C.c_int, [_D, C.c_int64, _D, _D, _D]),
    "cgb_selftest_sfcc_curve": (C.c_int, [_P, _D, C.c_int64, _D]),
    "cgb_selftest_explog": (C.c_int, [_D, C.c_int64, _D, _D]),
    "cgb_selftest_explog_tab": (C.c_int, [_D, C.c_int64, _D, _D]),
}

_lib = None


class CgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libcgb200 error {code}: {msg}")
        self.code = code


def load():
    """Load libcgb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "cryogrid.jl_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a):
    return a.ctypes.data_as(_D)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
