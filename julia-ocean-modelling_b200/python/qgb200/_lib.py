"""ctypes binding of libqgb200.so (include/qgb200.h).

This is the Python twin of the Julia ``ccall`` shim in ``julia/src/model.jl``: the same
symbols, the same argument order.  There is no fallback of any kind: if the shared library
is missing or no B200-class GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QGB200_LIB points at another build of the same library (e.g. one compiled with -DQG_K3_TRACE)
LIB_PATH = os.environ.get("QGB200_LIB") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib", "libqgb200.so"))

QG_NKERNELS = 8


class QGError(RuntimeError):
    """A libqgb200 call returned a negative status."""

    def __init__(self, code, message):
        super().__init__(f"libqgb200 error {code}: {message}")
        self.code = code
        self.message = message


class qg_params(C.Structure):
    _fields_ = [("M", C.c_int32), ("P", C.c_int32), ("dx", C.c_double), ("dt", C.c_double),
                ("visc", C.c_double), ("r", C.c_double), ("U", C.c_double), ("beta1", C.c_double),
                ("beta2", C.c_double), ("alpha", C.c_double), ("Pinv", C.c_double * 4),
                ("Pfwd", C.c_double * 4), ("H1", C.c_double), ("H2", C.c_double), ("S1", C.c_double)]


# every symbol declared in include/qgb200.h: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
SYMBOLS = {
    "qg_abi_version": (C.c_int, []),
    "qg_last_error": (C.c_char_p, [_P]),
    "qg_create": (C.c_int, [C.POINTER(qg_params), C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "qg_destroy": (C.c_int, [_P]),
    "qg_upload_state": (C.c_int, [_P, _P, _P, _P]),
    "qg_upload_initial_state": (C.c_int, [_P, _P, _P]),
    "qg_init_state": (C.c_int, [_P, C.c_uint64, C.c_double, C.c_double, C.c_double]),
    "qg_download_state": (C.c_int, [_P, _P, _P, _P]),
    "qg_snapshot_begin": (C.c_int, [_P, _P, _P]),
    "qg_snapshot_end": (C.c_int, [_P]),
    "qg_evolve_zeta": (C.c_int, [_P, C.c_int]),
    "qg_evolve_psi": (C.c_int, [_P]),
    "qg_step": (C.c_int, [_P, C.c_int, C.c_int]),
    "qg_sync": (C.c_int, [_P]),
    "qg_diagnostics": (C.c_int, [_P, _D, _D]),
    "qg_extrema": (C.c_int, [_P, _D]),
    "qg_solve": (C.c_int, [_P, C.c_int, _P, _P]),
    "qg_set_profiling": (C.c_int, [_P, C.c_int]),
    "qg_kernel_times": (C.c_int, [_P, _D, C.POINTER(C.c_int64)]),
    "qg_kernel_name": (C.c_char_p, [C.c_int]),
    "qg_launch_count": (C.c_int64, [_P]),
    "qg_plan_probe": (C.c_int, [C.POINTER(qg_params), C.c_int, _D, _D, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int)]),
    "qg_nccl_unique_id": (C.c_int, [_P]),
    "qg_dist_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "qg_dist_ipc_export": (C.c_int, [_P, _P]),
    "qg_dist_ipc_import": (C.c_int, [_P, _P]),
    "qg_dist_ipc_blobs_share_device": (C.c_int, [_P, C.c_int]),
    "qg_device_layout": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """Load libqgb200.so (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C julia-ocean-modelling_b200/csrc`.  qgb200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, rc):
    if rc != 0:
        msg = load().qg_last_error(handle)
        raise QGError(rc, msg.decode() if msg else "unknown error")
