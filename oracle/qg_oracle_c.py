"""ctypes wrapper of oracle/qg_oracle.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libqgoracle.so")


class qgo_params(C.Structure):
    _fields_ = [("M", C.c_int), ("P", C.c_int), ("dx", C.c_double), ("dt", C.c_double), ("visc", C.c_double),
                ("r", C.c_double), ("U", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("alpha", C.c_double), ("Pinv", C.c_double * 4), ("Pfwd", C.c_double * 4), ("H1", C.c_double),
                ("H2", C.c_double), ("S1", C.c_double)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            import subprocess
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.qgo_step.argtypes = [C.POINTER(qgo_params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.qgo_step.restype = None
        _lib.qgo_max_threads.restype = C.c_int
    return _lib


def params_of(m):
    """m: oracle.qg_oracle.BaroclinicModel"""
    import qg_oracle as o
    p = qgo_params()
    p.M, p.P, p.dx, p.dt, p.visc, p.r, p.U = m.M, m.P, m.dx, m.dt, m.visc, m.r, m.U
    p.beta1, p.beta2, p.alpha = o.beta_1(m), o.beta_2(m), o.S_eig(m)
    p.Pinv = (C.c_double * 4)(*o.P_inv_matrix(m).ravel())
    p.Pfwd = (C.c_double * 4)(*o.P_matrix(m.H_1, m.H_1).ravel())
    p.H1, p.H2, p.S1 = m.H_1, m.H_2, o.S1_plus(m)
    return p


def run_steps(m, zeta, psi, f_store, first_timestep, nsteps, nthreads=0):
    lib = load()
    for a in (zeta, psi, f_store):
        assert a.flags.f_contiguous and a.dtype == np.float64
    p = params_of(m)
    lib.qgo_step(C.byref(p), zeta.ctypes.data, psi.ctypes.data, f_store.ctypes.data, first_timestep, nsteps, nthreads)


def max_threads():
    return int(load().qgo_max_threads())
