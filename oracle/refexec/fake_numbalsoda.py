"""ORACLE (test infrastructure only - never imported by the product path).

`numbalsoda` for the reference's membrane.py:108-112: `lsoda(funcptr, u0, t_eval, data, rtol,
atol)` integrates the numba cfunc right-hand side (called through its C address, exactly as
numbalsoda does) from t_eval[0] to t_eval[-1].  numbalsoda (unpinned, pyproject.toml:14) wraps a
C++ port of ODEPACK's LSODA; scipy's LSODA is the Fortran original: same method, same error
control (rtol, atol = 0 -> floored at 1e-300 because scipy rejects 0).  The right-hand side
writes the channel currents into `data` as a side effect (e.g. mm_hh.py:154-159), so after the
call `data` holds the values of the integrator's LAST right-hand-side evaluation - the
reference's convention for the currents handed to the PDEs.
"""
import ctypes

import numpy as np
from numba import types as _t
from scipy.integrate import solve_ivp

lsoda_sig = _t.void(_t.double, _t.CPointer(_t.double), _t.CPointer(_t.double), _t.CPointer(_t.double))
_PROTO = ctypes.CFUNCTYPE(None, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_double))
_DP = ctypes.POINTER(ctypes.c_double)


def lsoda(funcptr, u0, t_eval, data=None, rtol=1e-3, atol=1e-6):
    f = _PROTO(funcptr)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    assert data.flags["C_CONTIGUOUS"] and data.dtype == np.float64
    pd = data.ctypes.data_as(_DP)
    dy = np.empty_like(u0)

    def rhs(t, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        f(float(t), y.ctypes.data_as(_DP), dy.ctypes.data_as(_DP), pd)
        return dy.copy()

    sol = solve_ivp(rhs, (float(t_eval[0]), float(t_eval[-1])), u0, method="LSODA", rtol=rtol,
                    atol=atol if atol > 0 else 1e-300, t_eval=np.asarray(t_eval, dtype=float))
    return sol.y.T.copy(), bool(sol.success)
