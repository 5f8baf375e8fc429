"""ODE-module protocol shared by the bundled membrane models.

A membrane model is a Python module exposing (consumed at
src/knpemidg/membrane.py:32-34, 88 and src/knpemidg/solver.py:248, 257, 1098):

    init_state_values(**overrides)      -> float64[ns]
    init_parameter_values(**overrides)  -> float64[np]
    state_indices(*names), parameter_indices(*names) -> int | list[int]
    rhs_numba                           -> callable (t, y, dy, p); numba cfunc when
                                           numba is importable (`.address`), the
                                           Python function is kept at `.py_func`

`build` creates those from two (name, default) tables and one Python
right-hand side.  The same Python source is what knpemidg.odegen translates to
the CUDA device function, so a bundled model and a user's Gotran/numba module
go through one mechanism.
"""
import numpy as np


def _indexer(table, what):
    index = {name: i for i, (name, _) in enumerate(table)}

    def indices(*names):
        out = []
        for name in names:
            if name not in index:
                raise ValueError("Unknown {0}: '{1}'".format(what, name))
            out.append(index[name])
        return out if len(out) > 1 else out[0]

    return indices, index


def _initialiser(table, index, what):
    def init(**values):
        arr = np.array([v for _, v in table], dtype=np.float64)
        for name, value in values.items():
            if name not in index:
                raise ValueError("{0} is not a {1}.".format(name, what))
            arr[index[name]] = value
        return arr

    return init


class _LazyCfunc:
    """numba cfunc compiled on first use of `.address` (import stays cheap)."""

    def __init__(self, py_func):
        self.py_func = py_func
        self._c = None

    @property
    def address(self):
        if self._c is None:
            from numba import cfunc
            from numbalsoda import lsoda_sig
            self._c = cfunc(lsoda_sig, nopython=True)(self.py_func)
        return self._c.address

    def __call__(self, t, y, dy, p):
        return self.py_func(t, y, dy, p)


def build(module_name, states, parameters, rhs):
    s_idx, s_map = _indexer(states, "state")
    p_idx, p_map = _indexer(parameters, "param")
    return {
        "init_state_values": _initialiser(states, s_map, "state"),
        "init_parameter_values": _initialiser(parameters, p_map, "parameter"),
        "state_indices": s_idx,
        "parameter_indices": p_idx,
        "rhs_numba": _LazyCfunc(rhs),
        "STATE_NAMES": tuple(n for n, _ in states),
        "PARAMETER_NAMES": tuple(n for n, _ in parameters),
    }
