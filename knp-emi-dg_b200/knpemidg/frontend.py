"""Light stand-ins for the dolfin objects that cross the reference's Solver API.

The reference passes dolfin `Constant`s, `Function`s and UFL expressions between the
run scripts, `Solver` and `MembraneModel`.  On the B200 path all fields
live on the device, so these classes are *handles*: they name a device field and fetch
or store it on demand.  Only what the run scripts and the `update_ode` hooks use is
provided (`Constant.assign/float`, `Function.vector().get_local()`, `c_prev_k.split()`,
`plus/minus/pcws_constant_project` results).
"""
from __future__ import annotations

import numpy as np

from . import _lib


class Constant:
    """dolfin.Constant look-alike (scalar or small vector).

    A scalar Constant is also a SYMBOL: arithmetic with it (or passing it to the symbolic
    sin/cos/... of knpemidg.symbolic) yields a sympy expression in which the Constant stays
    a free symbol whose CURRENT value is substituted whenever the expression is evaluated -
    that is how the time `t` of the reference's manufactured solutions (tests/mms_time.py:28-43)
    keeps following `t.assign`.  `float(c)` and `as_float(expr)` give numbers."""

    def __init__(self, value):
        self._v = np.asarray(float(value) if np.ndim(value) == 0 else value, dtype=float)
        self._sym = None

    def assign(self, value):
        self._v = np.asarray(float(value) if np.ndim(value) == 0 else value, dtype=float)

    def values(self):
        return np.atleast_1d(self._v)

    def __float__(self):
        return float(self._v)

    def _sympy_(self):
        if self._v.ndim != 0:
            raise TypeError("a vector Constant is not a scalar symbol")
        if self._sym is None:
            from . import symbolic
            self._sym = symbolic.constant_symbol(self)
        return self._sym

    def _bin(self, o, f):
        from . import symbolic
        if isinstance(o, symbolic.Vec):
            return NotImplemented
        return f(self._sympy_(), symbolic.to_sym(o))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: b / a)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __neg__(self): return -self._sympy_()
    def __pos__(self): return self._sympy_()

    def __repr__(self):
        return f"Constant({self._v})"


def as_float(v):
    """float of a number, a Constant (ours or dolfin's), a 0-d array, or a symbolic expression
    of Constants (their current values)"""
    try:
        return float(v)
    except TypeError:
        from . import symbolic
        return symbolic.numeric(v)


class _Vector:
    def __init__(self, getter, setter):
        self._get, self._set = getter, setter

    def get_local(self):
        return self._get()

    def set_local(self, values):
        self._set(np.asarray(values, dtype=float))

    def apply(self, mode="insert"):
        pass

    def __len__(self):
        return len(self._get())


class FunctionSpace:
    """Names a space: 'DG1' (cell field, nd dofs per cell) or 'DLT0' (one value per
    membrane facet).  Carries the engine so that MembraneModel can find the device."""

    def __init__(self, engine, kind, mesh=None):
        self.engine, self.kind, self._mesh = engine, kind, mesh

    def mesh(self):
        return self._mesh

    def dim(self):
        return self.engine.n if self.kind == "DG1" else self.engine.nm


from .symbolic import field_ops  # noqa: E402


@field_ops
class CellField:
    """DG-P1 field on the device: field id (which, idx) of include/knpemi.h."""

    def __init__(self, engine, which, idx=0, name="f"):
        self.engine, self.which, self.idx, self.name = engine, which, idx, name

    def nodal(self):
        """[nc, nd] nodal values (dof = nd*cell + local vertex)"""
        return self.engine.ctx.get_field(self.which, self.idx).reshape(self.engine.nc, self.engine.nd)

    def vector(self):
        return _Vector(lambda: self.engine.ctx.get_field(self.which, self.idx),
                       lambda v: self.engine.ctx.set_field(self.which, self.idx, v))

    def assign(self, other):
        if isinstance(other, CellField):
            self.engine.ctx.set_field(self.which, self.idx, other.vector().get_local())
        else:
            self.engine.ctx.set_field(self.which, self.idx, np.asarray(other, dtype=float))

    def function_space(self):
        return FunctionSpace(self.engine, "DG1", self.engine.mesh)

    def __call__(self, x):
        """point evaluation (host side, for post-processing only)"""
        return self.engine.evaluate(self, np.asarray(x, dtype=float))


class MixedCellField:
    """c / c_prev_k / c_prev_n of the reference: the solved ions as one object."""

    def __init__(self, engine, which):
        self.engine, self.which = engine, which

    def split(self, deepcopy=False):
        return tuple(CellField(self.engine, self.which, k, f"c{k}") for k in range(self.engine.N - 1))

    def sub(self, k):
        return CellField(self.engine, self.which, k, f"c{k}")

    def vector(self):
        def get():
            return np.concatenate([self.engine.ctx.get_field(self.which, k) for k in range(self.engine.N - 1)])
        return _Vector(get, None)


class FacetField:
    """Function on Q (one value per membrane facet) living on the device."""

    def __init__(self, engine, which, idx=0):
        self.engine, self.which, self.idx = engine, which, idx

    def vector(self):
        return _Vector(lambda: self.engine.ctx.get_field(self.which, self.idx),
                       lambda v: self.engine.ctx.set_field(self.which, self.idx, v))

    def function_space(self):
        return FunctionSpace(self.engine, "DLT0", self.engine.mesh)


class HostFacetField:
    """Function on Q held on the host (user-made, e.g. a constant initial phi_M)."""

    def __init__(self, engine, values=None):
        self.engine = engine
        self.values = np.zeros(engine.nm) if values is None else np.asarray(values, dtype=float).copy()

    def vector(self):
        return _Vector(lambda: self.values.copy(), lambda v: self.values.__setitem__(slice(None), v))

    def function_space(self):
        return FunctionSpace(self.engine, "DLT0", self.engine.mesh)


class InterfaceNormal:
    """n_g of the reference (utils.py:61-85): only its orientation convention is needed
    (lower cell tag -> higher cell tag), which the library fixed when it built the
    membrane table."""

    def __init__(self, engine=None):
        self.engine = engine


class Trace:
    """plus(f, n_g) / minus(f, n_g): one-sided trace of a cell field (utils.py:87-98)."""

    def __init__(self, field, side):
        self.field, self.side = field, side          # side 0 = plus / ECS, 1 = minus / ICS


class FacetMean:
    """pcws_constant_project(trace, Q) (utils.py:100-124): facet mean of a trace; evaluated
    on the device (inside the ODE kernel when linked to a parameter)."""

    def __init__(self, trace):
        self.trace = trace

    def vector(self):
        f = self.trace.field
        return _Vector(lambda: f.engine.ctx.facet_trace(f.which, f.idx, self.trace.side), None)

    def function_space(self):
        return FunctionSpace(self.trace.field.engine, "DLT0")
