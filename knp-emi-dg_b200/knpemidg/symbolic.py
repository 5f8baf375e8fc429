"""A small symbolic layer (sympy) for the UFL expressions that cross the reference's Solver API.

The reference's manufactured-solution tests (tests/mms_space.py, tests/mms_time.py,
tests/run_MMS_*.py) build their exact solutions, sources and interface data as UFL
expressions of `SpatialCoordinate(mesh)`, dolfin `Constant`s (among them the time `t`, which the
solver advances with `t.assign`) and `grad/div/dot/inner`, hand them to `Solver(mms=...)`
inside the ion dictionaries, and afterwards integrate error norms that mix those expressions
with the discrete solution (`inner(ca1 - uh_ca, ca1 - uh_ca)*dX(1, metadata=...)`).

Here the same objects are sympy expressions in the coordinates x0, x1, x2 and one symbol per
`Constant` (looked up at evaluation time, so `t.assign` is seen), `FieldExpr` trees where a
device field enters, and quadrature-backed `Measure` / `Form` / `assemble`.  Host-side, setup
and post-processing only; nothing here runs inside a time step except the evaluation of the
MMS load vectors (knpemidg/mms_loads.py), which the reference also re-assembles every step.
"""
from __future__ import annotations

import math
import re

import numpy as np
import sympy as sy

X = sy.symbols("x0 x1 x2", real=True)
_CONSTANTS = {}           # Symbol -> Constant (frontend.Constant registers itself)
_counter = [0]


def constant_symbol(const):
    s = sy.Symbol(f"_c{_counter[0]}", real=True)
    _counter[0] += 1
    _CONSTANTS[s] = const
    return s


def is_symbolic(v):
    return isinstance(v, (sy.Basic, Vec, FieldExpr)) or hasattr(v, "_sympy_")


def to_sym(v):
    """number / Constant / sympy -> sympy expression"""
    if isinstance(v, sy.Basic):
        return v
    if hasattr(v, "_sympy_"):
        return v._sympy_()
    return sy.sympify(v)


def numeric(v):
    """float of a number, a Constant or a symbolic expression of Constants (current values)"""
    if isinstance(v, sy.Basic):
        free = v.free_symbols
        if free:
            v = v.subs({s: float(_CONSTANTS[s]) for s in free if s in _CONSTANTS})
        return float(v)
    return float(v)


class Vec:
    """small vector of symbolic components (the value of grad(f), a flux, a constant normal)"""

    def __init__(self, comps):
        self.c = [to_sym(v) for v in comps]

    def __len__(self):
        return len(self.c)

    def __getitem__(self, i):
        return self.c[i]

    def __add__(self, o):
        return Vec([a + b for a, b in zip(self.c, as_vec(o).c)])

    __radd__ = __add__

    def __sub__(self, o):
        return Vec([a - b for a, b in zip(self.c, as_vec(o).c)])

    def __rsub__(self, o):
        return Vec([b - a for a, b in zip(self.c, as_vec(o).c)])

    def __neg__(self):
        return Vec([-a for a in self.c])

    def __mul__(self, o):
        s = to_sym(o)
        return Vec([a * s for a in self.c])

    __rmul__ = __mul__

    def __truediv__(self, o):
        s = to_sym(o)
        return Vec([a / s for a in self.c])

    # let sympy scalars hand the operation over to us (scalar * Vec)
    _op_priority = 20.0


def as_vec(v):
    if isinstance(v, Vec):
        return v
    if hasattr(v, "values") and not isinstance(v, sy.Basic):       # vector Constant
        return Vec([float(a) for a in v.values()])
    return Vec(list(v))


_gdim = [2]


def SpatialCoordinate(mesh):
    _gdim[0] = int(mesh.gdim)
    return tuple(X[: mesh.gdim])


def grad(f):
    return Vec([sy.diff(to_sym(f), X[i]) for i in range(_gdim[0])])


def div(v):
    v = as_vec(v)
    return sum(sy.diff(v[i], X[i]) for i in range(len(v)))


def dot(a, b):
    a_vec = isinstance(a, Vec) or (hasattr(a, "values") and np.size(a.values()) > 1)
    b_vec = isinstance(b, Vec) or (hasattr(b, "values") and np.size(b.values()) > 1)
    if a_vec or b_vec:
        a, b = as_vec(a), as_vec(b)
        return sum(x * y for x, y in zip(a.c, b.c))
    return inner(a, b)


def inner(a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        return dot(a, b)
    if isinstance(a, FieldExpr) or isinstance(b, FieldExpr) or _is_field(a) or _is_field(b):
        return FieldExpr("mul", a, b)
    return to_sym(a) * to_sym(b)


def _fn(sym_fn, num_fn):
    def f(v):
        if isinstance(v, (int, float, np.floating, np.integer)):
            return num_fn(v)
        return sym_fn(to_sym(v))
    return f


sin = _fn(sy.sin, math.sin)
cos = _fn(sy.cos, math.cos)
exp = _fn(sy.exp, math.exp)
ln = _fn(sy.log, math.log)
sqrt = _fn(sy.sqrt, math.sqrt)
pi = sy.pi


# ---- evaluation ---------------------------------------------------------------------------
def evaluate(expr, pts):
    """expr (number, Constant, sympy) at points pts[..., d] -> array pts.shape[:-1]"""
    pts = np.asarray(pts, dtype=float)
    if isinstance(expr, (int, float, np.floating)):
        return np.full(pts.shape[:-1], float(expr))
    e = to_sym(expr)
    free = e.free_symbols
    consts = sorted((s for s in free if s in _CONSTANTS), key=str)
    if consts:
        e = e.subs({s: float(_CONSTANTS[s]) for s in consts})
    coords = [pts[..., i] if i < pts.shape[-1] else np.zeros(pts.shape[:-1]) for i in range(3)]
    f = sy.lambdify(X, e, modules="numpy")
    return np.broadcast_to(np.asarray(f(*coords), dtype=float), pts.shape[:-1]).copy()


# ---- expressions that contain a device field --------------------------------------------
def _is_field(v):
    return hasattr(v, "nodal") and hasattr(v, "engine")


class FieldExpr:
    """expression tree mixing symbolic parts with P1 cell fields, evaluated cell by cell at
    quadrature points"""

    def __init__(self, op, a, b=None):
        self.op, self.a, self.b = op, a, b

    @staticmethod
    def _wrap(v):
        return v

    def __add__(self, o): return FieldExpr("add", self, o)
    def __radd__(self, o): return FieldExpr("add", o, self)
    def __sub__(self, o): return FieldExpr("sub", self, o)
    def __rsub__(self, o): return FieldExpr("sub", o, self)
    def __neg__(self): return FieldExpr("neg", self)

    def __mul__(self, o):
        if isinstance(o, MeasureTag):
            return NotImplemented
        return FieldExpr("mul", self, o)

    def __rmul__(self, o): return FieldExpr("mul", o, self)

    def eval(self, pts, bary, cells):
        a = _eval_any(self.a, pts, bary, cells)
        if self.op == "neg":
            return -a
        b = _eval_any(self.b, pts, bary, cells)
        return {"add": a + b, "sub": a - b, "mul": a * b}[self.op]


def _eval_any(v, pts, bary, cells):
    if isinstance(v, FieldExpr):
        return v.eval(pts, bary, cells)
    if _is_field(v):
        nodal = v.nodal()[cells]                       # [ncells, nd]
        return nodal @ bary.T                          # [ncells, nq]
    return evaluate(v, pts)


def field_ops(cls):
    """arithmetic of a device-field handle with symbolic expressions (class decorator)"""
    cls.__add__ = lambda s, o: FieldExpr("add", s, o)
    cls.__radd__ = lambda s, o: FieldExpr("add", o, s)
    cls.__sub__ = lambda s, o: FieldExpr("sub", s, o)
    cls.__rsub__ = lambda s, o: FieldExpr("sub", o, s)
    cls.__neg__ = lambda s: FieldExpr("neg", s)
    cls.__mul__ = lambda s, o: NotImplemented if isinstance(o, MeasureTag) else FieldExpr("mul", s, o)
    cls.__rmul__ = lambda s, o: FieldExpr("mul", o, s)
    return cls


# ---- integration --------------------------------------------------------------------------
def simplex_rule(dim, n):
    """collapsed tensor Gauss rule on the reference simplex: barycentric points [nq, dim+1],
    weights summing to 1; exact for polynomials of degree 2n - dim"""
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    if dim == 1:
        return np.column_stack([1 - x, x]), w
    if dim == 2:
        U, V = np.meshgrid(x, x, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        l1, l2 = U.ravel(), (V * (1 - U)).ravel()
        return np.column_stack([1 - l1 - l2, l1, l2]), (WU * WV * (1 - U)).ravel() * 2.0
    U, V, W = np.meshgrid(x, x, x, indexing="ij")
    WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
    l1, l2, l3 = U.ravel(), (V * (1 - U)).ravel(), (W * (1 - U) * (1 - V)).ravel()
    return (np.column_stack([1 - l1 - l2 - l3, l1, l2, l3]),
            (WU * WV * WW * (1 - U) ** 2 * (1 - V)).ravel() * 6.0)


class MeasureTag:
    def __init__(self, mesh, subdomains, tag, degree):
        self.mesh, self.subdomains, self.tag, self.degree = mesh, subdomains, tag, degree

    def __rmul__(self, integrand):
        return Form([(integrand, self)])


class Measure:
    """dolfin.Measure('dx', domain=mesh, subdomain_data=cell_function)"""

    def __init__(self, kind, domain=None, subdomain_data=None):
        if kind != "dx":
            raise NotImplementedError("only cell measures are integrated here")
        self.mesh, self.subdomains = domain, subdomain_data

    def __call__(self, tag=None, metadata=None):
        degree = (metadata or {}).get("quadrature_degree", 5)
        return MeasureTag(self.mesh, self.subdomains, tag, degree)

    def __rmul__(self, integrand):
        return Form([(integrand, self())])


class Form:
    def __init__(self, terms):
        self.terms = list(terms)

    def __add__(self, o):
        return Form(self.terms + o.terms)


def assemble(form):
    """value of a scalar Form"""
    total = 0.0
    for integrand, m in form.terms:
        mesh = m.mesh
        tags = np.asarray(m.subdomains.array()) if m.subdomains is not None else None
        cells = np.arange(mesh.num_cells()) if (m.tag is None or tags is None) else np.flatnonzero(tags == m.tag)
        if cells.size == 0:
            continue
        d = mesh.gdim
        bary, w = simplex_rule(d, m.degree // 2 + 2)
        pts = np.einsum("qa,cak->cqk", bary, mesh.coords[mesh.cells[cells]])
        vol = mesh.cell_volume()[cells]
        vals = _eval_any(integrand, pts, bary, cells)
        total += float((vals * w[None, :] * vol[:, None]).sum())
    return total


# ---- dolfin.Expression(C string) ----------------------------------------------------------
class Expression:
    """scalar dolfin Expression given as a C++ string in x[i], pi and user parameters; callable
    on a point (what the Solver's 'expression' initial conditions need)"""

    def __init__(self, code, degree=None, **params):
        self.params = params
        py = code.replace("\\\n", " ").replace("\n", " ")
        py = re.sub(r"\bpow\s*\(", "_pow(", py)
        self._code = compile(" ".join(py.split()), "<Expression>", "eval")

    def __call__(self, x):
        env = {"x": x, "pi": math.pi, "sin": math.sin, "cos": math.cos, "exp": math.exp, "sqrt": math.sqrt,
               "_pow": math.pow, "fabs": abs, "tanh": math.tanh, "log": math.log}
        for k, v in self.params.items():
            env[k] = numeric(v)
        return float(eval(self._code, {"__builtins__": {}}, env))
