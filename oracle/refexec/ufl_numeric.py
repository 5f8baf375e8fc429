"""ORACLE (test infrastructure only - never imported by the product path).

A small NUMERIC stand-in for the part of UFL + dolfin's assembler that the reference's form
code uses.  It exists so that the reference's own, unmodified `setup_varform_emi`,
`setup_varform_knp` (src/knpemidg/solver.py:270-403, 534-663), `interface_normal`, `plus`,
`minus`, `pcws_constant_project` (src/knpemidg/utils.py:61-124) and the step updates
(solver.py:794-847) can be EXECUTED in this container, where FEniCS cannot be installed:
the expression objects the reference builds with `grad/inner/dot/jump/avg/ln/abs/conditional`
are kept as trees and evaluated at quadrature points when `assemble` is called.  The manufactured
solutions of tests/mms_space.py / mms_time.py (sin/cos expressions of SpatialCoordinate and the time
Constant under grad/div/dot) are differentiated symbolically first, as UFL's apply_derivatives does.

What is restated here (and therefore not pinned by the reference itself) is only what the
reference delegates to third-party code that is absent from /root/reference:
  * UFL's degree estimation (sum of factor degrees, +2 for ln, sin, cos and non-integer powers of
    non-constant arguments, max over sums and conditional branches, SpatialCoordinate counts 1,
    grad of a finite-element function lowers the degree by one on affine cells);
  * FFC/FIAT's default quadrature rule for that degree (oracle/quadrature.py);
  * dolfin's assembler: element tensors over cells / interior facets ('+' = first cell of
    the facet) / exterior facets added into a global tensor, integrals of one form that share
    a measure are integrated with ONE rule of the largest estimated degree (UFL groups them).
Spaces: DG0, DG1 (scalar and mixed), 'Discontinuous Lagrange Trace' 0 (scalar and vector), on
affine simplices.  dof numbering: DG1 dof = nd*cell + local vertex (+ component*n for mixed
spaces); DLT0 dof = facet index (vector: d*facet + component).

A value is carried as an array [E, Q, V, T, U]: entities x quadrature points x value
components (1 or d) x test slots x trial slots.
"""
from __future__ import annotations

import numbers

import numpy as np
import scipy.sparse as sp

from .. import quadrature as quad


# --------------------------------------------------------------------------------------
# geometry cache
# --------------------------------------------------------------------------------------
class Geometry:
    def __init__(self, mesh):
        mesh.init_topology()
        self.mesh = mesh
        self.d = d = mesh.gdim
        self.nd = nd = d + 1
        self.nc = mesh.num_cells()
        self.nf = mesh.facet_cells.shape[0]
        X = mesh.coords[mesh.cells]
        T = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)
        Tinv = np.linalg.inv(T)
        g = np.empty((self.nc, nd, d))
        g[:, 1:, :] = Tinv
        g[:, 0, :] = -Tinv.sum(axis=1)
        self.grad = g
        self.X = X
        self.vol = mesh.cell_volume()
        self.h = mesh.cell_diameter()
        FX = mesh.coords[mesh.facet_verts]
        if d == 2:
            t = FX[:, 1] - FX[:, 0]
            area = np.linalg.norm(t, axis=1)
            n = np.column_stack([t[:, 1], -t[:, 0]]) / area[:, None]
        else:
            cr = np.cross(FX[:, 1] - FX[:, 0], FX[:, 2] - FX[:, 0])
            nrm = np.linalg.norm(cr, axis=1)
            area = 0.5 * nrm
            n = cr / nrm[:, None]
        c0 = mesh.facet_cells[:, 0]
        l0 = mesh.facet_local[:, 0]
        opp = X[c0, l0]
        sgn = np.sign(np.einsum("fk,fk->f", FX.mean(axis=1) - opp, n))
        self.normal0 = n * sgn[:, None]          # outward from the facet's first cell
        self.farea = area
        self.FX = FX
        self.fmid = FX.mean(axis=1)


def geometry_of(mesh):
    g = getattr(mesh, "_refexec_geometry", None)
    if g is None:
        g = Geometry(mesh)
        mesh._refexec_geometry = g
    return g


class Ctx:
    """Where an integrand is being evaluated."""

    def __init__(self, G, kind, ents, bary):
        self.G, self.kind, self.ents = G, kind, np.asarray(ents)
        self.E, self.Q = len(self.ents), bary.shape[0]
        mesh = G.mesh
        if kind == "cell":
            self.cells = {None: self.ents}
            self.x = np.einsum("qa,eak->eqk", bary, G.X[self.ents])
            self.scale = G.vol[self.ents]
        else:
            self.x = np.einsum("qa,eak->eqk", bary, G.FX[self.ents])
            self.scale = G.farea[self.ents]
            c0 = mesh.facet_cells[self.ents, 0]
            if kind == "interior_facet":
                self.cells = {"+": c0, "-": mesh.facet_cells[self.ents, 1]}
            else:
                self.cells = {None: c0}
        self._lam = {}
        self.memo = {}

    def cell(self, side):
        if side not in self.cells:
            if self.kind == "interior_facet":
                raise ValueError("a cell-wise quantity must be restricted ('+' or '-') in a dS integral")
            return self.cells[None]           # a restriction inside dx / ds has no effect
        return self.cells[side]

    def lam(self, side):
        """P1 basis of the side's cell at the quadrature points [E, Q, nd]"""
        key = side if side in self.cells else None
        if key not in self._lam:
            c = self.cell(side)
            G = self.G
            self._lam[key] = 1.0 + np.einsum("eak,eqak->eqa", G.grad[c],
                                             self.x[:, :, None, :] - G.X[c][:, None, :, :])
        return self._lam[key]


# --------------------------------------------------------------------------------------
# expression tree
# --------------------------------------------------------------------------------------
def as_expr(v):
    if isinstance(v, Expr):
        return v
    if isinstance(v, numbers.Number) or isinstance(v, np.generic):
        return Const(float(v))
    raise TypeError(f"cannot use {type(v)} in a form expression")


class Expr:
    ufl_shape = ()

    def __init_subclass__(cls, **kw):
        """every node's eval is memoised per (context, side): the reference's integrands reuse the same
        sub-expressions (kappa, alpha_sum, the traces of c) dozens of times"""
        super().__init_subclass__(**kw)
        raw = cls.__dict__.get("eval")
        if raw is None:
            return

        def eval(self, ctx, side, _raw=raw):
            key = (id(self), side)
            hit = ctx.memo.get(key)
            if hit is None:             # (the entry keeps the node alive: its id cannot be recycled while the context lives)
                hit = ctx.memo[key] = (self, _raw(self, ctx, side))
            return hit[1]
        cls.eval = eval

    # -- algebra --
    def __add__(self, o):
        if isinstance(o, (Form, Measure)):
            return NotImplemented
        return Sum(self, as_expr(o))

    def __radd__(self, o):
        return Sum(as_expr(o), self)

    def __sub__(self, o):
        return Sum(self, Neg(as_expr(o)))

    def __rsub__(self, o):
        return Sum(as_expr(o), Neg(self))

    def __neg__(self):
        return Neg(self)

    def __pos__(self):
        return self

    def __mul__(self, o):
        if isinstance(o, Measure):
            return o.__rmul__(self)
        return Product(self, as_expr(o))

    def __rmul__(self, o):
        return Product(as_expr(o), self)

    def __truediv__(self, o):
        return Division(self, as_expr(o))

    def __rtruediv__(self, o):
        return Division(as_expr(o), self)

    def __pow__(self, p):
        return Power(self, as_expr(p))

    def __abs__(self):
        return Abs(self)

    def __call__(self, side):
        assert side in ("+", "-")
        return Restricted(self, side)

    def __float__(self):
        return float(self.const_value())

    def __getitem__(self, i):
        return Component(self, i)

    def __iter__(self):
        if len(self.ufl_shape) != 1:
            raise TypeError("iteration over a scalar expression")
        return iter([Component(self, i) for i in range(self.ufl_shape[0])])

    # -- protocol --
    def const_value(self):
        raise TypeError(f"{type(self).__name__} is not a constant expression")

    def children(self):
        return ()

    def degree(self):
        raise NotImplementedError

    def eval(self, ctx, side):
        raise NotImplementedError

    def arguments(self):
        out = {}
        for c in self.children():
            out.update(c.arguments())
        return out


class Const(Expr):
    def __init__(self, v):
        self.v = float(v)

    def const_value(self):
        return self.v

    def degree(self):
        return 0

    def eval(self, ctx, side):
        return np.full((1, 1, 1, 1, 1), self.v)


class Constant(Const):
    """dolfin.Constant (scalar); assign() mutates it, as `t.assign(...)` in solver.py:845"""

    def __new__(cls, v, name=None):
        if isinstance(v, (tuple, list, np.ndarray)):          # Constant((-1, 0)): the MMS normals (mms_space.py:77)
            return ListVector([Const(float(x)) for x in v])
        return super().__new__(cls)

    def __init__(self, v, name=None):
        super().__init__(float(v))

    def assign(self, v):
        self.v = float(v)

    def values(self):
        return np.array([self.v])


class Sum(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        assert a.ufl_shape == b.ufl_shape, "sum of a scalar and a vector"
        self.ufl_shape = a.ufl_shape

    def children(self):
        return (self.a, self.b)

    def const_value(self):
        return self.a.const_value() + self.b.const_value()

    def degree(self):
        return max(self.a.degree(), self.b.degree())

    def eval(self, ctx, side):
        return self.a.eval(ctx, side) + self.b.eval(ctx, side)


class Neg(Expr):
    def __init__(self, a):
        self.a = a
        self.ufl_shape = a.ufl_shape

    def children(self):
        return (self.a,)

    def const_value(self):
        return -self.a.const_value()

    def degree(self):
        return self.a.degree()

    def eval(self, ctx, side):
        return -self.a.eval(ctx, side)


class Product(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        assert a.ufl_shape == () or b.ufl_shape == (), "product of two vectors: use inner/dot"
        self.ufl_shape = a.ufl_shape or b.ufl_shape

    def children(self):
        return (self.a, self.b)

    def const_value(self):
        return self.a.const_value() * self.b.const_value()

    def degree(self):
        return self.a.degree() + self.b.degree()

    def eval(self, ctx, side):
        return self.a.eval(ctx, side) * self.b.eval(ctx, side)


class Division(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b
        assert b.ufl_shape == ()
        self.ufl_shape = a.ufl_shape

    def children(self):
        return (self.a, self.b)

    def const_value(self):
        return self.a.const_value() / self.b.const_value()

    def degree(self):
        return self.a.degree() + self.b.degree()      # UFL: f/g is estimated like f*g

    def eval(self, ctx, side):
        return self.a.eval(ctx, side) / self.b.eval(ctx, side)


class Power(Expr):
    def __init__(self, a, p):
        self.a, self.p = a, p

    def children(self):
        return (self.a, self.p)

    def const_value(self):
        return self.a.const_value() ** self.p.const_value()

    def degree(self):
        try:
            p = self.p.const_value()
            if p >= 0 and float(p).is_integer():
                return self.a.degree() * int(p)
        except TypeError:
            pass
        return self.a.degree() + 2

    def eval(self, ctx, side):
        return self.a.eval(ctx, side) ** self.p.eval(ctx, side)


class Abs(Expr):
    def __init__(self, a):
        self.a = a

    def children(self):
        return (self.a,)

    def const_value(self):
        return abs(self.a.const_value())

    def degree(self):
        return self.a.degree()

    def eval(self, ctx, side):
        return np.abs(self.a.eval(ctx, side))


class Ln(Expr):
    def __init__(self, a):
        self.a = a

    def children(self):
        return (self.a,)

    def degree(self):
        return self.a.degree() + 2                    # UFL: math functions add 2

    def eval(self, ctx, side):
        return np.log(self.a.eval(ctx, side))


class Sqrt(Ln):
    def eval(self, ctx, side):
        return np.sqrt(self.a.eval(ctx, side))


class Inner(Expr):
    """inner(a, b) / dot(a, b) for scalars and vectors"""

    def __init__(self, a, b):
        self.a, self.b = a, b
        assert a.ufl_shape == b.ufl_shape, (a.ufl_shape, b.ufl_shape)

    def children(self):
        return (self.a, self.b)

    def degree(self):
        return self.a.degree() + self.b.degree()

    def eval(self, ctx, side):
        return (self.a.eval(ctx, side) * self.b.eval(ctx, side)).sum(axis=2, keepdims=True)


class Component(Expr):
    def __init__(self, a, i):
        assert len(a.ufl_shape) == 1
        self.a, self.i = a, int(i)

    def children(self):
        return (self.a,)

    def degree(self):
        return self.a.degree()

    def eval(self, ctx, side):
        return self.a.eval(ctx, side)[:, :, self.i:self.i + 1]


class Restricted(Expr):
    def __init__(self, a, side):
        self.a, self.side = a, side
        self.ufl_shape = a.ufl_shape

    def children(self):
        return (self.a,)

    def degree(self):
        return self.a.degree()

    def eval(self, ctx, side):
        return self.a.eval(ctx, self.side)


class Condition:
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, as_expr(a), as_expr(b)

    def eval(self, ctx, side):
        a, b = self.a.eval(ctx, side), self.b.eval(ctx, side)
        return {"ge": a >= b, "gt": a > b, "le": a <= b, "lt": a < b}[self.op]


class Conditional(Expr):
    def __init__(self, cond, t, f):
        self.cond, self.t, self.f = cond, as_expr(t), as_expr(f)
        assert self.t.ufl_shape == self.f.ufl_shape
        self.ufl_shape = self.t.ufl_shape

    def children(self):
        return (self.cond.a, self.cond.b, self.t, self.f)

    def arguments(self):
        out = {}
        for c in (self.t, self.f):
            out.update(c.arguments())
        return out

    def degree(self):
        return max(self.t.degree(), self.f.degree())  # UFL ignores the condition

    def eval(self, ctx, side):
        c = self.cond.eval(ctx, side)
        t, f = self.t.eval(ctx, side), self.f.eval(ctx, side)
        return np.where(c, t, f)


class ListVector(Expr):
    """as_vector([...]) / a vector Constant / the value of grad(scalar expression)"""

    def __init__(self, comps):
        self.comps = [as_expr(c) for c in comps]
        assert all(c.ufl_shape == () for c in self.comps)
        self.ufl_shape = (len(self.comps),)

    def children(self):
        return tuple(self.comps)

    def degree(self):
        return max(c.degree() for c in self.comps)

    def eval(self, ctx, side):
        vals = np.broadcast_arrays(*[c.eval(ctx, side) for c in self.comps])
        return np.concatenate(vals, axis=2)


class MathFunction(Expr):
    """sin / cos / exp of a scalar; UFL estimates degree(argument) + 2"""
    FN = {"sin": np.sin, "cos": np.cos, "exp": np.exp}

    def __init__(self, name, a):
        assert a.ufl_shape == ()
        self.name, self.a = name, a

    def children(self):
        return (self.a,)

    def const_value(self):
        return float(self.FN[self.name](self.a.const_value()))

    def degree(self):
        d = self.a.degree()
        return d + 2 if d else d                         # UFL: degree(sin(const)) == 0

    def eval(self, ctx, side):
        return self.FN[self.name](self.a.eval(ctx, side))


class SpatialCoordinate(Expr):
    def __init__(self, mesh):
        self.mesh = mesh
        self.ufl_shape = (mesh.gdim,)

    def degree(self):
        return 1                                        # affine cells

    def eval(self, ctx, side):
        return ctx.x[:, :, :, None, None]


# -- derivatives of expressions of the spatial coordinate ------------------------------------
# The reference's manufactured solutions (tests/mms_space.py:31-74, mms_time.py:28-74) are UFL
# expressions of SpatialCoordinate; grad/div of them are expanded by UFL's apply_derivatives BEFORE
# the quadrature degree is estimated, so the estimate sees products of sin/cos factors.  The same is
# done here: differentiate the tree, drop exact zeros, then estimate on the result.
def _is_zero(e):
    return type(e) is Const and e.v == 0.0


def _add(a, b):
    if _is_zero(a):
        return b
    if _is_zero(b):
        return a
    return Sum(a, b)


def _mul(a, b):
    if _is_zero(a) or _is_zero(b):
        return Const(0.0)
    if type(a) is Const and a.v == 1.0:
        return b
    if type(b) is Const and b.v == 1.0:
        return a
    return Product(a, b)


def _gdim(e):
    if isinstance(e, (SpatialCoordinate, FacetNormal)):
        return e.ufl_shape[0]
    for c in e.children():
        d = _gdim(c)
        if d:
            return d
    return 0


def component(v, k):
    """scalar expression of component k of a vector-valued expression"""
    if v.ufl_shape == ():
        raise ValueError("component of a scalar")
    if isinstance(v, ListVector):
        return v.comps[k]
    if isinstance(v, Sum):
        return _add(component(v.a, k), component(v.b, k))
    if isinstance(v, Neg):
        return Neg(component(v.a, k))
    if isinstance(v, Product):
        return _mul(v.a, component(v.b, k)) if v.a.ufl_shape == () else _mul(component(v.a, k), v.b)
    if isinstance(v, Division):
        return Division(component(v.a, k), v.b)
    return Component(v, k)


def deriv(e, k):
    """d e / d x_k of a scalar expression of SpatialCoordinate, Constants and numbers"""
    assert e.ufl_shape == ()
    if isinstance(e, Const):
        return Const(0.0)
    if isinstance(e, Component) and isinstance(e.a, SpatialCoordinate):
        return Const(1.0 if e.i == k else 0.0)
    if isinstance(e, Sum):
        return _add(deriv(e.a, k), deriv(e.b, k))
    if isinstance(e, Neg):
        d = deriv(e.a, k)
        return d if _is_zero(d) else Neg(d)
    if isinstance(e, Product):
        return _add(_mul(deriv(e.a, k), e.b), _mul(e.a, deriv(e.b, k)))
    if isinstance(e, Division):
        da, db = deriv(e.a, k), deriv(e.b, k)
        out = Const(0.0) if _is_zero(da) else Division(da, e.b)
        if not _is_zero(db):
            out = _add(out, Neg(Division(_mul(e.a, db), Product(e.b, e.b))))
        return out
    if isinstance(e, Power):
        da = deriv(e.a, k)
        if _is_zero(da):
            return da
        p = e.p.const_value()
        return _mul(_mul(Const(p), Power(e.a, Const(p - 1.0))), da)
    if isinstance(e, MathFunction):
        da = deriv(e.a, k)
        if _is_zero(da):
            return da
        if e.name == "sin":
            return _mul(MathFunction("cos", e.a), da)
        if e.name == "cos":
            return _mul(Neg(MathFunction("sin", e.a)), da)
        return _mul(e, da)
    raise NotImplementedError(f"derivative of {type(e).__name__}")


def sym_grad(e):
    d = _gdim(e)
    if not d:
        raise ValueError("grad of an expression without a spatial coordinate")
    return ListVector([deriv(e, k) for k in range(d)])


def sym_div(v):
    out = Const(0.0)
    for k in range(v.ufl_shape[0]):
        out = _add(out, deriv(component(v, k), k))
    return out


# -- geometric quantities ----------------------------------------------------------------
class FacetNormal(Expr):
    def __init__(self, mesh):
        self.mesh = mesh
        self.ufl_shape = (mesh.gdim,)

    def degree(self):
        return 0

    def eval(self, ctx, side):
        assert ctx.kind != "cell", "FacetNormal in a cell integral"
        n = ctx.G.normal0[ctx.ents]
        if ctx.kind == "interior_facet":
            if side not in ("+", "-"):
                raise ValueError("FacetNormal must be restricted in a dS integral")
            if side == "-":
                n = -n
        return n[:, None, :, None, None]


class CellQuantity(Expr):
    def __init__(self, mesh, what):
        self.mesh, self.what = mesh, what

    def degree(self):
        return 0

    def eval(self, ctx, side):
        G = ctx.G
        if self.what == "facet_area":
            assert ctx.kind != "cell"
            v = G.farea[ctx.ents]
        else:
            c = ctx.cell(side)
            v = {"diameter": G.h, "volume": G.vol}[self.what][c]
        return v[:, None, None, None, None]


# -- function spaces -----------------------------------------------------------------------
class Element:
    def __init__(self, family, degree, shape=(), subs=None):
        self._family, self._degree, self._shape, self.subs = family, degree, tuple(shape), subs

    def family(self):
        return self._family

    def degree(self):
        return self._degree

    def value_shape(self):
        return self._shape

    def value_size(self):
        return int(np.prod(self._shape)) if self._shape else 1

    def __eq__(self, o):
        return isinstance(o, Element) and (self._family, self._degree, self._shape) == (o._family, o._degree, o._shape) \
            and (self.subs is None) == (o.subs is None) and (self.subs is None or len(self.subs) == len(o.subs))

    def __hash__(self):
        return hash((self._family, self._degree, self._shape))


_FAMILY = {"DG": "Discontinuous Lagrange", "Discontinuous Lagrange": "Discontinuous Lagrange",
           "Discontinuous Lagrange Trace": "HDiv Trace", "HDiv Trace": "HDiv Trace", "DGT": "HDiv Trace"}


class DofMap:
    """what src/knpemidg/dlt_dof_extraction.py:18-48 and solver.py:1260-1298 ask of a dofmap"""

    def __init__(self, V):
        self.V = V

    def ownership_range(self):
        return (0, self.V.dim())

    def entity_dofs(self, mesh, dim):
        assert self.V.kind == "DLT0" and dim == mesh.gdim - 1
        return list(range(self.V.dim()))

    def tabulate_entity_dofs(self, dim, i):
        return [0] if self.V.ncomp == 1 else list(range(self.V.ncomp))

    def local_to_global_index(self, d):
        return int(d)

    def cell_dofs(self, cell):
        assert self.V.kind == "DG1"
        nd = self.V.G.nd
        return np.arange(nd * cell, nd * cell + nd)


class FunctionSpace:
    """kind: 'DG0' | 'DG1' | 'DLT0'; ncomp > 1: mixed DG1 (components of one mixed function) or
    vector-valued DLT0"""

    def __init__(self, mesh, kind, ncomp=1, vector=False, parent=None, index=None):
        self._mesh, self.kind, self.ncomp, self.vector = mesh, kind, ncomp, vector
        self.G = geometry_of(mesh)
        self.parent, self.index = parent, index
        G = self.G
        self.n1 = {"DG0": G.nc, "DG1": G.nd * G.nc, "DLT0": G.nf}[kind]

    def mesh(self):
        return self._mesh

    def dim(self):
        return self.n1 * self.ncomp

    def ufl_element(self):
        fam = {"DG0": "Discontinuous Lagrange", "DG1": "Discontinuous Lagrange", "DLT0": "HDiv Trace"}[self.kind]
        deg = 1 if self.kind == "DG1" else 0
        if self.ncomp > 1 and not self.vector:
            return Element("Mixed", deg, (self.ncomp,), subs=[Element(fam, deg)] * self.ncomp)
        return Element(fam, deg, (self.ncomp,) if self.vector else ())

    def dofmap(self):
        return DofMap(self)

    def sub(self, i):
        assert self.ncomp > 1 and not self.vector
        return FunctionSpace(self._mesh, self.kind, 1, parent=self, index=i)

    def collapse(self):
        return FunctionSpace(self._mesh, self.kind, 1)

    def tabulate_dof_coordinates(self):
        G = self.G
        if self.kind == "DLT0":
            return np.repeat(G.fmid, self.ncomp, axis=0) if self.vector else G.fmid.copy()
        if self.kind == "DG0":
            return G.X.mean(axis=1)
        return G.X.reshape(-1, G.d)

    @property
    def elem_degree(self):
        return 1 if self.kind == "DG1" else 0


class Vector:
    """dolfin GenericVector and, at once, the PETSc Vec behind it"""

    def __init__(self, n_or_array):
        self.array = np.zeros(n_or_array) if isinstance(n_or_array, (int, np.integer)) else n_or_array

    # dolfin side
    def get_local(self):
        return self.array.copy()

    def set_local(self, a):
        self.array[:] = a

    def apply(self, mode):
        pass

    def update_ghost_values(self):
        pass

    def vec(self):
        return self

    def size(self):
        return self.array.size

    def __len__(self):
        return self.array.size

    def __getitem__(self, i):
        return self.array[i]

    def __setitem__(self, i, v):
        self.array[i] = v.array if isinstance(v, Vector) else v

    def __sub__(self, o):
        return self.array - (o.array if isinstance(o, Vector) else o)

    def zero(self):
        self.array[:] = 0.0

    def norm(self, kind="l2"):
        return float(np.linalg.norm(self.array))

    # PETSc side
    @property
    def array_w(self):
        return self.array

    @property
    def array_r(self):
        return self.array

    def axpy(self, a, x):
        self.array += a * x.array

    def ghostUpdate(self, addv=None, mode=None):
        pass

    def duplicate(self):
        return Vector(np.zeros_like(self.array))

    def copy(self):
        return Vector(self.array.copy())


class Function(Expr):
    def __init__(self, V, name=None, _array=None, _comp=None):
        self.V = V
        self._comp = _comp                 # component of a mixed function (shares the parent's array)
        if _array is None:
            _array = np.zeros(V.dim())
        self._vector = Vector(_array)
        self.ufl_shape = (V.ncomp,) if V.vector else ()
        self._name = name

    def function_space(self):
        return self.V

    def vector(self):
        if self._comp is not None:
            raise NotImplementedError("vector() of a sub-function")
        return self._vector

    @property
    def values(self):
        if self._comp is None:
            return self._vector.array
        return self._vector.array[self._comp * self.V.n1:(self._comp + 1) * self.V.n1]

    def split(self, deepcopy=False):
        assert self.V.ncomp > 1 and not self.V.vector
        out = []
        for k in range(self.V.ncomp):
            sub = FunctionSpace(self.V.mesh(), self.V.kind, 1, parent=self.V, index=k)
            f = Function(sub, _array=self._vector.array, _comp=k)
            f.V_n1 = self.V.n1
            if deepcopy:
                f = Function(sub.collapse(), _array=f.values.copy())
            out.append(f)
        return tuple(out)

    def sub(self, k):
        return self.split()[k]

    def assign(self, other):
        if isinstance(other, Function):
            assert other.values.shape == self.values.shape
            self.values[:] = other.values
        else:
            raise NotImplementedError("assign of an expression")

    def copy(self, deepcopy=True):
        return Function(FunctionSpace(self.V.mesh(), self.V.kind, self.V.ncomp, self.V.vector), _array=self.values.copy())

    def rename(self, *a):
        pass

    # -- evaluation --
    def degree(self):
        return self.V.elem_degree

    def _cellwise(self, comp):
        V = self.V
        a = self.values if comp is None else self._vector.array[comp * V.n1:(comp + 1) * V.n1]
        return a

    def eval(self, ctx, side):
        V, G = self.V, ctx.G
        if V.ncomp > 1 and not V.vector:
            raise ValueError("a mixed function must be split before it is used in a form")
        vals = self.values
        if V.kind == "DG1":
            c = ctx.cell(side)
            dofs = vals.reshape(G.nc, G.nd)[c]
            return np.einsum("eqa,ea->eq", ctx.lam(side), dofs)[:, :, None, None, None]
        if V.kind == "DG0":
            return vals[ctx.cell(side)][:, None, None, None, None]
        assert ctx.kind != "cell", "a facet function in a cell integral"
        if V.vector:
            return vals.reshape(G.nf, V.ncomp)[ctx.ents][:, None, :, None, None]
        return vals[ctx.ents][:, None, None, None, None]

    def eval_grad(self, ctx, side):
        V, G = self.V, ctx.G
        if V.kind == "DG0":
            return np.zeros((1, 1, G.d, 1, 1))
        assert V.kind == "DG1"
        c = ctx.cell(side)
        dofs = self.values.reshape(G.nc, G.nd)[c]
        return np.einsum("eak,ea->ek", G.grad[c], dofs)[:, None, :, None, None]


class Argument(Expr):
    def __init__(self, V, number, comp=None):
        self.V, self.number, self.comp = V, number, comp
        self.ufl_shape = (V.ncomp,) if V.vector else ()
        if V.ncomp > 1 and not V.vector and comp is None:
            self.ufl_shape = (V.ncomp,)

    def function_space(self):
        return self.V

    def arguments(self):
        return {self.number: self.V}

    def degree(self):
        return self.V.elem_degree

    def __getitem__(self, k):
        if self.V.ncomp > 1 and not self.V.vector:
            return Argument(self.V, self.number, k)
        return Component(self, k)

    def _place(self, ctx, side, vals):
        """vals [E, Q, V, nloc] -> [E, Q, V, T, U] with the slots of this argument filled"""
        V, G = self.V, ctx.G
        nslot = slots_per_entity(V, ctx)
        out = np.zeros(vals.shape[:3] + (nslot,))
        if V.kind == "DLT0":
            out[...] = vals
        else:
            nloc = G.nd if V.kind == "DG1" else 1
            per_cell = nloc * V.ncomp
            base = (per_cell if (ctx.kind == "interior_facet" and side == "-") else 0) + (self.comp or 0) * nloc
            out[..., base:base + nloc] = vals
        return out[:, :, :, :, None] if self.number == 0 else out[:, :, :, None, :]

    def eval(self, ctx, side):
        V, G = self.V, ctx.G
        if V.kind == "DLT0":
            assert ctx.kind != "cell"
            if V.vector:     # slot j = component j
                vals = np.broadcast_to(np.eye(V.ncomp)[None, None], (ctx.E, 1, V.ncomp, V.ncomp))
            else:
                vals = np.ones((ctx.E, 1, 1, 1))
            return self._place(ctx, side, vals)
        if V.ncomp > 1 and self.comp is None:
            raise ValueError("index the argument of a mixed space (TestFunctions / TrialFunctions)")
        if ctx.kind == "interior_facet" and side not in ("+", "-"):
            raise ValueError("a DG argument must be restricted in a dS integral")
        if V.kind == "DG1":
            vals = ctx.lam(side)[:, :, None, :]
        else:
            vals = np.ones((ctx.E, 1, 1, 1))
        return self._place(ctx, side, vals)

    def eval_grad(self, ctx, side):
        V, G = self.V, ctx.G
        assert V.kind == "DG1" and not (V.ncomp > 1 and self.comp is None)
        if ctx.kind == "interior_facet" and side not in ("+", "-"):
            raise ValueError("a DG argument must be restricted in a dS integral")
        g = G.grad[ctx.cell(side)]                               # [E, nd, d]
        vals = np.swapaxes(g, 1, 2)[:, None, :, :]               # [E, 1, d, nd]
        return self._place(ctx, side, vals)


def slots_per_entity(V, ctx):
    if V.kind == "DLT0":
        return V.ncomp
    nloc = (ctx.G.nd if V.kind == "DG1" else 1) * V.ncomp
    return 2 * nloc if ctx.kind == "interior_facet" else nloc


def global_dofs(V, ctx):
    """[E, nslot] global dof of every slot"""
    G = ctx.G
    if V.kind == "DLT0":
        f = ctx.ents[:, None]
        return f * V.ncomp + np.arange(V.ncomp)[None, :]

    def of_cells(c):
        nloc = G.nd if V.kind == "DG1" else 1
        loc = nloc * c[:, None] + np.arange(nloc)[None, :]
        return np.concatenate([k * V.n1 + loc for k in range(V.ncomp)], axis=1)
    if ctx.kind == "interior_facet":
        return np.concatenate([of_cells(ctx.cells["+"]), of_cells(ctx.cells["-"])], axis=1)
    return of_cells(ctx.cells[None])


class Grad(Expr):
    def __init__(self, a):
        if not isinstance(a, (Function, Argument)) or a.ufl_shape != ():
            raise NotImplementedError("grad of a composite expression")
        self.a = a
        self.ufl_shape = (a.V.mesh().gdim,)

    def children(self):
        return (self.a,)

    def degree(self):
        return max(self.a.degree() - 1, 0)            # affine simplex

    def eval(self, ctx, side):
        return self.a.eval_grad(ctx, side)


# -- measures and forms ----------------------------------------------------------------------
class Measure:
    def __init__(self, kind, domain=None, subdomain_data=None, subdomain_id=None, metadata=None):
        self.kind = {"dx": "cell", "dS": "interior_facet", "ds": "exterior_facet"}.get(kind, kind)
        self.domain, self.data, self.id, self.metadata = domain, subdomain_data, subdomain_id, metadata

    def __call__(self, subdomain_id=None, metadata=None, domain=None, subdomain_data=None):
        return Measure(self.kind, domain or self.domain, subdomain_data if subdomain_data is not None else self.data,
                       self.id if subdomain_id is None else subdomain_id, metadata or self.metadata)

    def __rmul__(self, integrand):
        return Form([(as_expr(integrand), self)])


class Form:
    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, o):
        if isinstance(o, Form):
            return Form(self.integrals + o.integrals)
        if isinstance(o, numbers.Number) and o == 0:
            return self
        return NotImplemented

    __radd__ = __add__

    def __neg__(self):
        return Form([(Neg(e), m) for e, m in self.integrals])

    def __sub__(self, o):
        return self + (-o)

    def __rmul__(self, s):
        return Form([(as_expr(s) * e, m) for e, m in self.integrals])

    def arguments(self):
        out = {}
        for e, _ in self.integrals:
            out.update(e.arguments())
        return out

    def mesh(self):
        for e, m in self.integrals:
            if m.domain is not None:
                return m.domain
            for V in e.arguments().values():
                return V.mesh()
        raise ValueError("form without a mesh")

    def groups(self):
        """integrals sharing (measure type, subdomain id, subdomain data) are summed and get ONE
        quadrature degree - the largest of their estimates (UFL: group_form_integrals followed by
        attach_estimated_degrees)"""
        out = {}
        for e, m in self.integrals:
            key = (m.kind, m.id, id(m.data) if m.id is not None else None,
                   None if not m.metadata else m.metadata.get("quadrature_degree"))
            out.setdefault(key, [m, []])[1].append(e)
        return list(out.values())


def entities_of(G, m):
    mesh = G.mesh
    if m.kind == "cell":
        ents = np.arange(G.nc)
    else:
        interior = mesh.facet_cells[:, 1] >= 0
        ents = np.flatnonzero(interior if m.kind == "interior_facet" else ~interior)
    if m.id is not None:
        if m.data is None:
            raise ValueError("a measure with a subdomain id needs subdomain_data")
        ents = ents[np.asarray(m.data.array())[ents] == m.id]
    return ents


class Matrix:
    """dolfin GenericMatrix and the PETSc Mat behind it"""

    def __init__(self, A):
        self.A = A.tocsr()
        self.nullspace = None

    def mat(self):
        return self

    def array(self):
        return self.A.toarray()

    # PETSc side
    def createVecs(self):
        return Vector(self.A.shape[1]), Vector(self.A.shape[0])

    def setNearNullSpace(self, ns):
        self.nullspace = ns

    def convert(self, *a):
        return self

    def getSize(self):
        return self.A.shape


def assemble(form, tensor=None):
    """dolfin.assemble for forms of arity 0 (not used), 1 and 2"""
    if not isinstance(form, Form):
        raise TypeError("assemble: not a form")
    G = geometry_of(form.mesh())
    args = form.arguments()
    arity = len(args)
    assert sorted(args) == list(range(arity)), "a trial function without a test function"
    Vt = args.get(0)
    Vu = args.get(1)
    rows, cols, vals = [], [], []
    vec = np.zeros(Vt.dim()) if arity == 1 else None
    scalar = 0.0
    for m, exprs in form.groups():
        ents = entities_of(G, m)
        if len(ents) == 0:
            continue
        deg = max(e.degree() for e in exprs)
        if m.metadata and m.metadata.get("quadrature_degree") is not None:
            deg = m.metadata["quadrature_degree"]
        bary, w = (quad.cell_rule(G.d, deg) if m.kind == "cell" else quad.facet_rule(G.d, deg))
        ctx = Ctx(G, m.kind, ents, bary)
        total = None
        for e in exprs:
            v = e.eval(ctx, None)
            total = v if total is None else total + v
        assert total.shape[2] == 1, "the integrand is not a scalar"
        nt = slots_per_entity(Vt, ctx) if arity >= 1 else 1
        nu = slots_per_entity(Vu, ctx) if arity == 2 else 1
        total = np.broadcast_to(total, (ctx.E, ctx.Q, 1, nt, nu))
        elem = np.einsum("q,e,eqtu->etu", w, ctx.scale, total[:, :, 0])
        if arity == 0:
            scalar += elem.sum()
        elif arity == 1:
            np.add.at(vec, global_dofs(Vt, ctx), elem[:, :, 0])
        else:
            gt, gu = global_dofs(Vt, ctx), global_dofs(Vu, ctx)
            rows.append(np.repeat(gt[:, :, None], nu, 2).ravel())
            cols.append(np.repeat(gu[:, None, :], nt, 1).ravel())
            vals.append(elem.ravel())
    if arity == 0:
        return scalar
    if arity == 1:
        if tensor is not None:
            tensor.array[:] = vec
            return tensor
        return Vector(vec)
    n, mcols = Vt.dim(), Vu.dim()
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, mcols)).tocsr()
    if tensor is not None:
        tensor.A = A
        return tensor
    return Matrix(A)
