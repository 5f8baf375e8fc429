"""ORACLE (test infrastructure only - never imported by the product path).

The names the reference imports from `dolfin` (`from dolfin import *` in
src/knpemidg/solver.py:1, `import dolfin as df` in utils.py, membrane.py,
dlt_dof_extraction.py), backed by oracle/refexec/ufl_numeric.py.  Installed as
`sys.modules['dolfin']` by oracle.refexec.install() - in a process of its own, never in one
that runs the product.
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import ufl_numeric as U
from .ufl_numeric import (Constant, Function, FunctionSpace as _FS, Measure, assemble, Vector, Matrix,  # noqa: F401
                          Expr, as_expr)

parameters = {"ghost_mode": "none", "form_compiler": {}}


def info(msg):
    pass


# -- mesh side ---------------------------------------------------------------------------------
class _Comm:
    pass


class MPI:
    comm_world = _Comm()

    @staticmethod
    def min(comm, x):
        return x

    @staticmethod
    def max(comm, x):
        return x

    @staticmethod
    def rank(comm):
        return 0

    @staticmethod
    def size(comm):
        return 1


class _Dim:
    def __init__(self, d):
        self._d = d

    def dim(self):
        return self._d


class Mesh:
    """wraps a host simplex mesh (coords, cells + facet topology)"""

    def __init__(self, simplex):
        simplex.init_topology()
        self.__dict__["_m"] = simplex

    def __getattr__(self, name):
        return getattr(self.__dict__["_m"], name)

    def __setattr__(self, name, value):
        setattr(self.__dict__["_m"], name, value)

    def geometry(self):
        return _Dim(self._m.gdim)

    def topology(self):
        return _Dim(self._m.gdim)

    def coordinates(self):
        return self._m.coords

    def ufl_cell(self):
        return "triangle" if self._m.gdim == 2 else "tetrahedron"

    def mpi_comm(self):
        return MPI.comm_world

    def hmin(self):
        return float(self._m.cell_diameter().min())


class MeshFunction:
    def __init__(self, value_type, mesh, dim, value=0):
        self._mesh, self._dim = mesh, dim
        n = mesh.num_cells() if dim == mesh.gdim else mesh.facet_cells.shape[0]
        self._a = np.full(n, value, dtype=np.int64)

    @classmethod
    def from_array(cls, mesh, dim, array):
        f = cls("size_t", mesh, dim, 0)
        f._a[:] = array
        return f

    def mesh(self):
        return self._mesh

    def dim(self):
        return self._dim

    def array(self):
        return self._a

    def where_equal(self, v):
        return np.flatnonzero(self._a == v)

    def __getitem__(self, ent):
        return int(self._a[ent.index() if hasattr(ent, "index") else ent])


class _Cell:
    def __init__(self, i):
        self._i = i

    def index(self):
        return self._i


def cells(mesh):
    return (_Cell(i) for i in range(mesh.num_cells()))


# -- spaces -------------------------------------------------------------------------------------
class FiniteElement:
    def __init__(self, family, cell=None, degree=None):
        self.fam, self.deg = U._FAMILY[family], degree


class MixedElement:
    def __init__(self, elements):
        self.elements = list(elements)


def _kind(fam, deg):
    fam = U._FAMILY[fam] if fam in U._FAMILY else fam
    if fam == "Discontinuous Lagrange" and deg in (0, 1):
        return "DG%d" % deg
    if fam == "HDiv Trace" and deg == 0:
        return "DLT0"
    raise NotImplementedError(f"function space {fam} {deg}")


def FunctionSpace(mesh, element, degree=None):
    if isinstance(element, MixedElement):
        k = {_kind(e.fam, e.deg) for e in element.elements}
        assert len(k) == 1
        return _FS(mesh, k.pop(), ncomp=len(element.elements))
    if isinstance(element, FiniteElement):
        return _FS(mesh, _kind(element.fam, element.deg))
    return _FS(mesh, _kind(element, degree))


def VectorFunctionSpace(mesh, family, degree):
    return _FS(mesh, _kind(family, degree), ncomp=mesh.gdim, vector=True)


def TestFunction(V):
    return U.Argument(V, 0)


def TrialFunction(V):
    return U.Argument(V, 1)


def TestFunctions(V):
    return tuple(U.Argument(V, 0, k) for k in range(V.ncomp))


def TrialFunctions(V):
    return tuple(U.Argument(V, 1, k) for k in range(V.ncomp))


def split(f):
    return f.split()


def assign(target, source):
    target.assign(source)


class _Aux:
    class _PCWS:
        pass

    def PCWS(self):
        return _Aux._PCWS()


def compile_cpp_code(code):
    """utils.py:42 compiles a 35-line pybind11 Expression returning the cell tag (utils.py:5-39)"""
    return _Aux()


class CompiledExpression:
    def __init__(self, obj, degree=0, **kw):
        assert isinstance(obj, _Aux._PCWS)
        self.subdomains = kw["subdomains"]


class Expression(U.Expr):
    """A python callable f(x[npts, d]) -> values stands in for the C++ string expressions of the
    run scripts (e.g. the source window of run_tortuosity.py:180-200, which reads the time from a
    Constant it captured).  In a form it is evaluated at the quadrature points and counts with its
    declared `degree` in the degree estimate; dolfin would first interpolate it into P_degree on
    every cell, which is the same thing for data that is polynomial (here: constant) per cell."""

    def __init__(self, fn, degree=1, **kw):
        if isinstance(fn, str):
            fn = _compile_cpp_expression(fn, kw)
        self.fn, self._degree = fn, degree

    def degree(self):
        return self._degree

    def eval(self, ctx, side):
        x = ctx.x.reshape(-1, ctx.x.shape[-1])
        return np.asarray(self.fn(x), dtype=float).reshape(ctx.E, ctx.Q)[:, :, None, None, None]


class _Coords:
    def __init__(self, x):
        self._x = x

    def __getitem__(self, i):
        return self._x[:, i]


def _compile_cpp_expression(code, kw):
    """the C++ strings of the reference's MMS initial data (tests/mms_space.py:41-51, mms_time.py:46-52):
    arithmetic in x[0], x[1], pi, sin, cos, exp and the keyword parameters (numbers or Constants, read when
    the expression is evaluated)"""
    src = compile(" ".join(code.split()), "<Expression>", "eval")

    def fn(x):
        ns = {"x": _Coords(np.asarray(x)), "pi": np.pi, "sin": np.sin, "cos": np.cos, "exp": np.exp,
              "pow": np.power, "sqrt": np.sqrt, "__builtins__": {}}
        ns.update({k: float(v) for k, v in kw.items()})
        return eval(src, ns) + np.zeros(len(x))
    return fn


def interpolate(f, V):
    out = Function(V)
    G = V.G
    if isinstance(f, CompiledExpression):
        assert V.kind == "DG0"
        out.values[:] = f.subdomains.array()
    elif isinstance(f, U.Const):
        out.values[:] = f.v
    elif isinstance(f, Function):
        if f.V.kind == V.kind:
            out.values[:] = f.values
        elif f.V.kind == "DG0" and V.kind == "DG1":
            out.values[:] = np.repeat(f.values, G.nd)
        else:
            raise NotImplementedError("interpolate between these spaces")
    elif isinstance(f, Expression):
        assert V.kind == "DG1"
        out.values[:] = np.asarray(f.fn(G.X.reshape(-1, G.d)), dtype=float)
    else:
        raise NotImplementedError(f"interpolate({type(f)})")
    return out


def project(expr, V):
    """dolfin.project: L2 projection (mass matrix + right-hand side + direct solve), as at
    solver.py:837"""
    u, v = TrialFunction(V), TestFunction(V)
    dx_ = Measure("dx", domain=V.mesh())
    A = assemble(u * v * dx_)
    b = assemble(as_expr(expr) * v * dx_)
    out = Function(V)
    out.values[:] = spla.spsolve(A.A.tocsc(), b.array)
    return out


# -- operators ------------------------------------------------------------------------------------
def FacetNormal(mesh):
    return U.FacetNormal(mesh)


def CellDiameter(mesh):
    return U.CellQuantity(mesh, "diameter")


def CellVolume(mesh):
    return U.CellQuantity(mesh, "volume")


def FacetArea(mesh):
    return U.CellQuantity(mesh, "facet_area")


def grad(f):
    f = as_expr(f)
    if isinstance(f, (U.Function, U.Argument)):
        return U.Grad(f)
    return U.sym_grad(f)                 # an expression of SpatialCoordinate (the MMS data)


def div(v):
    return U.sym_div(as_expr(v))


def SpatialCoordinate(mesh):
    return U.SpatialCoordinate(mesh)


def sin(f):
    return U.MathFunction("sin", as_expr(f))


def cos(f):
    return U.MathFunction("cos", as_expr(f))


def exp(f):
    return U.MathFunction("exp", as_expr(f))


pi = np.pi


def inner(a, b):
    return U.Inner(as_expr(a), as_expr(b))


dot = inner


def jump(v, n=None):
    v = as_expr(v)
    if n is None:
        return v("+") - v("-")
    return v("+") * n("+") + v("-") * n("-")


def avg(v):
    v = as_expr(v)
    return 0.5 * (v("+") + v("-"))


def ln(f):
    return U.Ln(as_expr(f))


def sqrt(f):
    return U.Sqrt(as_expr(f))


def ge(a, b):
    return U.Condition("ge", a, b)


def gt(a, b):
    return U.Condition("gt", a, b)


def le(a, b):
    return U.Condition("le", a, b)


def lt(a, b):
    return U.Condition("lt", a, b)


def conditional(c, t, f):
    return U.Conditional(c, t, f)


def as_backend_type(x):
    return x


dx = Measure("dx")
ds = Measure("ds")
dS = Measure("dS")


class Timer:
    def __init__(self, name=None):
        self.start()

    def start(self):
        import time
        self._t0 = time.perf_counter()

    def stop(self):
        import time
        return time.perf_counter() - self._t0
