"""oracle/refexec/ufl_numeric.py: the symbolic derivatives and degree estimates that stand in for UFL's
apply_derivatives + estimate_total_polynomial_degree on the manufactured solutions of the reference's
tests/mms_space.py / mms_time.py (expressions of SpatialCoordinate, sin/cos, Constants)."""
import numpy as np

from common import kmesh
from oracle.refexec import ufl_numeric as U
from oracle import quadrature as quad


def _ctx(mesh, degree=4):
    G = U.geometry_of(mesh)
    bary, _ = quad.cell_rule(G.d, degree)
    return U.Ctx(G, "cell", np.arange(G.nc), bary)


def _vals(e, ctx):
    return np.broadcast_to(e.eval(ctx, None), (ctx.E, ctx.Q) + e.eval(ctx, None).shape[2:])[..., 0, 0]


def _mms_like(mesh):
    x = U.SpatialCoordinate(mesh)
    two_pi = 2 * np.pi
    k = 0.3 + 0.2 * U.MathFunction("sin", two_pi * x[0]) * U.MathFunction("sin", two_pi * x[1])
    phi = U.MathFunction("cos", two_pi * x[0]) * U.MathFunction("cos", two_pi * x[1])
    return x, k, phi


def test_derivatives_match_the_closed_forms():
    mesh, _, _ = kmesh.mms_mesh(2)
    ctx = _ctx(mesh)
    x, k, phi = _mms_like(mesh)
    X, Y = ctx.x[..., 0], ctx.x[..., 1]
    tp = 2 * np.pi
    assert np.allclose(_vals(U.deriv(k, 0), ctx)[..., 0], 0.2 * tp * np.cos(tp * X) * np.sin(tp * Y), rtol=0, atol=1e-14)
    assert np.allclose(_vals(U.deriv(k, 1), ctx)[..., 0], 0.2 * tp * np.sin(tp * X) * np.cos(tp * Y), rtol=0, atol=1e-14)
    # J = -D grad(k) - z D psi k grad(phi);  div J in closed form
    D, z = U.Constant(6.0), U.Constant(-1.0)
    J = -D * U.sym_grad(k) - z * D * k * U.sym_grad(phi)
    kx, ky = 0.2 * tp * np.cos(tp * X) * np.sin(tp * Y), 0.2 * tp * np.sin(tp * X) * np.cos(tp * Y)
    kk = 0.3 + 0.2 * np.sin(tp * X) * np.sin(tp * Y)
    px, py = -tp * np.sin(tp * X) * np.cos(tp * Y), -tp * np.cos(tp * X) * np.sin(tp * Y)
    lap_k = -2 * tp ** 2 * 0.2 * np.sin(tp * X) * np.sin(tp * Y)
    lap_p = -2 * tp ** 2 * np.cos(tp * X) * np.cos(tp * Y)
    expect = -6.0 * lap_k + 6.0 * (kx * px + ky * py + kk * lap_p)
    got = _vals(U.sym_div(J), ctx)[..., 0]
    assert np.abs(got - expect).max() < 1e-11 * np.abs(expect).max()
    # quotient and power rules
    q = (1.0 + x[0]) ** 2 / (2.0 + x[1])
    assert np.allclose(_vals(U.deriv(q, 0), ctx)[..., 0], 2 * (1 + X) / (2 + Y), rtol=1e-14)
    assert np.allclose(_vals(U.deriv(q, 1), ctx)[..., 0], -(1 + X) ** 2 / (2 + Y) ** 2, rtol=1e-14)


def test_degree_estimates_follow_ufl():
    mesh, _, _ = kmesh.mms_mesh(2)
    x, k, phi = _mms_like(mesh)
    t = U.Constant(0.25)
    assert x.degree() == 1 and (2 * np.pi * x[0]).degree() == 1
    assert U.MathFunction("sin", 2 * np.pi * x[0]).degree() == 3            # degree(argument) + 2
    assert U.MathFunction("cos", 2 * np.pi * t).degree() == 0               # sin(const) counts 0
    assert k.degree() == 6 and phi.degree() == 6
    J = -U.Constant(6.0) * U.sym_grad(k) - U.Constant(2.0) * k * U.sym_grad(phi)
    assert J.degree() == 12 and U.sym_div(J).degree() == 12                 # + 1 for the test function: the rule of degree 13
    lin = 1 + (x[0] + x[1]) + 0.2 * U.MathFunction("cos", 2 * np.pi * t)    # mms_time.py:28
    assert lin.degree() == 1
    g = U.sym_grad(lin)
    assert [c.const_value() for c in g.comps] == [1.0, 1.0] and U.sym_div(lin * g).degree() == 0


def test_zero_derivatives_are_dropped():
    """UFL simplifies 0 * f and f + 0 away before estimating; so does the stand-in (otherwise the degree of
    grad(k) * 0 would still count)"""
    mesh, _, _ = kmesh.mms_mesh(2)
    x, k, _ = _mms_like(mesh)
    t = U.Constant(0.5)
    e = (1 + t ** 2) * (1 + x[0] - x[1])                                    # mms_time.py:36
    d0, d1 = U.deriv(e, 0), U.deriv(e, 1)
    assert d0.degree() == 0 and abs(d0.const_value() - 1.25) < 1e-15 and abs(d1.const_value() + 1.25) < 1e-15
    assert U._is_zero(U.deriv(U.MathFunction("sin", 2 * np.pi * t), 0))


def test_vector_constant_and_iteration():
    mesh, _, _ = kmesh.mms_mesh(2)
    n = U.Constant((-1, 0))
    assert isinstance(n, U.ListVector) and n.ufl_shape == (2,)
    xs = list(U.SpatialCoordinate(mesh))
    assert len(xs) == 2 and all(c.ufl_shape == () for c in xs)
    ctx = _ctx(mesh)
    _, k, _ = _mms_like(mesh)
    dot = U.Inner(U.sym_grad(k), n)
    assert np.allclose(_vals(dot, ctx)[..., 0], -_vals(U.deriv(k, 0), ctx)[..., 0])
