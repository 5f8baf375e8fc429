"""ORACLE (test infrastructure only - never imported by the product path).

Quadrature rules on reference simplices, in barycentric coordinates.

The reference integrates every form with FFC-generated quadrature whose degree
is UFL's estimate of the integrand degree (call sites
src/knpemidg/solver.py:452-453, 477-479, 710, 730-731; utils.py:122).  All
matrix entries and the EMI right-hand side are polynomial, so any rule of
sufficient degree reproduces them to rounding; the only rule-sensitive
integrands are the rational membrane terms (solver.py:603-629, estimated
degree 5) and the Nernst logarithm (solver.py:299, 827, estimated degree 4),
for which the rules FFC/FIAT select by default are tabulated here:
interval: Gauss-Legendre ceil((deg+1)/2) points; triangle degree 4: 6-point
Dunavant rule; triangle degree 5: 7-point (Radon) rule.

FFC/FIAT are not vendored in /root/reference (unpinned dependency, environment.yml:5); the
rules are checked for monomial exactness in tests/test_oracle.py.  Which rule is used where
follows UFL's degree estimation as restated in oracle/refexec/ufl_numeric.py.
"""
import numpy as np
from math import factorial


def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def interval_rule(degree):
    n = max(1, (degree + 2) // 2)
    x, w = gauss_legendre_01(n)
    return np.column_stack([1.0 - x, x]), w


def duffy_rule(dim, n):
    """Collapsed tensor Gauss rule with n points per direction; weights sum
    to 1 (i.e. relative to the simplex measure).  Exact to degree 2n-1-(dim-1)."""
    x, w = gauss_legendre_01(n)
    if dim == 1:
        return np.column_stack([1.0 - x, x]), w
    if dim == 2:
        U, V = np.meshgrid(x, x, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        l1 = U.ravel()
        l2 = (V * (1 - U)).ravel()
        ww = (WU * WV * (1 - U)).ravel() * 2.0
        return np.column_stack([1 - l1 - l2, l1, l2]), ww
    if dim == 3:
        U, V, W = np.meshgrid(x, x, x, indexing="ij")
        WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
        l1 = U.ravel()
        l2 = (V * (1 - U)).ravel()
        l3 = (W * (1 - U) * (1 - V)).ravel()
        ww = (WU * WV * WW * (1 - U) ** 2 * (1 - V)).ravel() * 6.0
        return np.column_stack([1 - l1 - l2 - l3, l1, l2, l3]), ww
    raise ValueError(dim)


def triangle_deg4():
    """Dunavant 6-point, degree 4."""
    a, wa = 0.445948490915965, 0.223381589678011
    b, wb = 0.091576213509771, 0.109951743655322
    pts, w = [], []
    for s, ws in ((a, wa), (b, wb)):
        t = 1.0 - 2.0 * s
        pts += [(t, s, s), (s, t, s), (s, s, t)]
        w += [ws] * 3
    return np.array(pts), np.array(w)


def triangle_deg5():
    """Radon 7-point, degree 5 (closed forms)."""
    s15 = np.sqrt(15.0)
    a = (6.0 - s15) / 21.0
    b = (6.0 + s15) / 21.0
    wa = (155.0 - s15) / 1200.0
    wb = (155.0 + s15) / 1200.0
    pts = [(1.0 / 3, 1.0 / 3, 1.0 / 3)]
    w = [0.225]
    for s, ws in ((a, wa), (b, wb)):
        t = 1.0 - 2.0 * s
        pts += [(t, s, s), (s, t, s), (s, s, t)]
        w += [ws] * 3
    return np.array(pts), np.array(w)


def facet_rule(gdim, degree):
    """Default rule on a facet of a gdim-dimensional simplex."""
    if gdim == 2:
        return interval_rule(degree)
    if degree <= 4:
        return triangle_deg4()
    if degree == 5:
        return triangle_deg5()
    return duffy_rule(2, degree // 2 + 2)


def cell_rule(gdim, degree):
    return duffy_rule(gdim, degree // 2 + 2)


def monomial_integral(alpha):
    """int over the unit-measure simplex of prod lambda_i^alpha_i."""
    m = len(alpha) - 1
    num = factorial(m)
    for a in alpha:
        num *= factorial(a)
    return num / factorial(sum(alpha) + m)
