"""Pins of the CPU oracle (oracle/): the reference holds no golden matrices,
so the restatement is pinned by what the reference's own material implies:
  * its MMS convergence study (tests/run_MMS_space.py: L2 rate ~2 for DG-P1),
  * symmetry / constant null space of the EMI operator (solver.py:465-466),
  * the rest state of the shipped HH model at the shipped initial data,
  * the electroneutral elimination identity (solver.py:831-838),
and by monomial exactness of the quadrature rules it tabulates."""
import itertools

import numpy as np
import pytest

from knpemidg import mesh as kmesh
from knpemidg.models import mm_hh
from oracle import forms, mms as omms, quadrature as quad, stepper


def test_quadrature_rules_are_exact():
    for rule, dim, deg in ((quad.interval_rule(5), 1, 5), (quad.triangle_deg4(), 2, 4), (quad.triangle_deg5(), 2, 5),
                           (quad.duffy_rule(2, 4), 2, 6), (quad.duffy_rule(3, 4), 3, 5)):
        b, w = rule
        assert abs(w.sum() - 1.0) < 1e-14
        for alpha in itertools.product(range(deg + 1), repeat=dim + 1):
            if sum(alpha) > deg:
                continue
            val = (w * np.prod(b ** np.array(alpha), axis=1)).sum()
            assert abs(val - quad.monomial_integral(alpha)) < 1e-13


def _mms_errors(r):
    mesh, sub, surf = kmesh.mms_mesh(r)
    mm = omms.MMS("space", dt=1e-10)
    P = forms.Problem(mesh, sub.array(), surf.array(), **mm.problem_kwargs())
    c0 = np.stack([mm.exact_field(P, "c", k, t=0.0) for k in range(3)])
    S = stepper.OracleSolver(P, c0, mms=mm, splitting=False)
    S.run(2)                                            # run_MMS_space.py:16-17
    return [mm.l2_error(P, S.c[0], "c", 0), mm.l2_error(P, S.c[1], "c", 1),
            mm.l2_error(P, S.phi, "phi", mean_free=True)]


def test_mms_space_convergence_rate():
    e = np.array([_mms_errors(r) for r in (3, 4, 5)])
    rates = np.log(e[:-1] / e[1:]) / np.log(2.0)
    assert np.all(rates[-1] > 1.85) and np.all(rates[-1] < 2.2), rates


def test_emi_operator_structure():
    mesh, sub, surf = kmesh.neuron_2d_mesh(0)
    P = forms.Problem(mesh, sub.array(), surf.array(), F=96485.0, R=8.314, T=300.0, C_M=0.02, C_phi=200.0,
                      dt=1e-4, z=[1.0, -1.0, 1.0], D_sub=[{0: 2e-9, 1: 1e-9}] * 3, membrane_tags=(1,))
    rng = np.random.default_rng(0)
    c = 100.0 * (1 + 0.1 * rng.uniform(-1, 1, (3, P.nc, P.nd)))
    A, B, b = forms.assemble_emi(P, c, rng.standard_normal(P.nm), None)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    assert np.abs(A @ np.ones(P.ndof)).max() < 1e-10 * abs(A).max()
    assert abs(b.sum()) < 1e-10 * np.abs(b).sum()                    # compatible right-hand side
    w = np.linalg.eigvalsh(B.toarray())
    assert w.min() > 0                                                 # B = A + mass shift is SPD


def test_rest_state_known_answer():
    """HH + pump at the shipped initial data is a steady state (run_2D.py:81-87, mm_hh.py:12-15)."""
    y = mm_hh.init_state_values()
    p = mm_hh.init_parameter_values()
    R, T, F = 8.314, 300.0, 96485.0
    Na_i, Na_e, K_i, K_e = 12.838513108648856, 100.71925900027354, 124.15397583491901, 3.3236967382705265
    p[mm_hh.parameter_indices("E_Na")] = R * T / F * np.log(Na_e / Na_i)
    p[mm_hh.parameter_indices("E_K")] = R * T / F * np.log(K_e / K_i)
    p[mm_hh.parameter_indices("K_e")] = K_e
    p[mm_hh.parameter_indices("Na_i")] = Na_i
    p[mm_hh.parameter_indices("Cm")] = 0.02
    dy = np.zeros(4)
    mm_hh.rhs_numba.py_func(0.2, y, dy, p)                            # t > 0.125: no stimulus
    assert np.abs(dy[:3]).max() < 1e-9 and abs(dy[3]) < 1e-6


def test_electroneutral_elimination():
    mesh, sub, surf = kmesh.mms_mesh(2)
    P = forms.Problem(mesh, sub.array(), surf.array(), F=1, R=1, T=1, C_M=1, C_phi=1, dt=1, z=[1.0, -1.0, 2.0],
                      D_sub=[{0: 1, 1: 1}] * 3, rho_sub={0: 0.5, 1: -0.25})
    rng = np.random.default_rng(1)
    c = rng.uniform(1, 2, (2, P.nc, P.nd))
    ce = forms.eliminated_concentration(P, c)
    total = 1.0 * c[0] - 1.0 * c[1] + 2.0 * ce + P.rho[:, None]
    assert np.abs(total).max() < 1e-14
