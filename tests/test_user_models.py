"""User-supplied membrane models (the reference accepts any module that follows the mm_*.py
protocol, membrane.py:88): a module that is not bundled is translated to device code and compiled
into a library variant by Solver.setup_membrane_model (knpemidg._lib.variant_with)."""
import os
import sys
import textwrap
from collections import namedtuple

import numpy as np
import pytest

import solver_checks as sc
from common import kmesh
from knpemidg import _lib
from knpemidg.frontend import Constant

USER_MODEL = '''
    """leak membrane with a slow adaptation variable - not one of the bundled models"""
    import numpy as np
    from numbalsoda import lsoda_sig
    from numba import cfunc

    def init_state_values(**values):
        return np.array([-0.0743, 0.1], dtype=np.float64)

    def init_parameter_values(**values):
        return np.array([2.0, 0.5, 0.0, 0.0, 0.02, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 50.0], dtype=np.float64)

    _S = {"V": 0, "w": 1}
    _P = {"g_K": 0, "g_Na": 1, "E_Na": 2, "E_K": 3, "Cm": 4, "stim_amplitude": 5, "I_ch_Na": 6, "I_ch_K": 7,
          "I_ch_Cl": 8, "K_e": 9, "Na_i": 10, "E_Cl": 11, "tau_w": 12}

    def state_indices(*names):
        idx = [_S[n] for n in names]
        return idx[0] if len(idx) == 1 else idx

    def parameter_indices(*names):
        idx = [_P[n] for n in names]
        return idx[0] if len(idx) == 1 else idx

    @cfunc(lsoda_sig, nopython=True)
    def rhs_numba(t, states, values, parameters):
        i_Na = (parameters[1] + parameters[5] * np.exp(-np.mod(t, 0.03) / 0.002)) * (states[0] - parameters[2])
        i_K = parameters[0] * (1.0 + states[1]) * (states[0] - parameters[3])
        parameters[6] = i_Na
        parameters[7] = i_K
        parameters[8] = 0.0
        values[0] = -(i_K + i_Na) / parameters[4]
        values[1] = parameters[12] * (np.tanh(100.0 * (states[0] + 0.06)) - states[1])
'''


@pytest.fixture
def user_module(tmp_path, monkeypatch):
    (tmp_path / "mm_user_adapt.py").write_text(textwrap.dedent(USER_MODEL))
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.setenv("KNPEMIDG_VARIANT_DIR", str(tmp_path / "variants"))     # keep the source tree clean
    sys.modules.pop("mm_user_adapt", None)
    import mm_user_adapt
    yield mm_user_adapt
    sys.modules.pop("mm_user_adapt", None)


def test_user_model_is_compiled_and_integrated(emu_lib, user_module):
    lib, names = _lib.variant_with(emu_lib, [user_module])
    name = names[user_module]
    assert name.startswith("user_mm_user_adapt_") and name in lib.models()
    mid, ns, npar = lib.models()[name]
    assert (ns, npar) == (2, 13)
    assert set(emu_lib.models()) < set(lib.models())                 # the bundled models are still there
    # one ODE point: the library's adaptive step against scipy's LSODA on the module's own right-hand side
    from common import Case
    from scipy.integrate import solve_ivp
    cs = Case("2d", lib)
    ctx = cs.ctx
    y0 = user_module.init_state_values()
    p0 = user_module.init_parameter_values()
    p0[[2, 3, 5]] = 0.054, -0.088, 4.0
    h = ctx.membrane_register(mid, [0], y0[None, :], p0[None, :])
    ctx.membrane_outputs(h, 0, [7, 8, 6])
    dt = 1e-3
    ctx.ode_step(h, 0.0, dt, rtol=1e-9, atol=0.0, set_v=False)
    y1 = ctx.membrane_get(h, "states", (1, ns))[0]
    rhs = user_module.rhs_numba.py_func if hasattr(user_module.rhs_numba, "py_func") else user_module.rhs_numba._pyfunc

    def f(t, y):
        dy, p = np.zeros(2), p0.copy()
        rhs(t, y, dy, p)
        return dy
    ref = solve_ivp(f, (0.0, dt), y0, method="LSODA", rtol=1e-10, atol=1e-14).y[:, -1]
    np.testing.assert_allclose(y1, ref, rtol=1e-6)


def test_solver_builds_the_variant_on_demand(emu_lib, user_module):
    """run_2D.py-style flow with the user's module: setup_membrane_model finds no compiled
    counterpart, builds the variant, and the run proceeds on it"""
    params = namedtuple("params", "dt n_steps_ODE F psi phi_M_init C_phi C_M R temperature phi_M_init_type "
                                  "rho_sub")(sc.DT, 25, sc.F, sc.F / (sc.R * sc.T), Constant(-0.0743), sc.C_M / sc.DT,
                                             sc.C_M, sc.R, sc.T, "constant", {0: Constant(0), 1: Constant(0)})
    ion_list = [sc._ion("K", 1.0, 1.96e-9, sc.K_I, sc.K_E), sc._ion("Cl", -1.0, 2.03e-9, sc.NA_I + sc.K_I, sc.NA_E + sc.K_E),
                sc._ion("Na", 1.0, 1.33e-9, sc.NA_I, sc.NA_E)]
    stim = namedtuple("membrane_params", "g_syn_bar stimulus stimulus_locator")(
        4.0, {"stim_amplitude": 4.0}, lambda x: x[0] < 20e-6)
    sp = sc.SolverParams(False, False, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None)
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    S = sc.Solver2D(params, ion_list, lib=emu_lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(stim, {1: user_module})
    assert S.engine.ctx.lib is not emu_lib and S.engine.user_models[user_module].startswith("user_")
    t = Constant(0.0)
    S.solve_system_active(5 * sc.DT, t, sp)
    pm = S.phi_M_prev_PDE.vector().get_local()
    assert np.isfinite(pm).all() and pm.max() > -0.0743 + 1e-4        # the stimulated end depolarises
    w = S.mem_models[0]["ode"].states[:, 1]
    assert np.all(w < 0.1) and np.all(w > -1.0)                       # the adaptation variable relaxes towards tanh(...) < 0


STIFF_MODEL = '''
    """a deliberately STIFF membrane model: the gate w relaxes to its (V-dependent) target with a time
    constant of 1e-7 s, five orders of magnitude below the step it is integrated over"""
    import numpy as np
    from numbalsoda import lsoda_sig
    from numba import cfunc

    def init_state_values(**values):
        return np.array([-0.0743, 0.3], dtype=np.float64)

    def init_parameter_values(**values):
        return np.array([2.0, 0.5, 0.054, -0.088, 0.02, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0e7], dtype=np.float64)

    _S = {"V": 0, "w": 1}
    _P = {"g_K": 0, "g_Na": 1, "E_Na": 2, "E_K": 3, "Cm": 4, "stim_amplitude": 5, "I_ch_Na": 6, "I_ch_K": 7,
          "I_ch_Cl": 8, "K_e": 9, "Na_i": 10, "E_Cl": 11, "rate_w": 12}

    def state_indices(*names):
        idx = [_S[n] for n in names]
        return idx[0] if len(idx) == 1 else idx

    def parameter_indices(*names):
        idx = [_P[n] for n in names]
        return idx[0] if len(idx) == 1 else idx

    @cfunc(lsoda_sig, nopython=True)
    def rhs_numba(t, states, values, parameters):
        i_Na = parameters[1] * (states[0] - parameters[2])
        i_K = parameters[0] * (1.0 + states[1]) * (states[0] - parameters[3])
        parameters[6] = i_Na
        parameters[7] = i_K
        parameters[8] = 0.0
        values[0] = -(i_K + i_Na) / parameters[4]
        values[1] = parameters[12] * (0.5 + 0.4 * np.tanh(50.0 * (states[0] + 0.06)) - states[1])
'''


def test_stiff_user_model_takes_the_implicit_path(emu_lib, tmp_path, monkeypatch):
    """LSODA switches to BDF on a stiff right-hand side (membrane.py:108-112); the library's explicit pair
    detects that its step is stability-limited and finishes the interval with the Rosenbrock pair - same
    answer as scipy's implicit Radau solver, in a bounded number of steps"""
    (tmp_path / "mm_user_stiff.py").write_text(textwrap.dedent(STIFF_MODEL))
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.setenv("KNPEMIDG_VARIANT_DIR", str(tmp_path / "variants"))
    sys.modules.pop("mm_user_stiff", None)
    import mm_user_stiff as mod
    from common import Case
    from scipy.integrate import solve_ivp
    lib, names = _lib.variant_with(emu_lib, [mod])
    mid, ns, npar = lib.models()[names[mod]]
    ctx = Case("2d", lib).ctx
    y0, p0 = mod.init_state_values(), mod.init_parameter_values()
    h = ctx.membrane_register(mid, [0, 1], np.tile(y0, (2, 1)), np.tile(p0, (2, 1)))
    ctx.membrane_outputs(h, 0, [7, 8, 6])
    dt = 1e-2                                   # 1e5 relaxation times of w: ~3e4 explicit steps would be needed
    nsteps, nfev = ctx.ode_step(h, 0.0, dt, rtol=1e-8, atol=0.0, set_v=False)
    assert ctx.ode_stiff_facets == 2
    assert nsteps < 60000
    y1 = ctx.membrane_get(h, "states", (2, ns))
    rhs = mod.rhs_numba.py_func if hasattr(mod.rhs_numba, "py_func") else mod.rhs_numba._pyfunc

    def f(t, y):
        dy, p = np.zeros(2), p0.copy()
        rhs(t, y, dy, p)
        return dy
    ref = solve_ivp(f, (0.0, dt), y0, method="Radau", rtol=1e-11, atol=1e-14).y[:, -1]
    np.testing.assert_allclose(y1[0], ref, rtol=2e-6)
    np.testing.assert_allclose(y1[1], y1[0], rtol=0, atol=0)
    sys.modules.pop("mm_user_stiff", None)
