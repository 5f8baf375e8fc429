"""Pins the bundled membrane models against golden vectors generated from the
REFERENCE's own modules (tests/golden/make_ode_golden.py imports
/root/reference/examples/*/mm_*.py unmodified): default tables, name->index maps and
right-hand-side outputs (dy and the I_ch_* side effects in the parameter row).

The compiled device functions are generated from the same Python source
(knpemidg/odegen.py), so the host-emulation build is checked against the same vectors
through a dt -> 0 ODE step: (y(t+dt) - y(t)) / dt -> dy."""
import importlib
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ode_rhs_golden.json")))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_bundled_model_matches_reference_tables_and_rhs(name):
    g = GOLD[name]
    mod = importlib.import_module("knpemidg.models." + name)
    assert np.array_equal(mod.init_state_values(), np.array(g["init_states"]))
    assert np.array_equal(mod.init_parameter_values(), np.array(g["init_parameters"]))
    for key, idx in g["state_index"].items():
        assert mod.state_indices(key) == idx
    for key, idx in g["parameter_index"].items():
        assert mod.parameter_indices(key) == idx
    rhs = mod.rhs_numba.py_func
    for case in g["cases"]:
        y = np.array(case["y"]); p = np.array(case["p_in"]); dy = np.zeros_like(y)
        rhs(case["t"], y, dy, p)
        np.testing.assert_allclose(dy, case["dy"], rtol=1e-13, atol=0)
        np.testing.assert_allclose(p, case["p_out"], rtol=1e-13, atol=0)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_compiled_rhs_matches_reference(emu_lib, name):
    """the generated C function (same text nvcc compiles) through the library's ODE step"""
    from common import Case, _lib
    g = GOLD[name]
    mid, ns, npar = emu_lib.models()[name]
    cs = Case("2d", emu_lib)
    ctx = cs.ctx
    mod = importlib.import_module("knpemidg.models." + name)
    ncase = len(g["cases"])
    y0 = np.array([c["y"] for c in g["cases"]])
    p0 = np.array([c["p_in"] for c in g["cases"]])
    coupled = "V" in g["state_index"]                      # mm_calibration: free-standing, no V / I_ch_*
    ich = [mod.parameter_indices("I_ch_" + n) for n in ("K", "Cl", "Na")] if coupled else []
    for k, case in enumerate(g["cases"]):
        h = ctx.membrane_register(mid, [k], y0[k:k + 1], p0[k:k + 1])
        ctx.membrane_outputs(h, mod.state_indices("V") if coupled else 0, ich)
        scale = np.abs(np.array(case["dy"])).max()
        dt = 1e-7 * np.abs(y0[k]).max() / scale if scale > 0 else 1e-9
        ctx.ode_step(h, case["t"], dt, rtol=1e-10, atol=0.0, set_v=False)
        y1 = ctx.membrane_get(h, "states", (1, ns))[0]
        p1 = ctx.membrane_get(h, "params", (1, npar))[0]
        np.testing.assert_allclose((y1 - y0[k]) / dt, case["dy"], rtol=1e-3, atol=1e-5 * scale)   # first-order difference quotient
        # currents written by the final RHS evaluation at (t+dt, y1): ~ golden up to O(dt)
        if ich:
            np.testing.assert_allclose(p1[ich], np.array(case["p_out"])[ich], rtol=1e-3,
                                       atol=1e-5 * np.abs(np.array(case["p_out"])[ich]).max() + 1e-300)
