"""Generates tests/golden/ode_rhs_golden.json from the REFERENCE's own membrane-model
modules (examples/*/mm_*.py), imported unmodified from /root/reference with the
`numbalsoda` signature shim (numba is present, numbalsoda is not).

For every module: the default state / parameter tables, the name->index maps and the
right-hand side (`rhs_numba`, the function the reference hands to LSODA,
src/knpemidg/membrane.py:88) evaluated at seeded inputs: outputs = dy and the parameter row
after the call (the RHS stores the channel currents I_ch_* into it).

Run in the build container (the reference is not available on the GPU box):
    python tests/golden/make_ode_golden.py
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "knp-emi-dg_b200"))      # numbalsoda shim
REF = "/root/reference/examples"
if not hasattr(np, "float_"):          # the reference predates numpy 2 (np.float_ was removed)
    np.float_ = np.float64

# reference module -> bundled model that restates it
MODULES = {
    "idealized-geometries/mm_hh.py": "mm_hh",
    "idealized-geometries/mm_hh_no_stim.py": "mm_hh_no_stim",
    "rat-neuron/mm_leak.py": "mm_leak",
    "emix-simulations/mm_hh.py": "mm_hh_emix",
    "emix-simulations/mm_glial.py": "mm_glial_emix",
    "local-astrocyte-depolarization/mm_hh.py": "mm_hh_astro",
    "local-astrocyte-depolarization/mm_glial.py": "mm_glial_astro",
    "emix-simulations/mm_calibration.py": "mm_calibration",          # appended last: the seeded stream above is unchanged
}
STATE_NAMES = ["m", "h", "n", "V", "V_n", "V_g", "K_e", "K_n", "K_g", "Na_e", "Na_n", "Na_g"]


def load(path):
    spec = importlib.util.spec_from_file_location("ref_" + os.path.basename(path)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def names_of(indexer, candidates):
    out = {}
    for name in candidates:
        try:
            out[name] = int(indexer(name))
        except Exception:
            pass
    return out


PARAM_CANDIDATES = ["g_Na_bar", "g_K_bar", "g_leak_Na", "g_leak_K", "g_leak_Cl", "E_Na", "E_K", "E_Cl", "Cm",
                    "stim_amplitude", "I_ch_Na", "I_ch_K", "I_ch_Cl", "K_e", "Na_i", "m_K", "m_Na", "I_max",
                    "K_e_init", "K_i_init", "Na_i_init", "g_Kir", "E_K_init", "rho_pump", "P_Nai", "P_Ke",
                    "k_dec", "Cl_i", "Cl_e", "K_i", "Na_e", "g_Cl_leak", "g_K_leak", "g_Na_leak", "stim_start",
                    "stim_end", "g_syn_bar", "T", "F", "R", "z_K", "z_Na", "z_Cl", "psi", "E_Kir", "g_KCC1",
                    "i_pump", "g_leak", "E_leak", "phi_rest", "phi_M_init",
                    "g_leak_Na_n", "g_leak_K_n", "g_leak_Na_g", "g_leak_K_g", "I_max_n", "I_max_g"]


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    for rel, bundled in MODULES.items():
        mod = load(os.path.join(REF, rel))
        rhs = mod.rhs_numba._pyfunc if hasattr(mod.rhs_numba, "_pyfunc") else mod.rhs_numba.py_func
        s0 = np.asarray(mod.init_state_values(), dtype=float)
        p0 = np.asarray(mod.init_parameter_values(), dtype=float)
        cases = []
        mv = abs(s0[-1]) > 1.0          # mV / ms units
        for _ in range(6):
            y = s0 * (1.0 + 0.2 * rng.uniform(-1, 1, s0.size))
            p = p0.copy()
            pidx = names_of(mod.parameter_indices, PARAM_CANDIDATES)
            sc = 1e3 if mv else 1.0
            for key, val in (("E_Na", 0.054 * sc), ("E_K", -0.088 * sc), ("E_Cl", -0.07 * sc), ("K_e", 4.0),
                             ("Na_i", 12.0), ("Cm", 1.0 if mv else 0.02), ("stim_amplitude", 0.5 if mv else 10.0)):
                if key in pidx:
                    p[pidx[key]] = val * (1.0 + 0.1 * rng.uniform(-1, 1))
            t = float(rng.uniform(0, 0.05 * (1e3 if mv else 1.0)))
            dy = np.zeros_like(y)
            p_in = p.copy()
            rhs(t, y, dy, p)
            cases.append({"t": t, "y": y.tolist(), "p_in": p_in.tolist(), "dy": dy.tolist(), "p_out": p.tolist()})
        out[bundled] = {"reference_file": "examples/" + rel, "init_states": s0.tolist(), "init_parameters": p0.tolist(),
                        "state_index": names_of(mod.state_indices, STATE_NAMES),
                        "parameter_index": names_of(mod.parameter_indices, PARAM_CANDIDATES), "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ode_rhs_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, {k: len(v["cases"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
