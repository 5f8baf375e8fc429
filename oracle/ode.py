"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of MembraneModel.step_lsoda (src/knpemidg/membrane.py:84-119):
per membrane facet ("row"), optional stimulus overwrite of parameter columns,
then LSODA from self.time to self.time+dt with rtol=1e-8, atol=0 and a fresh
integrator per call; the right-hand side stores the channel currents into the
parameter row as a side effect (e.g. examples/idealized-geometries/mm_hh.py:154-159).

Third-party dependency: numbalsoda (unpinned, pyproject.toml:14) - a C++ port
of ODEPACK LSODA - is absent from /root/reference and from this image; scipy's
`LSODA` (the Fortran ODEPACK original) stands in.  Pins: the reference's own
membrane.py stepping its own mm_hh.py / mm_glial.py through this same integrator inside
the reference-executed runs (tests/golden/ref_run_*.npz), the calibration known answer the
reference hard-codes (examples/emix-simulations/mm_hh.py:11-14), the rest state.

Two conventions for the currents handed to the PDEs are provided
(SURVEY.md Appendix E): 'last_call' (reference behaviour: whatever the last
RHS evaluation wrote) and 'end_state' (I(y(t+dt), t+dt); the GPU kernel's
convention).
"""
import numpy as np
from scipy.integrate import solve_ivp


def rhs_python(module):
    f = getattr(module, "rhs_numba", module)
    for attr in ("py_func", "_pyfunc"):
        if hasattr(f, attr):
            return getattr(f, attr)
    return f


def step_rows(module, states, parameters, t0, dt, stim_mask=None, stimulus=None,
              rtol=1.0e-8, atol=0.0, current_convention="end_state", method="LSODA"):
    """In-place step of all rows; returns number of RHS evaluations."""
    rhs = rhs_python(module)
    ns = states.shape[1]
    nfev = 0
    for row in range(states.shape[0]):
        p = parameters[row]
        if stimulus and (stim_mask is None or stim_mask[row]):
            for key, value in stimulus.items():                 # membrane.py:102-104
                p[module.parameter_indices(key)] = value

        def f(t, y, p=p):
            dy = np.empty(ns)
            rhs(t, y, dy, p)
            return dy

        sol = solve_ivp(f, (t0, t0 + dt), states[row].copy(), method=method, rtol=rtol,
                        atol=atol if atol > 0 else 1e-300)
        assert sol.success                                       # membrane.py:113
        nfev += sol.nfev
        states[row, :] = sol.y[:, -1]
        if current_convention == "end_state":
            dy = np.empty(ns)
            rhs(t0 + dt, states[row], dy, p)
    return nfev
