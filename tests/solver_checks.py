"""Checks of the reference-facing Python API (knpemidg.Solver / MembraneModel / utils),
written the way the reference's run scripts use it (examples/idealized-geometries/run_2D.py,
tests/run_MMS_space.py), shared by the CPU (emulation) and GPU test modules."""
from collections import namedtuple

import numpy as np

from common import kmesh, rel_err
from knpemidg import Solver, plus, minus, pcws_constant_project
from knpemidg.frontend import Constant
from knpemidg.models import mm_hh
from oracle import forms, mms as omms, stepper

DT, C_M, F, R, T = 1e-4, 0.02, 96485, 8.314, 300
NA_I, NA_E, K_I, K_E = 12.838513108648856, 100.71925900027354, 124.15397583491901, 3.3236967382705265

SolverParams = namedtuple("solver_params", "direct_emi direct_knp resolution rtol_emi rtol_knp atol_emi atol_knp "
                                           "threshold_emi threshold_knp")


class Solver2D(Solver):
    """as examples/idealized-geometries/run_2D.py:30-50"""

    def __init__(self, params, ion_list, **kw):
        Solver.__init__(self, params, ion_list, degree_emi=1, degree_knp=1, mms=None, sf=1, **kw)

    def update_ode(self, ode_model):
        K_e = plus(self.c_prev_k.split()[0], self.n_g)
        ode_model.set_parameter("K_e", pcws_constant_project(K_e, self.Q))
        Na_i = minus(self.ion_list[-1]["c"], self.n_g)
        ode_model.set_parameter("Na_i", pcws_constant_project(Na_i, self.Q))


def _ion(name, z, D, ci, ce):
    return {"c_init_sub": {1: Constant(ci), 0: Constant(ce)}, "c_init_sub_type": "constant",
            "bdry": Constant((0, 0)), "z": z, "name": name, "D_sub": {1: Constant(D), 0: Constant(D)},
            "f_source": Constant(0)}


def run_2d_neuron(lib, nsteps, rtol_emi=1e-5, rtol_knp=1e-7, outdir=None, g_syn=10.0, resolution=0, trace=None):
    params = namedtuple("params", "dt n_steps_ODE F psi phi_M_init C_phi C_M R temperature phi_M_init_type "
                                  "rho_sub")(DT, 25, F, F / (R * T), Constant(-0.0743), C_M / DT, C_M, R, T,
                                             "constant", {0: Constant(0), 1: Constant(0)})
    ion_list = [_ion("K", 1.0, 1.96e-9, K_I, K_E), _ion("Cl", -1.0, 2.03e-9, NA_I + K_I, NA_E + K_E),
                _ion("Na", 1.0, 1.33e-9, NA_I, NA_E)]
    stim = namedtuple("membrane_params", "g_syn_bar stimulus stimulus_locator")(
        g_syn, {"stim_amplitude": g_syn}, lambda x: x[0] < 20e-6)
    sp = SolverParams(False, False, 0, rtol_emi, rtol_knp, 1e-40, 1e-40, None, None)
    mesh, sub, surf = kmesh.neuron_2d_mesh(resolution)
    S = Solver2D(params, ion_list, lib=lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(stim, {1: mm_hh})
    t = Constant(0.0)
    S.solve_system_active(nsteps * DT, t, sp, filename=outdir, save_fields=outdir is not None,
                          save_solver_stats=outdir is not None)
    assert abs(float(t) - nsteps * DT) < 1e-15
    # oracle run of the same problem
    kw = dict(F=F, R=R, T=T, C_M=C_M, C_phi=C_M / DT, dt=DT, z=[1.0, -1.0, 1.0],
              D_sub=[{0: 1.96e-9, 1: 1.96e-9}, {0: 2.03e-9, 1: 2.03e-9}, {0: 1.33e-9, 1: 1.33e-9}],
              rho_sub={0: 0.0, 1: 0.0})
    P = forms.Problem(mesh, sub.array(), surf.array(), membrane_tags=(1,), **kw)
    cinit = [{1: K_I, 0: K_E}, {1: NA_I + K_I, 0: NA_E + K_E}, {1: NA_I, 0: NA_E}]
    c0 = np.stack([np.where((sub.array() == 1)[:, None], ci[1], ci[0]) * np.ones((P.nc, P.nd)) for ci in cinit])
    O = stepper.OracleSolver(P, c0, models={1: mm_hh}, stimulus={"stim_amplitude": g_syn},
                             stimulus_locator=lambda x: x[0] < 20e-6, ion_names=["K", "Cl", "Na"])
    if trace is None:
        O.run(nsteps)
    else:                                    # membrane-potential trace of the oracle, step by step
        for _ in range(nsteps):
            O.step()
            trace.append(O.phi_M.copy())
    return S, O


class MMSLoads:
    """manufactured-solution data for Solver(mms=...): load vectors of the MMS-only terms
    (solver.py:365-374, 645-657) computed by the oracle's sympy restatement of
    tests/mms_space.py, plus the exact fields for the error norms."""

    def __init__(self, mesh, sub, surf, dt, kind="space", ufl_degree=None):
        self.mm = omms.MMS(kind, dt=dt, ufl_degree=ufl_degree)
        self.P = forms.Problem(mesh, sub.array(), surf.array(), **self.mm.problem_kwargs())

    def load_emi(self, t):
        self.mm.t = t
        return self.mm.emi_rhs(self.P)

    def load_knp(self, k, t):
        self.mm.t = t
        return self.mm.knp_rhs(self.P, k)


def run_mms(lib, r, dt=1e-10, nsteps=2, ufl_degree=None):
    """tests/run_MMS_space.py: passive system, two steps, direct solves"""
    mesh, sub, surf = kmesh.mms_mesh(r)
    L = MMSLoads(mesh, sub, surf, dt, ufl_degree=ufl_degree)
    mm, P = L.mm, L.P
    params = namedtuple("params", "dt F psi C_phi C_M R temperature phi_M_init_type rho_sub")(
        dt, 1.0, 1.0, 1.0 / dt, 1.0, 1.0, 1.0, "constant", {0: Constant(0), 1: Constant(0)})
    exact = [mm.exact_field(P, "c", k, t=0.0) for k in range(3)]
    ion_list = []
    for k, name in enumerate("abc"):
        ion_list.append({"c_init_sub": exact[k].ravel(), "c_init_sub_type": "function", "z": mm.z[k], "name": name,
                         "D_sub": {1: Constant(mm.D1[k]), 0: Constant(mm.D2[k])},
                         "C_sub": {1: Constant(mm.C1[k]), 0: Constant(mm.C2[k])}, "f_source": Constant(0)})
    S = Solver(params, ion_list, mms=L, lib=lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    sp = SolverParams(True, True, r, None, None, None, None, None, None)
    t = Constant(0.0)
    uh, c_elim = S.solve_system_passive(nsteps * dt, t, sp, None)
    errs = [mm.l2_error(P, uh[0].nodal(), "c", 0), mm.l2_error(P, uh[1].nodal(), "c", 1),
            mm.l2_error(P, uh[2].nodal(), "phi", mean_free=True)]
    return np.array(errs), S, L


def run_mms_time(lib, i, r=3, dt0=1.0e-2, script_initial_data=False):
    """tests/run_MMS_time.py: fixed mesh, dt = dt0 / 2^i, Tstop = 2 dt0, passive system, direct
    solves; the exact solution is linear in space, so the error is the time error.
    script_initial_data: start the eliminated ion from the script's own k_c1_init (mms_time.py:48), which has the
    0.2 / 0.3 of k_a1 / k_b1 swapped and therefore is -0.1 in the ICS where the exact solution is +0.1 (it only
    enters the conductivity of the first EMI solve) - needed to compare with the reference-executed numbers"""
    dt = dt0 / 2 ** i
    nsteps = int(round(2 * dt0 / dt))
    mesh, sub, surf = kmesh.mms_mesh(r)
    L = MMSLoads(mesh, sub, surf, dt, kind="time")
    mm, P = L.mm, L.P
    params = namedtuple("params", "dt F psi C_phi C_M R temperature phi_M_init_type rho_sub")(
        dt, 1.0, 1.0, 1.0 / dt, 1.0, 1.0, 1.0, "expression", {0: Constant(0), 1: Constant(0)})
    exact = [mm.exact_field(P, "c", k, t=0.0) for k in range(3)]
    if script_initial_data:
        exact[2] = np.where((P.cell_tag == 1)[:, None], -0.1, 0.0) * np.ones_like(exact[2])
    ion_list = []
    for k, name in enumerate("abc"):
        ion_list.append({"c_init_sub": exact[k].ravel(), "c_init_sub_type": "function", "z": mm.z[k], "name": name,
                         "D_sub": {1: Constant(mm.D1[k]), 0: Constant(mm.D2[k])},
                         "C_sub": {1: Constant(mm.C1[k]), 0: Constant(mm.C2[k])}, "f_source": Constant(0)})
    S = Solver(params, ion_list, mms=L, lib=lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    sp = SolverParams(True, True, r, None, None, None, None, None, None)
    t = Constant(0.0)
    uh, c_elim = S.solve_system_passive(nsteps * dt, t, sp, None)
    mm.t = float(t)
    errs = [mm.l2_error(P, uh[0].nodal(), "c", 0), mm.l2_error(P, uh[1].nodal(), "c", 1),
            mm.l2_error(P, uh[2].nodal(), "phi", mean_free=True)]
    return np.array(errs), S, L


def run_emix_block(lib, M, nsteps, rtol_emi=1e-12, rtol_knp=1e-13):
    """bench.py's second workload (BASELINE configs[4]: EMIx-like block, three cell tags, mm_glial on tag 1 and
    mm_hh on tag 2, ms / cm / mV units, calibrated initial state, synaptic stimulus) through the engine
    and through the oracle; returns (engine, oracle)"""
    import bench
    from knpemidg.models import mm_glial_emix, mm_hh_emix
    eng = bench.build_engine_emix(M, 0, lib=lib)
    eng.rtol_emi, eng.rtol_knp = rtol_emi, rtol_knp
    mesh = eng.global_mesh
    sub, surf = eng._global_cell_tags, None
    mesh2, sub2, surf2 = kmesh.emix_like_mesh(M, n_cells=100, length=1.0e-3)          # same seeded generator
    assert np.array_equal(mesh2.cells, mesh.cells) and np.array_equal(sub2.array(), sub)
    P = forms.Problem(mesh2, sub2.array(), surf2.array(), membrane_tags=(1, 2), **bench.EMIX_PHYS)
    c0 = np.stack([np.choose(sub2.array(), [ci[0], ci[1], ci[2]])[:, None] * np.ones((P.nc, P.nd))
                   for ci in bench.EMIX_C_INIT])
    O = stepper.OracleSolver(P, c0, models={1: mm_glial_emix, 2: mm_hh_emix}, stimulus={"stim_amplitude": 5.0},
                             stimulus_locator=lambda x: x[0] < 3.0e-4, ion_names=["K", "Cl", "Na"])
    for _ in range(nsteps):
        eng.step()
    O.run(nsteps)
    return eng, O


def check_picard(lib):
    """solve_for_time_step_picard (solver.py:850-927): converges in a few iterations at the
    reference's time step and stays close to the split step it refines"""
    import bench
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh
    from common import kmesh

    def make():
        mesh, sub, surf = kmesh.neuron_2d_mesh(1)
        eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1,), lib=lib, **bench.PHYS)
        eng.set_concentrations_by_tag(bench.C_INIT)
        eng.add_membrane_model(1, mm_hh, bench.ION_NAMES, stimulus=bench.STIMULUS, stimulus_locator=bench.stim_locator)
        eng.rtol_emi, eng.rtol_knp = 1e-10, 1e-11
        eng.initialize(pc=1)
        return eng
    a, b = make(), make()
    for _ in range(3):
        a.step()
        b.ode_phase()
        it = b.pde_phase_picard()
        b.k += 1
        assert 1 <= it <= 6
    assert rel_err(b.phi_M(), a.phi_M()) < 1e-3
    for k in range(3):
        assert rel_err(b.concentration(k), a.concentration(k)) < 1e-4
    # the fixed point: one more Picard sweep from the converged state changes nothing above tol
    assert b.picard_iterations <= 3


def _neuron_engine(lib, resolution=1, tight=True):
    import bench
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh
    from common import kmesh
    mesh, sub, surf = kmesh.neuron_2d_mesh(resolution)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1,), lib=lib, **bench.PHYS)
    eng.set_concentrations_by_tag(bench.C_INIT)
    eng.add_membrane_model(1, mm_hh, bench.ION_NAMES, stimulus=bench.STIMULUS, stimulus_locator=bench.stim_locator)
    if tight:
        eng.rtol_emi, eng.rtol_knp = 1e-10, 1e-11
    eng.initialize(pc=1)
    return eng


def check_picard_then_regular(lib):
    """a Picard step writes c_prev_n separately (KNP_F_CN); the regular steps after it must again
    take the time derivative against the concentrations of the step before (c_prev_n.assign(c),
    solver.py:810) - not against the ones the Picard step left behind"""
    from knpemidg import _lib
    a, b = _neuron_engine(lib), _neuron_engine(lib)
    for eng in (a, b):
        eng.step()
    # b: one Picard step; a: the same state transplanted into it afterwards
    b.ode_phase()
    b.pde_phase_picard()
    b.k += 1
    a.ode_phase()
    a.pde_phase_picard()
    a.k += 1
    fresh = _neuron_engine(lib)                          # never touched KNP_F_CN
    for k in range(3):
        fresh.ctx.set_field(_lib.F_C, k, a.ctx.get_field(_lib.F_C, k))
    fresh.ctx.set_field(_lib.F_PHI, 0, a.ctx.get_field(_lib.F_PHI))
    fresh.ctx.set_field(_lib.F_PHIM, 0, a.ctx.get_field(_lib.F_PHIM))
    for k in range(3):
        fresh.ctx.set_field(_lib.F_ICH, k, a.ctx.get_field(_lib.F_ICH, k))
    fresh.ctx.post_step(_lib.POST_NERNST)
    for _ in range(2):                                   # regular PDE steps (no ODE: same currents on both sides)
        b.pde_phase()
        fresh.pde_phase()
        for k in range(2):
            assert np.array_equal(b.ctx.get_field(_lib.F_CN, k), b.ctx.get_field(_lib.F_C, k))
    for k in range(3):
        assert rel_err(b.concentration(k), fresh.concentration(k)) < 1e-9
    assert rel_err(b.phi_M(), fresh.phi_M()) < 1e-7


def check_membrane_shape_validation(lib):
    """tables whose shape does not match the compiled model are refused before any pointer is taken;
    a module is matched to a compiled model by what it computes, not by its name"""
    import types
    import pytest
    from knpemidg import _lib
    from knpemidg.models import mm_hh, mm_leak
    eng = _neuron_engine(lib, resolution=0, tight=False)
    mid, ns, npar = lib.models()["mm_hh"]
    rows = np.arange(4, dtype=np.int32)
    with pytest.raises(_lib.KnpError):
        eng.ctx.membrane_register(mid, rows, np.zeros((4, ns + 1)), np.zeros((4, npar)))
    with pytest.raises(_lib.KnpError):
        eng.ctx.membrane_register(mid, rows, np.zeros((4, ns)), np.zeros((3, npar)))
    h = eng.members[0].handle
    with pytest.raises(_lib.KnpError):
        eng.ctx.membrane_set(h, "params", np.zeros(5))
    # a module CALLED mm_hh that is in fact the leak model resolves to the compiled leak model
    fake = types.ModuleType("mm_hh")
    for name in ("init_state_values", "init_parameter_values", "state_indices", "parameter_indices", "rhs_numba"):
        setattr(fake, name, getattr(mm_leak, name))
    assert eng.resolve_model(fake) == "mm_leak"
    assert eng.resolve_model(mm_hh) == "mm_hh"


def check_solver_emi(lib):
    """SolverEMI (solver_emi.py): the run-script flow with the EMI sub-problem only - the
    membrane fires, the concentrations never move"""
    from collections import namedtuple
    from knpemidg import SolverEMI
    from knpemidg.frontend import Constant
    from knpemidg.models import mm_hh
    from common import kmesh

    class EMI2D(SolverEMI):
        def update_ode(self, ode_model):
            Solver2D.update_ode(self, ode_model)

    params = namedtuple("params", "dt n_steps_ODE F psi phi_M_init C_phi C_M R temperature phi_M_init_type "
                                  "rho_sub")(DT, 25, F, F / (R * T), Constant(-0.0743), C_M / DT,
                                             C_M, R, T, "constant", {0: Constant(0), 1: Constant(0)})
    ion_list = [_ion("K", 1.0, 1.96e-9, K_I, K_E), _ion("Cl", -1.0, 2.03e-9, NA_I + K_I, NA_E + K_E),
                _ion("Na", 1.0, 1.33e-9, NA_I, NA_E)]
    stim = namedtuple("membrane_params", "g_syn_bar stimulus stimulus_locator")(
        10.0, {"stim_amplitude": 10.0}, lambda x: x[0] < 20e-6)
    sp = SolverParams(False, False, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None)
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    S = EMI2D(params, ion_list, lib=lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(stim, {1: mm_hh})
    c0 = [S.c.split()[k].nodal().copy() for k in range(2)]
    t = Constant(0.0)
    S.solve_system_active(30 * DT, t, sp)
    pm = S.phi_M_prev_PDE.vector().get_local()
    assert pm.max() > 0.0                                         # the spike (peak at step ~29 in the full model)
    for k in range(2):
        assert np.array_equal(S.c.split()[k].nodal(), c0[k])
    assert all(n == 0 for n in S.engine.stats["knp_niter"]) and abs(float(t) - 30 * DT) < 1e-15


def check_solver_emi_against_reference(lib, rtol=1e-6):
    """the run-script flow of SolverEMI, 25 steps of the 2D neuron, against tests/golden/ref_run_2d_emi.npz:
    membrane-potential traces of the reference's own solver_emi.py (direct solves there, tight CG here)"""
    import os
    from knpemidg import SolverEMI
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_run_2d_emi.npz"))
    trace = []

    class EMI2D(SolverEMI):
        def update_ode(self, ode_model):
            Solver2D.update_ode(self, ode_model)

        def solve_for_time_step(self, k, t):
            SolverEMI.solve_for_time_step(self, k, t)
            trace.append(self.phi_M_prev_PDE.vector().get_local().copy())

    params = namedtuple("params", "dt n_steps_ODE F psi phi_M_init C_phi C_M R temperature phi_M_init_type "
                                  "rho_sub")(DT, 25, F, F / (R * T), Constant(-0.0743), C_M / DT, C_M, R, T, "constant",
                                             {0: Constant(0), 1: Constant(0)})
    ion_list = [_ion("K", 1.0, 1.96e-9, K_I, K_E), _ion("Cl", -1.0, 2.03e-9, NA_I + K_I, NA_E + K_E),
                _ion("Na", 1.0, 1.33e-9, NA_I, NA_E)]
    stim = namedtuple("membrane_params", "g_syn_bar stimulus stimulus_locator")(
        10.0, {"stim_amplitude": 10.0}, lambda x: x[0] < 20e-6)
    sp = SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None)     # "direct": iterate to ~1e-11
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    S = EMI2D(params, ion_list, lib=lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(stim, {1: mm_hh})
    t = Constant(0.0)
    n = int(g["nsteps"])
    S.solve_system_active(n * DT, t, sp)
    ref = g["phi_M_trace"]
    got = np.stack(trace)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / (ref.max() - ref.min()) < rtol


def check_passive_run_against_reference(lib, rtol=1e-6):
    """Solver.solve_system_passive (PDE steps only, non-splitting Robin forms) against the reference's own
    solve_system_passive executed on oracle/refexec (tests/golden/ref_run_2d_passive.npz): 10 steps from a
    perturbed membrane potential"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_run_2d_passive.npz"))
    trace = []

    class Passive2D(Solver2D):
        def solve_for_time_step(self, k, t):
            Solver2D.solve_for_time_step(self, k, t)
            trace.append(self.phi_M_prev_PDE.vector().get_local().copy())

    params = namedtuple("params", "dt n_steps_ODE F psi phi_M_init C_phi C_M R temperature phi_M_init_type "
                                  "rho_sub")(DT, 25, F, F / (R * T), Constant(-0.0743), C_M / DT, C_M, R, T, "constant",
                                             {0: Constant(0), 1: Constant(0)})
    ion_list = [_ion("K", 1.0, 1.96e-9, K_I, K_E), _ion("Cl", -1.0, 2.03e-9, NA_I + K_I, NA_E + K_E),
                _ion("Na", 1.0, 1.33e-9, NA_I, NA_E)]
    stim = namedtuple("membrane_params", "g_syn_bar stimulus stimulus_locator")(
        10.0, {"stim_amplitude": 10.0}, lambda x: x[0] < 20e-6)
    sp = SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None)
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    S = Passive2D(params, ion_list, lib=lib)
    S.setup_domain(mesh, sub, surf)
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(stim, {1: mm_hh})
    S.phi_M_prev_PDE.vector().set_local(g["phi_M0"])
    t = Constant(0.0)
    uh, c_elim = S.solve_system_passive(int(g["nsteps"]) * DT, t, sp, None)
    ref = g["phi_M_trace"]
    got = np.stack(trace)
    assert np.abs(got - ref).max() / (np.abs(ref).max()) < rtol
    n = ref.shape[1]
    cfin = np.concatenate([uh[k].nodal().ravel() for k in range(2)])
    assert rel_err(cfin, g["final_c"]) < 1e-8
    assert rel_err(c_elim.nodal().ravel(), g["final_c_elim"]) < 1e-8
    phi = uh[-1].nodal().ravel()
    assert np.abs((phi - phi.mean()) - (g["final_phi"] - g["final_phi"].mean())).max() < 1e-7 * np.abs(g["final_phi"]).max()
