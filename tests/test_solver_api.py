"""Reference-facing API on the host-emulation library (CPU suite): the run-script flow of
run_2D.py against the oracle, the output files, and the MMS convergence study of
tests/run_MMS_space.py through the product code path."""
import os

import numpy as np

import solver_checks as sc
from common import rel_err


def test_run_2d_flow_matches_oracle(emu_lib, tmp_path):
    out = str(tmp_path) + "/"
    S, O = sc.run_2d_neuron(emu_lib, 4, rtol_emi=1e-12, rtol_knp=1e-13, outdir=out)
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 1e-6      # north_star trace tolerance
    for k in range(2):
        assert rel_err(S.c.split()[k].nodal(), O.c[k]) < 1e-9
    assert rel_err(S.ion_list[-1]["c"].nodal(), O.c_elim) < 1e-9
    for k in range(3):
        assert rel_err(S.ion_list[k]["E"].vector().get_local(), O.E[k]) < 1e-7
    # membrane model tables and stimulus mask (membrane.py:92-104)
    ode = S.mem_models[0]["ode"]
    P = ode.parameters
    col = ode.ode.parameter_indices("stim_amplitude")
    assert np.array_equal(P[:, col] > 0, ode.dof_locations[:, 0] < 20e-6)
    assert abs(ode.time - 4 * sc.DT) < 1e-15
    # statistics files keep the reference's format (solver.py:1146-1198)
    txt = open(os.path.join(out, "solver", "emi_niter_0.txt")).read().splitlines()
    assert txt[0].startswith("num cells:") and txt[1].startswith("dofs:") and len(txt) == 2 + 4
    # field time series: HDF5 in the reference's layout (solver.py:1214-1242), vector_0 = the initial state
    from common import load_h5_series
    d = load_h5_series(os.path.join(out, "results.h5"))
    assert d["potential"].shape[0] == 5 and d["concentrations"].shape[:2] == (5, 2)
    assert np.array_equal(d["potential"][-1].ravel(), S.phi.nodal().ravel())
    assert np.array_equal(d["concentrations"][-1][1].ravel(), S.c.split()[1].nodal().ravel())
    assert np.array_equal(d["elim_concentration"][-1].ravel(), S.ion_list[-1]["c"].nodal().ravel())
    assert np.array_equal(d["subdomains"], np.asarray(S._cell_tags)) and np.all(d["potential"][0] == 0.0)


def test_reference_tolerances_give_traces_within_solver_tolerance(emu_lib):
    S, O = sc.run_2d_neuron(emu_lib, 3)                # rtol 1e-5 / 1e-7 as run_2D.py:185-192
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 5e-4
    assert rel_err(S.c.split()[0].nodal(), O.c[0]) < 1e-5


def test_mms_space_rates_through_the_product_path(emu_lib):
    errs = np.array([sc.run_mms(emu_lib, r)[0] for r in (3, 4, 5)])
    rates = np.log(errs[:-1] / errs[1:]) / np.log(2.0)
    assert np.all(rates[-1] > 1.85) and np.all(rates[-1] < 2.2), rates


def test_mms_time_rates_through_the_product_path(emu_lib):
    """tests/run_MMS_time.py: first order in time (the reference prints the rates, expected ~1)"""
    errs = np.array([sc.run_mms_time(emu_lib, i)[0] for i in (2, 3, 4, 5)])
    rates = np.log(errs[:-1] / errs[1:]) / np.log(2.0)
    assert np.all(rates[-1] > 0.93) and np.all(rates[-1] < 1.07), rates
    assert np.all(np.diff(rates, axis=0) > 0)          # approaching 1 from below


def _bundle_run(lib, env, nsteps=6, mesh_fn=None):
    """a few steps of the small 3D bundle (or of mesh_fn()) under the given environment switches"""
    import os
    import bench
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh, mm_hh_no_stim
    from common import kmesh
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        mesh, sub, surf = mesh_fn() if mesh_fn else kmesh.bundle_3d_mesh(0)
        eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), lib=lib, **bench.PHYS)
        eng.set_concentrations_by_tag(bench.C_INIT)
        eng.add_membrane_model(1, mm_hh, bench.ION_NAMES, stimulus=bench.STIMULUS, stimulus_locator=bench.stim_locator)
        eng.add_membrane_model(2, mm_hh_no_stim, bench.ION_NAMES, stimulus=bench.STIMULUS,
                               stimulus_locator=bench.stim_locator)
        eng.rtol_emi, eng.rtol_knp = 1e-10, 1e-11
        for _ in range(nsteps):
            eng.step()
        return eng.phi_M().copy(), [eng.concentration(k).copy() for k in range(3)], dict(eng.stats)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_solver_engineering_switches_do_not_change_the_solution(emu_lib):
    """lagged preconditioner refresh, time-extrapolated initial guess, V(0,1) inside GMRES: all
    of them change HOW a system is solved, none may change WHAT a solve returns beyond the
    Krylov tolerance (the switches are read when a context is created)"""
    ref_pm, ref_c, ref_st = _bundle_run(emu_lib, {"KNP_AMG_REFRESH_PERIOD": "1", "KNP_EXTRAPOLATE": "0",
                                                  "KNP_KNP_PRESMOOTH": "1"})
    pm, c, st = _bundle_run(emu_lib, {"KNP_AMG_REFRESH_PERIOD": "8", "KNP_EXTRAPOLATE": "1",
                                      "KNP_KNP_PRESMOOTH": "0"})
    assert rel_err(pm, ref_pm) < 1e-7
    for k in range(3):
        assert rel_err(c[k], ref_c[k]) < 1e-8
    # and they must not cost iterations on this case
    assert sum(st["emi_niter"]) <= sum(ref_st["emi_niter"]) + 6
    # single-precision copies of the level-0 matrices inside the V-cycle (KNP_AMG_FP32=1; the Krylov
    # operators and every vector stay fp64): same answer, same iteration counts
    pm32, c32, st32 = _bundle_run(emu_lib, {"KNP_AMG_FP32": "1"})
    # degree-2 Chebyshev smoothing in the EMI V-cycle (KNP_AMG_CHEBY=2), together with the fp32 copies
    pmc, cc, stc = _bundle_run(emu_lib, {"KNP_AMG_CHEBY": "2", "KNP_AMG_FP32": "1"})
    assert rel_err(pmc, pm) < 1e-7
    for k in range(3):
        assert rel_err(cc[k], c[k]) < 1e-8
    assert sum(stc["emi_niter"]) <= sum(st["emi_niter"])
    assert rel_err(pm32, pm) < 1e-7
    for k in range(3):
        assert rel_err(c32[k], c[k]) < 1e-8
    assert sum(st32["emi_niter"]) <= sum(st["emi_niter"]) + 2
    assert sum(st32["knp_niter"]) <= sum(st["knp_niter"]) + 2


def test_smaller_preconditioner_shift_for_compact_cells(emu_lib):
    """KNP_EMI_LP_SCALE: compact cells in a large box (the EMIx geometry) - the mass shift kappa/Lp^2 of
    the reference's preconditioner matrix B outweighs the membrane coupling of each cell and costs CG
    iterations; a 100x smaller shift gives the same solution in fewer iterations"""
    from common import kmesh

    def mesh_fn():
        mesh, sub, surf = kmesh.emix_like_mesh(12, n_cells=6, length=4.0e-5)
        sub.array()[sub.array() == 2] = 1            # one intracellular tag (bench.C_INIT / D_sub know 0 and 1)
        return mesh, sub, surf
    pm, c, st = _bundle_run(emu_lib, {"KNP_EXTRAPOLATE": "0", "KNP_EMI_LP_SCALE": "1"}, nsteps=3, mesh_fn=mesh_fn)
    pm2, c2, st2 = _bundle_run(emu_lib, {"KNP_EXTRAPOLATE": "0", "KNP_EMI_LP_SCALE": "10"}, nsteps=3, mesh_fn=mesh_fn)
    assert rel_err(pm2, pm) < 1e-7
    for k in range(3):
        assert rel_err(c2[k], c[k]) < 1e-8
    assert sum(st2["emi_niter"]) < sum(st["emi_niter"]), (st["emi_niter"], st2["emi_niter"])
    # "auto": the scale that brings the shift of every region down to its membrane coupling ...
    pm3, c3, st3 = _bundle_run(emu_lib, {"KNP_EXTRAPOLATE": "0", "KNP_EMI_LP_SCALE": "auto"}, nsteps=3, mesh_fn=mesh_fn)
    assert rel_err(pm3, pm) < 1e-7
    assert sum(st3["emi_niter"]) < sum(st["emi_niter"])
    # ... and 1 (the reference's B, bit for bit the same run) for the long thin axons of the bench geometry
    a = _bundle_run(emu_lib, {"KNP_EMI_LP_SCALE": "auto"}, nsteps=2)
    b = _bundle_run(emu_lib, {"KNP_EMI_LP_SCALE": "1"}, nsteps=2)
    assert np.array_equal(a[0], b[0]) and a[2] == b[2]


def test_bench_emix_workload_builder(emu_lib, monkeypatch):
    """bench.py --emix M (BASELINE configs[4], the A/B workload for solver switches): glia stay at their
    calibrated rest, stimulated neurons depolarise, and the geometry-aware shift saves CG iterations"""
    import bench
    its = {}
    for mode in ("1", "auto"):
        monkeypatch.setenv("KNP_EMI_LP_SCALE", mode)
        eng = bench.build_engine_emix(12, 0, lib=emu_lib)
        for _ in range(3):
            eng.step()
        its[mode] = sum(eng.stats["emi_niter"])
        pm = eng.phi_M()
        assert abs(pm.min() + 83.085) < 0.05 and pm.max() > -60.0            # mV
        assert max(eng.stats["knp_niter"]) <= 8
    assert its["auto"] < its["1"], its


def test_emix_block_matches_oracle(emu_lib):
    """three cell tags, glial + neuronal membrane models side by side, ms/cm/mV units, synaptic stimulus:
    the engine against the oracle's time loop (north_star trace tolerance 1e-6)"""
    eng, O = sc.run_emix_block(emu_lib, 9, 2)
    assert eng.phi_M().max() > -60.0 and abs(eng.phi_M().min() + 83.085) < 0.01      # a neuron fires, glia rest
    assert rel_err(eng.phi_M(), O.phi_M) < 1e-6
    for k in range(2):
        assert rel_err(eng.concentration(k), O.c[k]) < 1e-9
    assert rel_err(eng.concentration(2), O.c_elim) < 1e-9


def test_picard_variant(emu_lib):
    sc.check_picard(emu_lib)


def test_picard_then_regular_steps_use_the_current_c_prev_n(emu_lib):
    sc.check_picard_then_regular(emu_lib)


def test_membrane_table_shapes_are_validated(emu_lib):
    sc.check_membrane_shape_validation(emu_lib)


def test_traces_over_an_action_potential_match_oracle(emu_lib):
    """north_star: "membrane-potential and concentration traces over a full run must agree within
    1e-6 relative".  100 steps (10 ms: stimulus, action potential, after-hyperpolarisation) on the
    resolved 2D neuron; both sides solve to tight Krylov tolerances, the ODE integrators differ
    (adaptive Dormand-Prince on the device, scipy LSODA in the oracle, both rtol 1e-8)."""
    S, O = sc.run_2d_neuron(emu_lib, 100, rtol_emi=1e-12, rtol_knp=1e-13, resolution=1)
    assert O.phi_M.max() < -0.07 and O.phi_M.min() < -0.08          # past the spike, in the after-hyperpolarisation
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 1e-6
    for k in range(2):
        assert rel_err(S.c.split()[k].nodal(), O.c[k]) < 1e-9
    assert rel_err(S.ion_list[-1]["c"].nodal(), O.c_elim) < 1e-9


def test_solver_emi_keeps_concentrations_frozen(emu_lib):
    sc.check_solver_emi(emu_lib)
