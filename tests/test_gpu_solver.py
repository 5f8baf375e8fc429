"""GPU: the reference-facing API (Solver / MembraneModel / utils) on the CUDA library:
run_2D.py flow against the oracle, the MMS study of tests/run_MMS_space.py, and
size-independent properties at the bench workload's scale."""
import numpy as np
import pytest

import solver_checks as sc
from common import rel_err

pytestmark = pytest.mark.gpu


def test_run_2d_flow_matches_oracle(gpu_lib, tmp_path):
    S, O = sc.run_2d_neuron(gpu_lib, 6, rtol_emi=1e-12, rtol_knp=1e-13, outdir=str(tmp_path) + "/")
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 1e-6      # north_star trace tolerance
    for k in range(2):
        assert rel_err(S.c.split()[k].nodal(), O.c[k]) < 1e-6
    assert rel_err(S.ion_list[-1]["c"].nodal(), O.c_elim) < 1e-6


def test_run_2d_flow_at_the_reference_resolution(gpu_lib):
    """run_2D.py:58 runs resolution 2 (3 968 cells, 248 ODE points): 10 steps against the oracle loop"""
    S, O = sc.run_2d_neuron(gpu_lib, 10, rtol_emi=1e-12, rtol_knp=1e-13, resolution=2)
    assert S.mem_models[0]["ode"].nodes == 248
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 1e-6
    for k in range(2):
        assert rel_err(S.c.split()[k].nodal(), O.c[k]) < 1e-8


def test_reference_tolerances(gpu_lib):
    S, O = sc.run_2d_neuron(gpu_lib, 3)
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 5e-4
    assert rel_err(S.c.split()[0].nodal(), O.c[0]) < 1e-5


def test_mms_space_rates(gpu_lib):
    errs = np.array([sc.run_mms(gpu_lib, r)[0] for r in (3, 4, 5, 6)])
    rates = np.log(errs[:-1] / errs[1:]) / np.log(2.0)
    assert np.all(rates[-1] > 1.9) and np.all(rates[-1] < 2.1), rates


def test_mms_time_rates(gpu_lib):
    errs = np.array([sc.run_mms_time(gpu_lib, i, r=4)[0] for i in (2, 3, 4, 5)])
    rates = np.log(errs[:-1] / errs[1:]) / np.log(2.0)
    assert np.all(rates[-1] > 0.93) and np.all(rates[-1] < 1.07), rates


def test_rest_state_is_preserved_at_scale(gpu_lib):
    """with no stimulus the coupled system stays at rest (phi_M = -74.386 mV,
    concentrations constant); also electroneutrality of the eliminated ion.  Bundle mesh
    r=0 (186,624 DOFs), 5 steps."""
    import bench
    from knpemidg import _lib
    from knpemidg.engine import Engine
    from knpemidg import mesh as kmesh
    from knpemidg.models import mm_hh_no_stim
    mesh, sub, surf = kmesh.bundle_3d_mesh(0)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), lib=gpu_lib, **bench.PHYS)
    eng.set_concentrations_by_tag(bench.C_INIT)
    for tag in (1, 2):
        eng.add_membrane_model(tag, mm_hh_no_stim, bench.ION_NAMES)
    c0 = [eng.concentration(k).copy() for k in range(3)]
    for _ in range(5):
        eng.step()
    pm = eng.phi_M()
    # 5 steps at the reference's KSP tolerances (CG rtol 1e-5): within 0.01 mV of rest
    assert np.abs(pm + 0.07438609374462003).max() < 1e-5
    # no pumps in the HH model: the resting Na+/K+ leak currents move the concentrations next
    # to the membranes by a few 1e-5 (relative) over 5 steps; anything larger is a solver fault
    for k in range(3):
        assert rel_err(eng.concentration(k), c0[k]) < 1e-4
    z = bench.PHYS["z"]
    total = sum(z[k] * eng.concentration(k) for k in range(3))
    assert np.abs(total).max() < 1e-9 * np.abs(c0[1]).max()


def test_picard_variant(gpu_lib):
    """solve_for_time_step_picard (solver.py:850-927) on the CUDA path"""
    sc.check_picard(gpu_lib)


def test_picard_then_regular_steps_use_the_current_c_prev_n(gpu_lib):
    sc.check_picard_then_regular(gpu_lib)


def test_solver_emi_keeps_concentrations_frozen(gpu_lib):
    """SolverEMI (solver_emi.py:52-822) on the CUDA path"""
    sc.check_solver_emi(gpu_lib)


def test_membrane_table_shapes_are_validated(gpu_lib):
    sc.check_membrane_shape_validation(gpu_lib)
