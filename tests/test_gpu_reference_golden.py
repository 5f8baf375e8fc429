"""CUDA library (C ABI) against the reference-EXECUTED golden fixtures (tests/golden/ref_*.npz):
assembled tensors entry by entry at 1e-12, one PDE step, and the 40-step membrane-potential
traces of the reference's own time loop at 1e-6."""
import os

import numpy as np
import pytest

import golden_checks as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gc.FORM_CASES)
def test_cuda_forms_match_the_reference(gpu_lib, name):
    assert gpu_lib.is_cuda()
    ctx = gc.check_library_forms(gpu_lib, name)
    assert ctx.launch_count() > 0


def test_cuda_run_matches_the_reference(gpu_lib):
    g = gc.run_golden()
    ref, n = g["phi_M_trace"], int(g["nsteps"])
    tr, eng = gc.library_run(gpu_lib, n, 1e-10, 1e-11)
    assert gc.trace_deviation(tr, ref) < 1e-6
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    assert gc.rel_err(cfin, g["final_c"]) < 1e-7


def test_cuda_astro_run_matches_the_reference(gpu_lib):
    """BASELINE configs[3] through the CUDA path: rho != 0, three membrane tags, glial + neuronal
    membranes, tortuosity, the time-windowed source - against the reference's own loop"""
    g = gc.astro_golden()
    tr, eng = gc.library_run_astro(gpu_lib, int(g["nsteps"]), int(g["M"]), 1e-10, 1e-11)
    assert gc.trace_deviation(tr, g["phi_M_trace"]) < 1e-6
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    assert gc.rel_err(cfin, g["final_c"]) < 1e-7
    assert gc.rel_err(eng.concentration(2).reshape(-1), g["final_c_elim"]) < 1e-7


def test_cuda_picard_run_matches_the_reference(gpu_lib):
    """the Picard variant (solver.py:850-927) on the CUDA path against the reference's own Picard loop"""
    g = np.load(os.path.join(gc.GOLDEN, "ref_run_2d_picard.npz"))
    tr, its, eng = gc.library_run_picard(gpu_lib, int(g["nsteps"]))
    assert max(its) <= 4
    assert gc.trace_deviation(tr, g["phi_M_trace"]) < 1e-6
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    assert gc.rel_err(cfin, g["final_c"]) < 1e-7


def test_cuda_solver_emi_matches_the_reference(gpu_lib):
    """SolverEMI on the CUDA path against the reference's own solver_emi.py (executed on oracle/refexec)"""
    import solver_checks as sc
    sc.check_solver_emi_against_reference(gpu_lib)


def test_cuda_passive_run_matches_the_reference(gpu_lib):
    """solve_system_passive (PDE steps only, non-splitting Robin forms) on the CUDA path"""
    import solver_checks as sc
    sc.check_passive_run_against_reference(gpu_lib)


def test_cuda_mms_matches_the_reference(gpu_lib):
    """BASELINE configs[0]: step-0 MMS-mode matrices entrywise, L2 errors of the space (r = 2..5) and time study 1e-6"""
    gc.check_library_mms(gpu_lib, resolutions=(2, 3, 4, 5))


def test_cuda_emix_run_matches_the_reference(gpu_lib):
    """BASELINE configs[4], the headline workload of bench.py (its own build_engine_emix, M = 9), through the CUDA path
    against the reference's own time loop on the problem of run_EMIx_simulation.py: 15 steps, one action potential"""
    gc.check_library_emix(gpu_lib)


def test_cuda_3d_bundle_run_matches_the_reference(gpu_lib):
    """BASELINE configs[2]: run_3D.py's four-axon bundle (mm_hh + mm_hh_no_stim) through the CUDA path against the
    reference's own time loop on the same mesh"""
    gc.check_library_3d(gpu_lib)
