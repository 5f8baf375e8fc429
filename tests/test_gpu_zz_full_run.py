"""GPU: the complete run of examples/idealized-geometries/run_2D.py against the oracle (added after the
last GPU session of round 1; its CPU twin, tests/test_solver_api.py::
test_traces_over_an_action_potential_match_oracle, and a 200-step run on the host emulation - 6e-8 on the
membrane potential - pass).  Runs last so that a surprise here cannot hide the rest of the GPU suite."""
import pytest

import edge_checks as ec
import parity_checks as pc
import solver_checks as sc
from common import rel_err

pytestmark = pytest.mark.gpu


def test_full_run_traces_match_oracle(gpu_lib):
    """the run of examples/idealized-geometries/run_2D.py (Tstop = 20 ms, 200 steps, one action
    potential) on the resolved 2D neuron: final membrane potential and concentrations against the
    oracle within the north_star trace tolerance (1e-6 relative)"""
    S, O = sc.run_2d_neuron(gpu_lib, 200, rtol_emi=1e-12, rtol_knp=1e-13, resolution=1)
    assert abs(O.phi_M.mean() + 0.07515) < 2e-4                      # back at rest after the spike
    assert rel_err(S.phi_M_prev_PDE.vector().get_local(), O.phi_M) < 1e-6
    for k in range(2):
        assert rel_err(S.c.split()[k].nodal(), O.c[k]) < 1e-9
    assert rel_err(S.ion_list[-1]["c"].nodal(), O.c_elim) < 1e-9


@pytest.mark.parametrize("name", ["unstr2d", "unstr3d"])
def test_unstructured_mesh_with_permuted_numbering(gpu_lib, name):
    """GPU twin of tests/test_emu_parity.py::test_unstructured_mesh_with_permuted_numbering (added
    after the last GPU session of round 1, hence in this last-running file)"""
    pc.check_assembly(gpu_lib, name, splitting=True, D_scale=(1.0, 0.5))
    pc.check_post_step(gpu_lib, name)
    pc.check_solvers(gpu_lib, name, pcs=(1,))


def test_calibration_run_lands_on_the_reference_values(gpu_lib):
    """known answer from the reference's own calibration run (tests/parity_checks.py:check_calibration_kat);
    100 000 launches of the ODE kernel on one membrane point"""
    pc.check_calibration_kat(gpu_lib)


def test_emix_block_matches_oracle(gpu_lib):
    """GPU twin of tests/test_solver_api.py::test_emix_block_matches_oracle (bench.py --emix workload at
    test size: glial + neuronal membranes, ms/cm/mV units), four steps"""
    eng, O = sc.run_emix_block(gpu_lib, 9, 4)
    assert eng.phi_M().max() > -60.0
    assert rel_err(eng.phi_M(), O.phi_M) < 1e-6
    for k in range(2):
        assert rel_err(eng.concentration(k), O.c[k]) < 1e-9
    assert rel_err(eng.concentration(2), O.c_elim) < 1e-9


def test_membrane_model_without_facets(gpu_lib):
    """GPU twin of tests/test_edge_cases.py::test_membrane_model_without_facets (empty ODE tables)"""
    ec.check_model_without_facets(gpu_lib)
