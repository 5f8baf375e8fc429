"""CPU test-suite: the library sources compiled for the host (tests/emu, g++ -DKNP_EMU)
run the same per-thread kernel bodies and the same host logic (mesh tables, AMG plan,
Krylov loops, C ABI) as the CUDA build; checked here against the oracle.  The GPU
parity tests proper are tests/test_gpu_parity.py."""
import pytest

import parity_checks as pc


@pytest.mark.parametrize("name,splitting,D_scale", [
    ("2d", True, (1.0, 1.0)), ("2d", False, (1.0, 0.5)),
    ("emix", True, (1.0, 0.5)), ("3d_small", False, (1.0, 1.0)),
])
def test_assembly_matches_oracle(emu_lib, name, splitting, D_scale):
    pc.check_assembly(emu_lib, name, splitting=splitting, D_scale=D_scale)


@pytest.mark.parametrize("name", ["2d", "emix"])
def test_post_step_matches_oracle(emu_lib, name):
    pc.check_post_step(emu_lib, name)


def test_solvers_2d(emu_lib):
    its = pc.check_solvers(emu_lib, "2d", pcs=(0, 1))
    assert its[("emi", 1)] < its[("emi", 0)]


def test_solvers_3d_amg(emu_lib):
    pc.check_solvers(emu_lib, "emix", pcs=(1,), max_emi_it=25)


@pytest.mark.parametrize("name", ["unstr2d", "unstr3d"])
def test_unstructured_mesh_with_permuted_numbering(emu_lib, name):
    """Delaunay (2D) / jittered (3D) mesh with random vertex, cell and local-vertex numbering
    (tests/common.py:unstructured_mesh): assembly, post-step and both solvers against the oracle."""
    pc.check_assembly(emu_lib, name, splitting=True, D_scale=(1.0, 0.5))
    pc.check_post_step(emu_lib, name)
    pc.check_solvers(emu_lib, name, pcs=(1,))


def test_ode_models(emu_lib):
    pc.check_ode(emu_lib, ["mm_hh", "mm_glial_emix", "mm_leak"], nsteps=3)


def test_ode_links_and_stimulus(emu_lib):
    pc.check_ode_links(emu_lib)


def test_calibration_run_lands_on_the_reference_values(emu_lib):
    pc.check_calibration_kat(emu_lib)
