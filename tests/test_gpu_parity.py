"""GPU parity tests: the CUDA library (libknpemi.so, sm_100a) through its C ABI against
the CPU oracle on seeded inputs."""
import pytest

import parity_checks as pc

pytestmark = pytest.mark.gpu


def test_library_is_cuda_build(gpu_lib):
    assert gpu_lib.is_cuda()


@pytest.mark.parametrize("name,splitting,D_scale", [
    ("2d", True, (1.0, 1.0)), ("2d", False, (1.0, 0.5)), ("2d_r1", True, (1.0, 0.5)),
    ("emix", True, (1.0, 0.5)), ("emix", False, (1.0, 1.0)),
    ("3d_small", True, (1.0, 1.0)), ("3d_r0", True, (1.0, 0.5)),
])
def test_assembly_matches_oracle(gpu_lib, name, splitting, D_scale):
    pc.check_assembly(gpu_lib, name, splitting=splitting, D_scale=D_scale)


@pytest.mark.parametrize("name", ["2d", "emix", "3d_r0"])
def test_post_step_matches_oracle(gpu_lib, name):
    pc.check_post_step(gpu_lib, name)


def test_solvers_2d(gpu_lib):
    its = pc.check_solvers(gpu_lib, "2d", pcs=(0, 1))
    assert its[("emi", 1)] < its[("emi", 0)]


@pytest.mark.parametrize("name", ["emix", "3d_r0"])
def test_solvers_3d_amg(gpu_lib, name):
    pc.check_solvers(gpu_lib, name, pcs=(1,), max_emi_it=30)


def test_ode_models(gpu_lib):
    pc.check_ode(gpu_lib, ["mm_hh", "mm_hh_no_stim", "mm_leak", "mm_hh_emix", "mm_glial_emix",
                           "mm_hh_astro", "mm_glial_astro"], nsteps=4)


def test_ode_links_and_stimulus(gpu_lib):
    pc.check_ode_links(gpu_lib)
