"""GPU, 2+ devices: the partitioned path over NCCL (one process per GPU) against a single-GPU
run of the same library.  Skipped on a one-GPU box; the host logic of the same path is
covered on CPU by tests/test_dist_gloo.py."""
import pytest

from test_dist_gloo import launch

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,case", [(2, "bundle_r0"), (2, "bundle_r0_quad"), (2, "unstr3d")])
def test_nccl_partitioned_run_matches_single_gpu(gpu_lib, tmp_path, world, case):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    res = launch(world, "gpu", case, 3, str(tmp_path / "out.json"), timeout=900)
    for key, err in res["errs"].items():
        assert err < 1e-8, (key, err, res)
    assert res["dist"]["world"] == world and res["dist"]["halos"] > 0
