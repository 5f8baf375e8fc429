"""bench.py's CPU arm (`--impl reference`: the C++/OpenMP port of the library's kernels on the same
workload) runs without a GPU; its JSON line carries the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", *extra], capture_output=True, text=True, timeout=900, env=env, check=True)
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line_on_a_small_block():
    line = _run("--size", "12")
    assert line["impl"] == "reference" and line["metric"] == "dof_steps_per_s" and line["unit"] == "DOF-steps/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["scaling"] == "strong"
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "12^3 x 6 tets" in line["config"]["workload"] and "configs[4]" in line["config"]["workload"]
    assert line["iterations"]["knp"][0] >= 5                     # ksp_min_it (solver.py:686)


def test_reference_arm_runs_the_astrocyte_workload():
    line = _run("--workload", "astro", "--size", "8")
    assert "configs[3]" in line["config"]["workload"] and line["value"] > 0
