"""Short driver for ncu captures: builds a bench workload (bench.py), steps it once and launches each
hot kernel a few times (knp_bench_kernel).

    python profiles/prof_kernels.py [emix|bundle|astro] [size] [reps]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "emix"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
args = argparse.Namespace(workload=workload, size=size, dims=None)
eng = bench.make_engine(args, 0)
eng.step()
for kid, name in ((0, "bell_spmv"), (3, "bell_block_jacobi"), (1, "emi_assembly"), (2, "knp_assembly")):
    ms, nbytes = eng.ctx.bench_kernel(kid, reps)
    print(f"{name}: {ms * 1e3:.1f} us/launch, {nbytes / 1e6:.1f} MB algorithmic, {nbytes / ms / 1e6:.0f} GB/s")
