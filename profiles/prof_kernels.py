"""Short driver for ncu captures: builds the bench workload (bench.py, 3D axon bundle),
steps it once and launches each hot kernel a few times (knp_bench_kernel).

    python profiles/prof_kernels.py [nx,ny,nz] [reps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dims = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else bench.WORKLOAD_DIMS
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = bench.build_engine(dims, 0)
eng.step()
for kid, name in ((0, "bell_spmv"), (3, "bell_block_jacobi"), (1, "emi_assembly"), (2, "knp_assembly")):
    ms, nbytes = eng.ctx.bench_kernel(kid, reps)
    print(f"{name}: {ms * 1e3:.1f} us/launch, {nbytes / 1e6:.1f} MB algorithmic, {nbytes / ms / 1e6:.0f} GB/s")
