#!/usr/bin/env python
"""bench.py - KNP-EMI DOF-steps/s (and time-steps/s) on the 3D axon-bundle workload
(BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling weak|strong]

One "step" = one full time step of the reference's loop (solver.py:1072-1127): membrane
ODE step -> EMI assembly + CG/AMG solve -> KNP assembly + GMRES/AMG solves -> post-step.
Workload: 32 x 0.9 x 0.9 um box with four axons (make_mesh_3D.py:81-111) at 96 x 27 x 27 x 6
tetrahedra = 419,904 cells, 5.04 M DOFs (3 fields x 4 dofs x cells), Hodgkin-Huxley membranes
with the synaptic stimulus of run_3D.py, dt = 0.1 ms, CG rtol 1e-5, GMRES(30) rtol 1e-7.

N > 1 (one process per GPU, torchrun): the mesh is partitioned by cell, one part per GPU,
DG halos over NCCL send/recv and Krylov dots over NCCL allreduce inside libknpemi.so.
Weak scaling (default): N copies of the four-axon block side by side (32 x 0.9 N x 0.9 um,
96 x 27 N x 27 x 6 tetrahedra, 4 N axons), so every GPU holds the N = 1 workload (5.04 M DOFs); `--scaling strong` keeps the N = 1 mesh.

Prints ONE JSON line (rank 0).  `value` = DOF-steps/s = (DOFs of the whole mesh) x steps /
time with everything resident in HBM (`steps_per_s` beside it); `e2e` = the same loop
driven with HOST buffers through the C ABI (state uploaded from pinned memory before and
downloaded after every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "knp-emi-dg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# spin-waiting exchange kernels and lazily loaded CUDA modules do not mix (csrc/knp_solve.cu,
# preload_solver_kernels): ask for eager loading before anything initialises CUDA
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("KNP_CONCURRENT_IONS") == "1":
    os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")   # (the opt-in cross-rank concurrency)

import numpy as np  # noqa: E402

WORKLOAD_DIMS = (96, 27, 27)
SAMPLE_DIMS = (32, 9, 9)          # CPU arms: bundle r=0 (make_mesh_3D.py resolution 0)
DT, C_M = 1.0e-4, 0.02
PHYS = dict(F=96485.0, R=8.314, T=300.0, C_M=C_M, C_phi=C_M / DT, dt=DT, z=[1.0, -1.0, 1.0],
            D_sub=[{0: 1.96e-9, 1: 1.96e-9}, {0: 2.03e-9, 1: 2.03e-9}, {0: 1.33e-9, 1: 1.33e-9}],
            rho_sub={0: 0.0, 1: 0.0})
NA_I, NA_E, K_I, K_E = 12.838513108648856, 100.71925900027354, 124.15397583491901, 3.3236967382705265
C_INIT = [{1: K_I, 0: K_E}, {1: NA_I + K_I, 0: NA_E + K_E}, {1: NA_I, 0: NA_E}]   # K, Cl, Na (run_3D.py)
ION_NAMES = ["K", "Cl", "Na"]
STIMULUS = {"stim_amplitude": 10.0}
# dram__bytes_read.sum + dram__bytes_write.sum of one BellSpmvKernel<4> launch on the N = 1 workload, from
# the `ncu --set full` capture summarised in profiles/kernels_r01_solver.md
TRAFFIC_SPMV = 3.119e8     # 299.4 MB read + 12.5 MB written (profiles/kernels_r01_solver.md, launch #1)


def stim_locator(x):
    return x[0] < 20.0e-6


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_engine(dims, device, transport=None, nblocks=1):
    from knpemidg import mesh as kmesh
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh, mm_hh_no_stim
    mesh, sub, surf = kmesh.bundle_3d_mesh(dims=dims, nblocks=nblocks)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), device=device, transport=transport,
                 **PHYS)
    eng.set_concentrations_by_tag(C_INIT)
    eng.add_membrane_model(1, mm_hh, ION_NAMES, stimulus=STIMULUS, stimulus_locator=stim_locator)
    eng.add_membrane_model(2, mm_hh_no_stim, ION_NAMES, stimulus=STIMULUS, stimulus_locator=stim_locator)
    eng.initialize(pc=1)
    return eng


# BASELINE configs[4]: EMIx-style tissue block, ms / cm / mV units, calibrated initial state
# (examples/emix-simulations/run_EMIx_simulation.py:56-147, 249); a second workload for A/B runs of the
# solver switches on compact cells (`--emix M`), not the headline line
EMIX_DT, EMIX_CM = 0.1, 2.0
EMIX_PHYS = dict(F=96485e3, R=8.314e3, T=300e3, C_M=EMIX_CM, C_phi=EMIX_CM / EMIX_DT, dt=EMIX_DT, z=[1.0, -1.0, 1.0],
                 D_sub=[{t: d for t in (0, 1, 2)} for d in (1.96e-8, 2.03e-8, 1.33e-8)],
                 rho_sub={0: 0.0, 1: 0.0, 2: 0.0})
_K = {0: 3.3236967382613933, 1: 102.75563828644862, 2: 124.15397583492471}       # ECS, glia, neuron
_NA = {0: 100.71925900028181, 1: 12.39731187972181, 2: 12.838513108606818}
EMIX_C_INIT = [_K, {t: _K[t] + _NA[t] for t in _K}, _NA]                           # K, Cl, Na


def build_engine_emix(M, device, transport=None, lib=None):
    from knpemidg import mesh as kmesh
    from knpemidg.engine import Engine
    from knpemidg.models import mm_glial_emix, mm_hh_emix
    mesh, sub, surf = kmesh.emix_like_mesh(M, n_cells=100, length=1.0e-3)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), device=device, transport=transport,
                 lib=lib, **EMIX_PHYS)
    eng.set_concentrations_by_tag(EMIX_C_INIT)
    stim = {"stim_amplitude": 5.0}
    locator = lambda x: x[0] < 3.0e-4                                              # noqa: E731
    eng.add_membrane_model(1, mm_glial_emix, ION_NAMES, stimulus=stim, stimulus_locator=locator)
    eng.add_membrane_model(2, mm_hh_emix, ION_NAMES, stimulus=stim, stimulus_locator=locator)
    eng.initialize(pc=1)
    return eng


def cpu_reference_steps(nsteps, dims=SAMPLE_DIMS):
    """The CPU restatement (oracle/) stepping the same kind of workload on a bounded
    sample; returns (seconds per step, dofs of the sample, timers)."""
    from knpemidg import mesh as kmesh
    from knpemidg.models import mm_hh, mm_hh_no_stim
    from oracle import forms, stepper
    mesh, sub, surf = kmesh.bundle_3d_mesh(dims=dims)
    P = forms.Problem(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), **PHYS)
    c0 = np.stack([np.where((sub.array() == 1)[:, None], ci[1], ci[0]) * np.ones((P.nc, P.nd)) for ci in C_INIT])
    O = stepper.OracleSolver(P, c0, models={1: mm_hh, 2: mm_hh_no_stim}, stimulus=STIMULUS,
                             stimulus_locator=stim_locator, ion_names=ION_NAMES, direct=False)
    O.step()                                   # warm-up (numpy/scipy first-call costs, AMG plan)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        O.step()
    dt = (time.perf_counter() - t0) / nsteps
    return dt, 3 * P.ndof, {"emi_niter": O.niter["emi"][-1:], "knp_niter": O.niter["knp"][-2:]}


def workload_dofs(dims, nblocks=1):
    return 3 * 4 * 6 * dims[0] * dims[1] * dims[2] * nblocks


def _cpu_worker(nsteps, q):
    try:
        try:                                    # one thread per copy: the copies fill the cores
            from threadpoolctl import threadpool_limits
            threadpool_limits(1)
        except Exception:
            pass
        q.put(cpu_reference_steps(nsteps))
    except Exception as e:  # pragma: no cover
        q.put(e)


def cpu_reference_parallel(nsteps, nproc):
    """`nproc` independent copies of the CPU restatement stepping the sample at the same time
    (what `mpirun -n nproc` of the reference could reach at best: perfect scaling, no
    communication).  Returns aggregate DOF-steps/s, seconds per step of the slowest copy, dofs."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(nsteps, q)) for _ in range(nproc)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    for r in res:
        if isinstance(r, Exception):
            raise r
    sec = max(r[0] for r in res)
    dofs = res[0][1]
    return nproc * dofs / sec, sec, dofs, res[0][2]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    """The reference's own path is dolfin + PETSc + numbalsoda, none of which exists in this
    image: the reference arm times the CPU restatement (oracle/) of the same
    time step on the host cores, one copy per core."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    cores = max(1, min(host_cores(), 32))
    value, sec, dofs, info = cpu_reference_parallel(steps, cores)
    sample = (f"{cores} independent copies (one per host core) of the oracle/ restatement on the bundle "
              f"{SAMPLE_DIMS[0]}x{SAMPLE_DIMS[1]}x{SAMPLE_DIMS[2]}x6 tets ({dofs} DOFs each), {steps} steps of "
              f"{sec:.2f} s after 1 warm-up step; DOF-steps/s summed over the copies")
    full = workload_dofs(WORKLOAD_DIMS)
    line = {"impl": "reference", "metric": "dof_steps_per_s", "value": value, "unit": "DOF-steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": 1e3 * full / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "3D axon bundle 96x27x27x6 tets, 5.04M DOFs, HH membranes (BASELINE configs[2])",
                       "note": "dolfin+PETSc+numbalsoda cannot be installed here; CPU restatement (oracle/) on "
                               "numpy/scipy on a bounded sample; ms_per_step = the full workload at this rate"},
            "steps_per_s": value / full,
            "cpu_baseline": {"value": value, "unit": "DOF-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "detail": info}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scaling", default="weak", choices=("weak", "strong"))
    ap.add_argument("--dims", default=None, help="nx,ny,nz override (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--emix", type=int, default=0, metavar="M",
                    help="second workload (BASELINE configs[4]): EMIx-like block of M^3 x 6 tets, ~100 cells; "
                         "M = 66 gives 20.7 M DOFs.  For A/B runs of solver switches; fixed mesh (strong scaling)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    dims = tuple(int(v) for v in args.dims.split(",")) if args.dims else WORKLOAD_DIMS
    warmup = max(args.warmup, 3)

    import torch
    dist = None
    transport = None
    if world > 1:
        import torch.distributed as dist
        from knpemidg.partition import TorchTransport
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        transport = TorchTransport()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def progress(msg):                 # stderr breadcrumbs: where a run was if it is ever killed by a timeout
        print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    nblocks = world if args.scaling == "weak" else 1
    progress("building the engine")
    if args.emix:
        args.scaling, args.no_cpu_baseline = "strong", True
        eng = build_engine_emix(args.emix, local_rank, transport)
    else:
        eng = build_engine(dims, local_rank, transport, nblocks)
    progress("engine ready")
    ctx = eng.ctx
    dofs = eng.dofs()                                  # of the whole (partitioned) mesh
    for _ in range(warmup):
        eng.step()
    progress("warm-up done")
    # ---- timed region: K steps, state resident in HBM -----------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.timers(reset=True)
    barrier()
    ctx.sync()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        eng.step()
    ms = ctx.timer_stop()
    ctx.sync()
    barrier()
    launches = ctx.launch_count() - l0
    progress("timed steps done")
    phase = ctx.timers()
    comm_info = ctx.dist_info()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    steps_per_s = args.steps / (ms * 1e-3)
    value = dofs * steps_per_s

    # ---- e2e: the same loop with host buffers through the C ABI -------------------
    from knpemidg import _lib
    n, nm, N = eng.n, eng.nm, eng.N
    host = {("c", k): torch.empty(n, dtype=torch.float64).pin_memory().numpy() for k in range(N)}
    host["phi"] = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    host["phiM"] = torch.empty(nm, dtype=torch.float64).pin_memory().numpy()

    def download():                       # device -> pinned host buffers, no staging copy
        for k in range(N):
            ctx.get_field(_lib.F_C, k, out=host[("c", k)])
        ctx.get_field(_lib.F_PHI, out=host["phi"])
        ctx.get_field(_lib.F_PHIM, out=host["phiM"])

    def upload():
        for k in range(N):
            ctx.set_field(_lib.F_C, k, host[("c", k)])
        ctx.set_field(_lib.F_PHI, 0, host["phi"])
        ctx.set_field(_lib.F_PHIM, 0, host["phiM"])

    download()
    e2e_steps = max(2, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(e2e_steps):
        upload()
        eng.step()
        download()
    ms_e2e = ctx.timer_stop()
    barrier()
    wall_e2e = time.perf_counter() - t0
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e * 1e3))
    e2e_value = dofs * e2e_steps / (ms_e2e * 1e-3)
    bytes_dir = sum_over_ranks(8.0 * ((N + 1) * n + nm))
    progress("e2e steps done")

    # ---- roofline of the hot kernels (CUDA events on the library's stream) ----------
    peak, peak_src = read_peaks()
    kern = {}
    for kid, name in ((0, "bell_spmv"), (3, "bell_block_jacobi_sweep"), (1, "emi_assembly"), (2, "knp_assembly")):
        kms, kbytes = ctx.bench_kernel(kid, reps=20)
        kern[name] = {"ms": kms, "algorithmic_bytes": kbytes, "gbs": kbytes / (kms * 1e-3) / 1e9,
                      "frac": kbytes / (kms * 1e-3) / 1e9 / peak}
    comm_us = None
    if world > 1:       # latency of one exchange of each kind, back to back (collective)
        comm_us = {}
        for kid, name in ((4, "dg_halo"), (5, "allreduce_4"), (6, "amg_tail_allgather")):
            barrier()
            kms, kbytes = ctx.bench_kernel(kid, reps=200)
            comm_us[name] = {"us": kms * 1e3, "bytes": kbytes}
    dom = kern["bell_spmv"]
    roofline = {"bound": "hbm", "kernel": "knp::BellSpmvKernel<4> (block-ELL fp64 SpMV)",
                "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": TRAFFIC_SPMV,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
                "ms_per_launch": dom["ms"], "other_kernels": kern}
    launches = int(sum_over_ranks(float(launches)))
    barrier()
    progress("kernel microbenchmarks done")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sec, sdofs, info = cpu_reference_steps(2)
        cpu = {"value": sdofs / sec, "unit": "DOF-steps/s", "cores": 1, "kind": "port",
               "sample": f"oracle/ restatement (numpy/scipy, one core), bundle {SAMPLE_DIMS} x6 tets ({sdofs} DOFs), "
                         f"2 steps of {sec:.2f} s after 1 warm-up step",
               "steps_per_s_on_workload": sdofs / sec / dofs}
    line = {"metric": "dof_steps_per_s", "value": value, "unit": "DOF-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D axon bundle {dims[0]}x{dims[1] * nblocks}x{dims[2]}x6 tets (BASELINE configs[2]"
                                   + (f", {nblocks} four-axon blocks side by side" if nblocks > 1 else "") + "), "
                                   f"{eng.nc_global} cells, {dofs} DOFs, HH membranes, dt=1e-4 s",
                       "parallelism": (f"cell partition, {world} parts (recursive bisection), NCCL halo + allreduce"
                                       if world > 1 else "single"),
                       "l2_policy": "inputs larger than L2 (matrices 3 x %.0f MB per GPU)" % (ctx.nnz * 8 / 1e6),
                       "solver": "CG rtol 1e-5 / GMRES(30) rtol 1e-7 (min 5 its), aggregation AMG: plan built once, values "
                                 "refreshed every %s solves or when iterations grow" % os.environ.get("KNP_AMG_REFRESH_PERIOD", "4")},
            "steps_per_s": steps_per_s,
            "seconds_per_step": {k: v / args.steps for k, v in phase.items()},
            "iterations": {"emi": eng.stats["emi_niter"][-args.steps:], "knp": eng.stats["knp_niter"][-args.steps:]},
            "e2e": {"value": e2e_value, "unit": "DOF-steps/s", "h2d_bytes_per_step": bytes_dir,
                    "d2h_bytes_per_step": bytes_dir, "steps": e2e_steps, "steps_per_s": e2e_steps / (ms_e2e * 1e-3)},
            "gpu_launches": launches, "comm": comm_info, "comm_latency": comm_us, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    if args.emix:
        line["config"]["workload"] = (f"EMIx-like tissue block {args.emix}^3 x 6 tets, ~100 cells (BASELINE configs[4]), "
                                      f"{eng.nc_global} cells, {dofs} DOFs, mm_hh + mm_glial membranes, dt=0.1 ms")
        line["roofline"]["traffic"] = None          # the ncu capture behind TRAFFIC_SPMV is of the bundle workload
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
