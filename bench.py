#!/usr/bin/env python
"""bench.py - KNP-EMI DOF-steps/s (and time-steps/s) on the EMIx-style tissue block
(BASELINE.json configs[4], the workload north_star sets its targets on).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload emix|bundle|astro] [--size M] [--scaling strong|weak]

One "step" = one full time step of the reference's loop (solver.py:1072-1127): membrane
ODE step -> EMI assembly + CG/AMG solve -> KNP assembly + GMRES/AMG solves -> post-step.

Default workload (`emix`, M = 104): a 10 um block of 104^3 x 6 = 6,749,184 tetrahedra with ~100
compact cells (alternately glial, `mm_glial`, and neuronal, `mm_hh`; ms/cm/mV units, calibrated
initial state, run_EMIx_simulation.py:56-147, 249): 81.0 M DOFs (north_star: "at least 20 M";
`--size 66` is the 20.7 M-DOF block), dt = 0.1 ms, CG rtol 1e-5, GMRES(30) rtol 1e-7.  N > 1 (one process per GPU, torchrun): the SAME mesh partitioned by cell
(strong scaling), DG halos and Krylov dots over NVLink peer memory / NCCL inside libknpemi.so.
Other workloads: `bundle` = BASELINE configs[2] (96x27x27x6 tets, 5.04 M DOFs, HH membranes; weak
scaling = N four-axon blocks side by side), `astro` = configs[3] (three membrane tags, glial +
neuronal models, rho != 0, tortuosity, time-windowed source; run_tortuosity.py).

Prints ONE JSON line (rank 0).  `value` = DOF-steps/s = (DOFs of the whole mesh) x steps /
time with everything resident in HBM (`steps_per_s` beside it); `e2e` = the same loop
driven with HOST buffers through the C ABI (state uploaded from pinned memory before and
downloaded after every step).  `cpu_baseline` / `--impl reference`: the reference itself
(dolfin + PETSc + numbalsoda) cannot be installed here; the CPU arm is the C++/OpenMP port of
this library's own kernels (oracle/_port, `g++ -DKNP_EMU -fopenmp`) stepping the SAME workload
on all host cores.
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

import numpy as np  # noqa: E402

WORKLOAD_DIMS = (96, 27, 27)      # bundle
EMIX_M = 104                      # emix: 104^3 x 6 tets = 6.75 M cells = 81.0 M DOFs (north_star: >= 20 M; 66 -> 20.7 M)
ASTRO_M = 48                      # astro: 48^3 x 6 tets = 7.96 M DOFs
DT, C_M = 1.0e-4, 0.02
PHYS = dict(F=96485.0, R=8.314, T=300.0, C_M=C_M, C_phi=C_M / DT, dt=DT, z=[1.0, -1.0, 1.0],
            D_sub=[{0: 1.96e-9, 1: 1.96e-9}, {0: 2.03e-9, 1: 2.03e-9}, {0: 1.33e-9, 1: 1.33e-9}],
            rho_sub={0: 0.0, 1: 0.0})
NA_I, NA_E, K_I, K_E = 12.838513108648856, 100.71925900027354, 124.15397583491901, 3.3236967382705265
C_INIT = [{1: K_I, 0: K_E}, {1: NA_I + K_I, 0: NA_E + K_E}, {1: NA_I, 0: NA_E}]   # K, Cl, Na (run_3D.py)
ION_NAMES = ["K", "Cl", "Na"]
STIMULUS = {"stim_amplitude": 10.0}
# dram__bytes_read.sum + dram__bytes_write.sum of one BellSpmvKernel<4> launch on the N = 1 workloads, from
# the `ncu --set full` captures summarised in profiles/ (None: not captured for that workload)
TRAFFIC_SPMV = {("bundle", 0): 3.119e8,   # 299.4 MB read + 12.5 MB written (profiles/kernels_r01_solver.md, launch #1)
                ("emix", 66): 1.2352e9,   # 1181.3 MB read + 53.9 MB written (profiles/kernels_r02_emix.md, launch #0)
                ("emix", 104): 4.8861e9,  # 4669.9 MB read + 216.2 MB written (profiles/kernels_r02_emix104.md, launch #0)
                }


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


def build_engine(dims, device, transport=None, nblocks=1, lib=None):
    from knpemidg import mesh as kmesh
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh, mm_hh_no_stim
    mesh, sub, surf = kmesh.bundle_3d_mesh(dims=dims, nblocks=nblocks)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), device=device, transport=transport,
                 lib=lib, **PHYS)
    eng.set_concentrations_by_tag(C_INIT)
    eng.add_membrane_model(1, mm_hh, ION_NAMES, stimulus=STIMULUS, stimulus_locator=stim_locator)
    eng.add_membrane_model(2, mm_hh_no_stim, ION_NAMES, stimulus=STIMULUS, stimulus_locator=stim_locator)
    eng.initialize(pc=1)
    return eng


# BASELINE configs[4]: EMIx-style tissue block, ms / cm / mV units, calibrated initial state
# (examples/emix-simulations/run_EMIx_simulation.py:56-147, 249)
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


# BASELINE configs[3]: examples/local-astrocyte-depolarization/run_tortuosity.py:81-298 on the synthetic
# mesh knpemidg.mesh.astro_like_mesh (ions K, Na, Cl with Cl eliminated; rho != 0; tortuosity)
ASTRO = dict(dt=0.1, C_M=1.0, T=307e3, F=96500e3, R=8.315e3, g_syn=26.0, t_syn=1.2, lambda_i=3.2 * 4, lambda_e=1.6 * 4)
ASTRO_D = [1.96e-8, 1.33e-8, 2.03e-8]
ASTRO_C = [(3.092970607490389, 124.13988964240784, 99.3100014897692),
           (144.60625137617149, 12.850454639128186, 15.775818906083778),
           (133.62525154406637, 5.0, 5.203660274163705)]


def build_engine_astro(M, device, transport=None, lib=None):
    from knpemidg import mesh as kmesh
    from knpemidg.engine import Engine
    from knpemidg.models import mm_glial_astro, mm_hh_astro
    A = ASTRO
    lam = [A["lambda_e"], A["lambda_i"], A["lambda_i"]]
    mesh, sub, surf = kmesh.astro_like_mesh(M)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2, 3), device=device, transport=transport, lib=lib,
                 F=A["F"], R=A["R"], T=A["T"], C_M=A["C_M"], C_phi=A["C_M"] / A["dt"], dt=A["dt"], z=[1.0, 1.0, -1.0],
                 D_sub=[{t: D / lam[t] ** 2 for t in range(3)} for D in ASTRO_D],
                 rho_sub={t: -(ASTRO_C[1][t] + ASTRO_C[0][t] - ASTRO_C[2][t]) for t in range(3)})
    eng.set_concentrations_by_tag([{t: ci[t] for t in range(3)} for ci in ASTRO_C])
    names = ["K", "Na", "Cl"]
    links = (("K_e", 0, "plus"), ("Na_i", 1, "minus"))                              # run_tortuosity.py:38-49
    for tag, mod in ((1, mm_hh_astro), (2, mm_glial_astro), (3, mm_hh_astro)):
        eng.add_membrane_model(tag, mod, names, stimulus={"stim_amplitude": 0.0}, links=links)
    # the K+/Na+ source of run_tortuosity.py:180-200 inside its time window (constant on the source box)
    lo, hi = mesh.source_box
    mid = eng.mesh.coords[eng.mesh.cells].mean(axis=1)
    inside = np.all((mid >= lo) & (mid <= hi), axis=1) & (eng.cell_tags == 0)
    vol = eng.mesh.cell_volume()
    for k, sign in ((0, 1.0), (1, -1.0)):
        load = np.zeros((eng.nc, eng.nd))
        load[inside] = (sign * A["g_syn"] * vol[inside] / 4.0)[:, None]
        eng.ctx.set_field(_lib_mod().F_LOAD_KNP, k, load)
    eng.initialize(pc=1)
    return eng


def _lib_mod():
    from knpemidg import _lib
    return _lib


def make_engine(args, device, transport=None, lib=None, nblocks=1):
    if args.workload == "emix":
        return build_engine_emix(args.size or EMIX_M, device, transport, lib)
    if args.workload == "astro":
        return build_engine_astro(args.size or ASTRO_M, device, transport, lib)
    dims = tuple(int(v) for v in args.dims.split(",")) if args.dims else WORKLOAD_DIMS
    return build_engine(dims, device, transport, nblocks, lib)


def workload_name(args, eng, nblocks=1):
    dofs = eng.dofs()
    if args.workload == "emix":
        M = args.size or EMIX_M
        return (f"EMIx-like tissue block {M}^3 x 6 tets, ~100 cells, mm_hh + mm_glial membranes, calibrated initial "
                f"state, dt=0.1 ms (BASELINE configs[4]), {eng.nc_global} cells, {dofs} DOFs")
    if args.workload == "astro":
        M = args.size or ASTRO_M
        return (f"astrocyte depolarisation block {M}^3 x 6 tets, 2 neurons + 1 glial cell, 3 membrane tags, rho != 0, "
                f"tortuosity, K+/Na+ source on, dt=0.1 ms (BASELINE configs[3]), {eng.nc_global} cells, {dofs} DOFs")
    dims = tuple(int(v) for v in args.dims.split(",")) if args.dims else WORKLOAD_DIMS
    return (f"3D axon bundle {dims[0]}x{dims[1] * nblocks}x{dims[2]}x6 tets (BASELINE configs[2]"
            + (f", {nblocks} four-axon blocks side by side" if nblocks > 1 else "") + "), "
            f"{eng.nc_global} cells, {dofs} DOFs, HH membranes, dt=1e-4 s")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def omp_port_path():
    return os.path.join(ROOT, "oracle", "_port", "libknpemi_omp.so")


def run_reference(args, rank):
    """CPU arm.  The reference's own path is dolfin + PETSc + numbalsoda, none of which exists in
    this image: what is timed is the C++/OpenMP port of this library's kernels and host
    logic (oracle/_port/libknpemi_omp.so: same sources, g++ -DKNP_EMU -fopenmp) stepping the SAME
    workload - same mesh, membranes, tolerances - on all host cores."""
    if rank != 0:
        return
    cores = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "false")
    path = omp_port_path()
    if not os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("knp_build", os.path.join(ROOT, "knp-emi-dg_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_omp()
    from knpemidg import _lib
    lib = _lib.Lib(path)
    assert not lib.is_cuda()
    steps = max(1, min(args.steps, 20))          # bounded: ~3 s per step of the 81 M-DOF workload on 16 cores
    warmup = max(1, min(args.warmup, 3))
    t0 = time.perf_counter()
    eng = make_engine(args, 0, None, lib=lib)
    setup_s = time.perf_counter() - t0
    for _ in range(warmup):
        eng.step()
    eng.ctx.timers(reset=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.step()
    sec = (time.perf_counter() - t0) / steps
    phase = eng.ctx.timers()
    dofs = eng.dofs()
    value = dofs / sec
    sample = (f"C++/OpenMP port of the library's kernels (oracle/_port, g++ -O3 -fopenmp), the full workload, {steps} steps of "
              f"{sec:.2f} s after {warmup} warm-up step(s), {cores} threads; setup {setup_s:.1f} s not counted")
    line = {"impl": "reference", "metric": "dof_steps_per_s", "value": value, "unit": "DOF-steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * sec,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, eng),
                       "note": "dolfin+PETSc+numbalsoda cannot be installed here; the CPU arm is the C++/OpenMP port "
                               "of this library's kernels on the same workload, all host cores"},
            "steps_per_s": 1.0 / sec,
            "seconds_per_step": {k: v / steps for k, v in phase.items()},
            "iterations": {"emi": eng.stats["emi_niter"][-steps:], "knp": eng.stats["knp_niter"][-steps:]},
            "cpu_baseline": {"value": value, "unit": "DOF-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "DOF-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args):
    """the CPU arm on 2 steps of the same workload, in a process of its own (its OpenMP runtime and
    ~10 GB of host matrices stay out of the GPU process)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--workload", args.workload, "--scaling", args.scaling]
    if args.size:
        cmd += ["--size", str(args.size)]
    if args.dims:
        cmd += ["--dims", args.dims]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as e:  # the baseline is a reported number, not a reason to lose the GPU line
        return {"value": None, "unit": "DOF-steps/s", "cores": host_cores(), "kind": "port", "sample": f"failed: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="emix", choices=("emix", "bundle", "astro"))
    ap.add_argument("--size", type=int, default=0, metavar="M", help="emix / astro: M^3 x 6 tetrahedra (default 104 / 48)")
    ap.add_argument("--scaling", default=None, choices=("weak", "strong"),
                    help="default strong (the same mesh on every N); weak is available for the bundle")
    ap.add_argument("--dims", default=None, help="bundle: nx,ny,nz override")
    ap.add_argument("--emix", type=int, default=0, metavar="M", help="same as --workload emix --size M")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.emix:
        args.workload, args.size = "emix", args.emix
    if args.scaling is None:
        args.scaling = "strong"
    if args.scaling == "weak" and args.workload != "bundle":
        ap.error("weak scaling is defined for the bundle workload only")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
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
    eng = make_engine(args, local_rank, transport, nblocks=nblocks)
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
                "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": TRAFFIC_SPMV.get((args.workload, (args.size or EMIX_M) if args.workload == "emix" else args.size))
                if (world == 1 and not args.dims) else None,
                "peak_source": peak_src,
                "peak_note": "the peak is the driver's COPY bandwidth (half reads, half writes); this kernel is 96 % reads and a "
                             "plain read-only reduction reaches 6.9 TB/s on the same box (profiles/bandwidth_probe.py), so frac can "
                             "exceed 1",
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
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
        progress("CPU baseline (C++/OpenMP port, same workload, 2 steps)")
        cpu = cpu_baseline_subprocess(args)
    line = {"metric": "dof_steps_per_s", "value": value, "unit": "DOF-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, eng, nblocks),
                       "parallelism": (f"cell partition, {world} parts (recursive bisection), peer-memory / NCCL halo + allreduce"
                                       if world > 1 else "single"),
                       "l2_policy": "inputs larger than L2 (matrices 3 x %.0f MB per GPU)" % (ctx.nnz * 8 / 1e6),
                       "solver": "CG rtol 1e-5 / GMRES(30) rtol 1e-7 (min 5 its), aggregation AMG: plan built once, values "
                                 "refreshed every %s solves or when iterations grow" % os.environ.get("KNP_AMG_REFRESH_PERIOD", "8")},
            "steps_per_s": steps_per_s,
            "seconds_per_step": {k: v / args.steps for k, v in phase.items()},
            "iterations": {"emi": eng.stats["emi_niter"][-args.steps:], "knp": eng.stats["knp_niter"][-args.steps:]},
            "e2e": {"value": e2e_value, "unit": "DOF-steps/s", "h2d_bytes_per_step": bytes_dir,
                    "d2h_bytes_per_step": bytes_dir, "steps": e2e_steps, "steps_per_s": e2e_steps / (ms_e2e * 1e-3)},
            "gpu_launches": launches, "comm": comm_info, "comm_latency": comm_us, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
