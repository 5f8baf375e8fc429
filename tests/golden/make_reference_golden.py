#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz by EXECUTING the reference (adajel/KNP-EMI-DG, /root/reference,
unmodified) on top of the numeric dolfin / petsc4py / numbalsoda stand-ins of oracle/refexec.

    python tests/golden/make_reference_golden.py [outdir] [--only=name,name]      (this container only)

Every fixture holds the mesh, the parameters and the seeded input fields of one case together
with what the reference's own code produced from them:

  ref_forms_<case>.npz   assemble(S.a_emi), assemble(S.L_emi), assemble(S.B_emi), assemble(S.A_knp),
                         assemble(S.L_knp) after S.setup_varform_emi() / S.setup_varform_knp()
                         (src/knpemidg/solver.py:270-403, 534-663) and the state after one
                         S.solve_for_time_step() (solver.py:794-847: phi, c, phi_M, Nernst potentials,
                         eliminated ion) - cases 2d (splitting), 2d_nosplit, 3d (three cell tags,
                         two membrane tags, region-dependent D, rho != 0, a source term);
  ref_run_2d.npz         S.solve_system_active() (solver.py:1014-1135) for 40 steps of the 2D neuron
                         of examples/idealized-geometries/run_2D.py on its resolution-1 mesh (resolution 0 does not resolve the membrane) with the
                         reference's mm_hh.py: membrane potential at every membrane facet and step,
                         final fields;
  ref_run_2d_picard.npz  the same problem, 12 steps, with the reference's Picard variant solve_for_time_step_picard
                         (solver.py:850-927) in place of the split PDE step;
  ref_run_2d_emi.npz     the same problem, 25 steps, with the reference's EMI-only SolverEMI (solver_emi.py:52-822);
  ref_run_2d_passive.npz S.solve_system_passive() (solver.py:930-1011: PDE steps only, non-splitting forms), 10 steps;
  ref_mms.npz            the reference's manufactured-solution study, BASELINE configs[0]: its own setup_mms
                         (tests/mms_space.py, tests/mms_time.py), Solver(mms=...) and solve_system_passive as driven by
                         tests/run_MMS_space.py / run_MMS_time.py - step-0 tensors and fields of the resolution-2 space case
                         and of one time case, the L2 errors the scripts print for resolutions 2..5 and dt_0/4..dt_0/16;
  ref_run_3d.npz         S.solve_system_active() for 8 steps of the four-axon bundle of examples/idealized-geometries/run_3D.py
                         (BASELINE configs[2]) on its own resolution-0 mesh, mm_hh + mm_hh_no_stim;
  ref_run_emix.npz       S.solve_system_active() for 15 steps of the problem of examples/emix-simulations/run_EMIx_simulation.py
                         (BASELINE configs[4], the workload of bench.py's headline number: glial + neuronal membrane models,
                         ms / cm / mV units, synaptic stimulus) on the synthetic block knpemidg.mesh.emix_like_mesh(9);
  ref_run_astro.npz      S.solve_system_active() for 16 steps of the problem of
                         examples/local-astrocyte-depolarization/run_tortuosity.py (BASELINE configs[3]: three
                         membrane tags, neuronal + glial models, rho != 0, tortuosity, the time-windowed K+/Na+
                         source) on the synthetic mesh knpemidg.mesh.astro_like_mesh(8).

dof numbering of the stored tensors: DG1 dof = nd*cell + local vertex; the mixed KNP space
stacks the ions (ion k at offset k*nd*ncells); facet quantities by facet index.
"""
import importlib.util
import os
import sys
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refexec  # noqa: E402

ref = refexec.install()
import dolfin as df  # noqa: E402  (the stand-in)
from knpemidg import Solver, SolverEMI  # noqa: E402  (the reference)
from knpemidg.utils import plus, minus, pcws_constant_project  # noqa: E402


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


kmesh = load_by_path("_kmesh_for_golden", os.path.join(ROOT, "knp-emi-dg_b200", "knpemidg", "mesh.py"))
REF_EX = "/root/reference/examples"
sys.path.insert(0, os.path.join(REF_EX, "idealized-geometries"))
mm_hh = load_by_path("mm_hh", os.path.join(REF_EX, "idealized-geometries", "mm_hh.py"))
mm_hh_no_stim = load_by_path("mm_hh_no_stim", os.path.join(REF_EX, "idealized-geometries", "mm_hh_no_stim.py"))


class RefSolver(Solver):
    """the sub-class of examples/idealized-geometries/run_2D.py:30-50"""

    def __init__(self, params, ion_list, degree_emi=1, degree_knp=1, mms=None, sf=1):
        Solver.__init__(self, params, ion_list, degree_emi=1, degree_knp=1, mms=None, sf=1)
        self.trace = []

    def update_ode(self, ode_model):
        K_e = plus(self.c_prev_k.split()[0], self.n_g)
        ode_model.set_parameter('K_e', pcws_constant_project(K_e, self.Q))
        Na_i = minus(self.ion_list[-1]['c'], self.n_g)
        ode_model.set_parameter('Na_i', pcws_constant_project(Na_i, self.Q))

    def solve_for_time_step(self, k, t):
        Solver.solve_for_time_step(self, k, t)
        self.trace.append(self.phi_M_prev_PDE.vector().get_local().copy())


Params = namedtuple('params', ('dt', 'n_steps_ODE', 'F', 'psi', 'phi_M_init', 'C_phi', 'C_M', 'R', 'temperature',
                               'phi_M_init_type', 'rho_sub'))
SolverParams = namedtuple('solver_params', ('direct_emi', 'direct_knp', 'resolution', 'rtol_emi', 'rtol_knp',
                                            'atol_emi', 'atol_knp', 'threshold_emi', 'threshold_knp'))
StimParams = namedtuple('membrane_params', ('g_syn_bar', 'stimulus', 'stimulus_locator'))

PHYS = dict(dt=1.0e-4, C_M=0.02, T=300.0, F=96485.0, R=8.314)
D_PHYS = {"K": 1.96e-9, "Cl": 2.03e-9, "Na": 1.33e-9}
NA_I, NA_E, K_I, K_E = 12.838513108648856, 100.71925900027354, 124.15397583491901, 3.3236967382705265


def build_solver(mesh, sub, surf, tags, ode_models, D_scale=None, rho=None, f_source=None, cls=None):
    """reference Solver set up as run_2D.py / run_3D.py do (ions K, Cl, Na; Na eliminated)"""
    dt, C_M = PHYS["dt"], PHYS["C_M"]
    rho_sub = {int(t): df.Constant((rho or {}).get(int(t), 0.0)) for t in tags}
    params = Params(dt, 25, PHYS["F"], PHYS["F"] / (PHYS["R"] * PHYS["T"]), df.Constant(-0.0743860937), C_M / dt, C_M,
                    PHYS["R"], PHYS["T"], 'constant', rho_sub)
    ci = {"K": K_I, "Na": NA_I, "Cl": K_I + NA_I}
    ce = {"K": K_E, "Na": NA_E, "Cl": K_E + NA_E}
    ions = []
    for name, z in (("K", 1.0), ("Cl", -1.0), ("Na", 1.0)):
        scale = D_scale or {}
        ions.append({'c_init_sub': {int(t): df.Constant(ce[name] if t == 0 else ci[name]) for t in tags},
                     'c_init_sub_type': 'constant', 'bdry': None, 'z': z, 'name': name,
                     'D_sub': {int(t): df.Constant(D_PHYS[name] * scale.get(int(t), 1.0)) for t in tags},
                     'f_source': (f_source or {}).get(name, df.Constant(0))})
    S = (cls or RefSolver)(params, ions)
    dmesh = df.Mesh(mesh)
    S.setup_domain(dmesh, df.MeshFunction.from_array(dmesh, mesh.gdim, sub), df.MeshFunction.from_array(dmesh, mesh.gdim - 1, surf))
    S.setup_parameters()
    S.setup_FEM_spaces()
    stim = StimParams(10, {'stim_amplitude': 10}, lambda x: x[0] < 20e-6)
    S.setup_membrane_model(stim, ode_models)
    return S, ions


def coo(M):
    A = M.A.tocoo()
    A.sum_duplicates()
    return A.row.astype(np.int32), A.col.astype(np.int32), A.data


def forms_case(name, mesh, sub, surf, ode_models, splitting=True, D_scale=None, rho=None, f_src=None, seed=0):
    mesh.init_topology()
    tags = np.unique(sub)
    fs = None
    if f_src is not None:
        fs = {"K": df.Constant(f_src[0]), "Cl": df.Constant(f_src[1])}
    S, ions = build_solver(mesh, sub, surf, tags, ode_models, D_scale, rho, fs)
    nc, nd = mesh.num_cells(), mesh.nd
    n = nc * nd
    rng = np.random.default_rng(seed)
    ics = (sub != 0)[:, None]
    base_i = [K_I, K_I + NA_I, NA_I]
    base_e = [K_E, K_E + NA_E, NA_E]
    c_all = np.stack([np.where(ics, base_i[k], base_e[k]) * (1.0 + 0.01 * rng.uniform(-1, 1, (nc, nd))) for k in range(3)])
    c_n = c_all[:2] * (1.0 + 0.001 * rng.uniform(-1, 1, (2, nc, nd)))
    phi = np.where(ics, -0.07, 0.0) * (1.0 + 0.01 * rng.uniform(-1, 1, (nc, nd)))
    nf = mesh.facet_cells.shape[0]
    phi_M = -0.07 * (1.0 + 0.01 * rng.uniform(-1, 1, nf))
    I_ch = 1e-3 * rng.uniform(-1, 1, (3, nf))
    # the seeded state goes into the reference's own Functions
    S.c_prev_k.vector().set_local(c_all[:2].reshape(-1))
    S.c_prev_n.vector().set_local(c_n.reshape(-1))
    S.c.vector().set_local(c_all[:2].reshape(-1))
    ions[-1]['c'].vector().set_local(c_all[2].reshape(-1))
    S.phi.vector().set_local(phi.reshape(-1))
    S.phi_M_prev_PDE.vector().set_local(phi_M)
    for mm in S.mem_models:
        for k, nm_ in enumerate(("K", "Cl", "Na")):
            mm['I_ch_k'][nm_].vector().set_local(I_ch[k])
    S.splitting_scheme = splitting
    S.setup_varform_emi()
    S.setup_varform_knp()
    A, B, b = df.assemble(S.a_emi), df.assemble(S.B_emi), df.assemble(S.L_emi)
    Ak, bk = df.assemble(S.A_knp), df.assemble(S.L_knp)
    out = dict(coords=mesh.coords, cells=mesh.cells, cell_tag=sub.astype(np.int32), facet_tag=surf.astype(np.int32),
               membrane_tags=np.array(sorted(ode_models), dtype=np.int32), splitting=np.array(int(splitting)),
               F=PHYS["F"], R=PHYS["R"], T=PHYS["T"], C_M=PHYS["C_M"], dt=PHYS["dt"], z=np.array([1.0, -1.0, 1.0]),
               tags=tags.astype(np.int32),
               D=np.array([[D_PHYS[nm_] * (D_scale or {}).get(int(t), 1.0) for t in tags] for nm_ in ("K", "Cl", "Na")]),
               rho=np.array([(rho or {}).get(int(t), 0.0) for t in tags]),
               f_source=np.array(f_src if f_src is not None else [0.0, 0.0]),
               c_all=c_all, c_n=c_n, phi=phi, phi_M=phi_M, I_ch=I_ch,
               E0=np.stack([ion['E'].vector().get_local() for ion in ions]),   # Nernst potentials of the input state
               n_g=S.n_g.vector().get_local().reshape(nf, mesh.gdim),
               b_emi=b.get_local(), b_knp=bk.get_local())
    for key, M in (("A_emi", A), ("B_emi", B), ("A_knp", Ak)):
        r, c, v = coo(M)
        out[key + "_row"], out[key + "_col"], out[key + "_val"] = r, c, v
    # one PDE step of the reference (direct solves), solver.py:794-847
    S.direct_emi = S.direct_knp = True
    S.save_solver_stats = False
    S.setup_solver_emi()
    S.setup_solver_knp()
    S.solve_for_time_step(0, df.Constant(0.0))
    out.update(step_phi=S.phi.vector().get_local(), step_c=S.c.vector().get_local(),
               step_phi_M=S.phi_M_prev_PDE.vector().get_local(),
               step_E=np.stack([ion['E'].vector().get_local() for ion in ions]),
               step_c_elim=ions[-1]['c'].vector().get_local())
    return out


def mesh_3d_two_cells():
    """6 x 4 x 4 boxes (x6 tets), a 'glial' block (cell tag 1, membrane tag 1) and a 'neuronal' block
    (cell tag 2, membrane tag 2) that do not touch; exterior facets 5; scaled to micrometres"""
    m = kmesh.box_mesh((0.0, 0.0, 0.0), (6.0, 4.0, 4.0), 6, 4, 4)
    m.init_topology()
    mid = m.cell_midpoints()
    sub = np.zeros(m.num_cells(), dtype=np.int64)
    inside = lambda lo, hi: np.all((mid > lo) & (mid < hi), axis=1)   # noqa: E731
    sub[inside(np.array([1.0, 1.0, 1.0]), np.array([2.0, 3.0, 3.0]))] = 1
    sub[inside(np.array([3.0, 1.0, 1.0]), np.array([5.0, 3.0, 2.0]))] = 2
    fc = m.facet_cells
    interior = fc[:, 1] >= 0
    t0 = sub[fc[:, 0]]
    t1 = np.where(interior, sub[np.maximum(fc[:, 1], 0)], t0)
    surf = np.zeros(fc.shape[0], dtype=np.int64)
    surf[interior & (t0 != t1)] = np.maximum(t0, t1)[interior & (t0 != t1)]
    surf[~interior] = 5
    m2 = kmesh.SimplexMesh(m.coords * 1e-6, m.cells)
    m2.init_topology()
    return m2, sub, surf


class RefSolverPicard(RefSolver):
    """the Picard variant the reference keeps beside the split step (solver.py:850-927; its call is commented
    out at :1123): every PDE step of the loop goes through solve_for_time_step_picard"""

    def solve_for_time_step(self, k, t):
        Solver.solve_for_time_step_picard(self, k, t)
        self.trace.append(self.phi_M_prev_PDE.vector().get_local().copy())


class RefSolverEMI(SolverEMI):
    """the reference's EMI-only solver (src/knpemidg/solver_emi.py:52-822), concentrations frozen"""

    def __init__(self, params, ion_list, degree_emi=1, degree_knp=1, mms=None, sf=1):
        SolverEMI.__init__(self, params, ion_list)
        self.trace = []

    def solve_for_time_step(self, k, t):
        SolverEMI.solve_for_time_step(self, k, t)
        self.trace.append(self.phi_M_prev_PDE.vector().get_local().copy())


def run_emi_case(nsteps=25):
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    S, ions = build_solver(mesh, sub, surf, np.unique(sub), {1: mm_hh}, cls=RefSolverEMI)
    t = df.Constant(0.0)
    S.solve_system_active(nsteps * PHYS["dt"], t, SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None))
    mem = np.flatnonzero(surf == 1)
    return dict(nsteps=np.array(nsteps), mem_facets=mem.astype(np.int32), phi_M_trace=np.stack(S.trace)[:, mem],
                final_phi=S.phi.vector().get_local(), t_end=np.array(float(t)))


def run_passive_case(nsteps=10):
    """solve_system_passive (solver.py:930-1011): PDE steps only, the non-splitting Robin forms (:337, 621-622),
    membrane currents as the ODE tables hold them initially; perturbed initial membrane potential so that
    something relaxes"""
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    S, ions = build_solver(mesh, sub, surf, np.unique(sub), {1: mm_hh})
    nf = mesh.facet_cells.shape[0]
    phi_M0 = -0.07 * (1.0 + 0.05 * np.sin(np.arange(nf)))
    S.phi_M_prev_PDE.vector().set_local(phi_M0)
    t = df.Constant(0.0)
    uh, c_elim = S.solve_system_passive(nsteps * PHYS["dt"], t, SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None), None)
    mem = np.flatnonzero(surf == 1)
    return dict(nsteps=np.array(nsteps), mem_facets=mem.astype(np.int32), phi_M0=phi_M0[mem],
                phi_M_trace=np.stack(S.trace)[:, mem], final_phi=uh[-1].vector().get_local(),
                final_c=S.c.vector().get_local(), final_c_elim=c_elim.vector().get_local())


def run_case(nsteps=40, picard=False):
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    S, ions = build_solver(mesh, sub, surf, np.unique(sub), {1: mm_hh}, cls=RefSolverPicard if picard else RefSolver)
    sp = SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 1e-40, None, None)
    t = df.Constant(0.0)
    S.solve_system_active(nsteps * PHYS["dt"], t, sp)
    mem = np.flatnonzero(surf == 1)
    return dict(nsteps=np.array(nsteps), mem_facets=mem.astype(np.int32), phi_M_trace=np.stack(S.trace)[:, mem],
                final_phi=S.phi.vector().get_local(), final_c=S.c.vector().get_local(),
                final_c_elim=ions[-1]['c'].vector().get_local(),
                final_E=np.stack([ion['E'].vector().get_local()[mem] for ion in ions]),
                final_states=S.mem_models[0]['ode'].states.copy(), t_end=np.array(float(t)))


def run_3d_case(nsteps=8):
    """BASELINE configs[2]: the problem of examples/idealized-geometries/run_3D.py (four axons; mm_hh on the membrane of
    the first, mm_hh_no_stim on the other three, run_3D.py:196) on its own resolution-0 mesh (make_mesh_3D.py:81-111:
    15 552 tetrahedra, 1 472 membrane facets)"""
    mesh, sub, surf = kmesh.bundle_3d_mesh(0)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    S, ions = build_solver(mesh, sub, surf, np.unique(sub), {1: mm_hh, 2: mm_hh_no_stim})
    t = df.Constant(0.0)
    S.solve_system_active(nsteps * PHYS["dt"], t, SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 2e-40, 0.9, 0.75))
    mem = np.flatnonzero(np.isin(surf, (1, 2)))
    return dict(nsteps=np.array(nsteps), mem_facets=mem.astype(np.int32), mem_tag=surf[mem].astype(np.int32),
                phi_M_trace=np.stack(S.trace)[:, mem], final_phi=S.phi.vector().get_local(),
                final_c=S.c.vector().get_local(), final_c_elim=ions[-1]['c'].vector().get_local(),
                final_E=np.stack([ion['E'].vector().get_local()[mem] for ion in ions]), t_end=np.array(float(t)))


# ---- BASELINE configs[3]: examples/local-astrocyte-depolarization/run_tortuosity.py ------------------------
ASTRO = dict(dt=0.1, C_M=1.0, T=307e3, F=96500e3, R=8.315e3, g_syn=26.0, t_syn=1.2, lambda_i=3.2 * 4, lambda_e=1.6 * 4,
             D={"K": 1.96e-8, "Na": 1.33e-8, "Cl": 2.03e-8},
             c={"K": (3.092970607490389, 124.13988964240784, 99.3100014897692),       # ECS, neuron, glia
                "Na": (144.60625137617149, 12.850454639128186, 15.775818906083778),
                "Cl": (133.62525154406637, 5.0, 5.203660274163705)})


class RefSolverAstro(RefSolver):
    """run_tortuosity.py:30-51: Na_i is the trace of the SECOND solved ion"""

    def update_ode(self, ode_model):
        K_e = plus(self.c_prev_k.split()[0], self.n_g)
        ode_model.set_parameter('K_e', pcws_constant_project(K_e, self.Q))
        Na_i = minus(self.c_prev_k.split()[1], self.n_g)
        ode_model.set_parameter('Na_i', pcws_constant_project(Na_i, self.Q))


def run_astro_case(nsteps=16, M=8):
    """the problem definition of run_tortuosity.py:81-298 (ions K, Na, Cl with Cl eliminated, rho != 0,
    tortuosity-scaled D, the K+/Na+ source window, membrane models {1: mm_hh, 2: mm_glial, 3: mm_hh},
    stimulus 0 everywhere) on the synthetic mesh knpemidg.mesh.astro_like_mesh(M)"""
    adir = os.path.join(REF_EX, "local-astrocyte-depolarization")
    a_hh = load_by_path("astro_mm_hh", os.path.join(adir, "mm_hh.py"))
    a_glial = load_by_path("astro_mm_glial", os.path.join(adir, "mm_glial.py"))
    mesh, sub, surf = kmesh.astro_like_mesh(M)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    A = ASTRO
    c = A["c"]
    rho = {t: -(c["Na"][t] + c["K"][t] - c["Cl"][t]) for t in range(3)}
    rho_sub = {t: df.Constant(rho[t]) for t in range(3)}
    params = namedtuple('params', ('dt', 'n_steps_ODE', 'F', 'psi', 'C_phi', 'C_M', 'R', 'temperature',
                                   'phi_M_init_type', 'rho_sub'))(
        A["dt"], 25, A["F"], A["F"] / (A["R"] * A["T"]), A["C_M"] / A["dt"], A["C_M"], A["R"], A["T"], 'constant', rho_sub)
    t = df.Constant(0.0)
    lo, hi = mesh.source_box

    def window(sign):
        def f(x):
            inside = np.all((x >= lo[None, :]) & (x <= hi[None, :]), axis=1)
            return sign * A["g_syn"] * inside * (0.2 <= float(t)) * (float(t) <= A["t_syn"])
        return df.Expression(f, degree=4)

    lam = {0: A["lambda_e"], 1: A["lambda_i"], 2: A["lambda_i"]}
    src = {"K": window(1.0), "Na": window(-1.0), "Cl": df.Constant(0)}
    ions = [{'c_init_sub': {tg: df.Constant(c[nm_][tg]) for tg in range(3)}, 'c_init_sub_type': 'constant', 'bdry': None,
             'z': z, 'name': nm_, 'D_sub': {tg: df.Constant(A["D"][nm_] / lam[tg] ** 2) for tg in range(3)},
             'f_source': src[nm_]} for nm_, z in (("K", 1.0), ("Na", 1.0), ("Cl", -1.0))]
    S = RefSolverAstro(params, ions)
    dmesh = df.Mesh(mesh)
    S.setup_domain(dmesh, df.MeshFunction.from_array(dmesh, 3, sub), df.MeshFunction.from_array(dmesh, 2, surf))
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(StimParams(0, {'stim_amplitude': 0}, lambda x: True), {1: a_hh, 2: a_glial, 3: a_hh})
    S.solve_system_active(nsteps * A["dt"], t, SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 1e-40, 0.9, 0.75))
    mem = np.flatnonzero(np.isin(surf, (1, 2, 3)))
    return dict(nsteps=np.array(nsteps), M=np.array(M), mem_facets=mem.astype(np.int32), mem_tag=surf[mem].astype(np.int32),
                phi_M_trace=np.stack(S.trace)[:, mem], final_phi=S.phi.vector().get_local(),
                final_c=S.c.vector().get_local(), final_c_elim=ions[-1]['c'].vector().get_local(),
                final_E=np.stack([ion['E'].vector().get_local()[mem] for ion in ions]), t_end=np.array(float(t)))


# ---- BASELINE configs[4] (bench.py's headline workload): examples/emix-simulations/run_EMIx_simulation.py ------
def run_emix_case(nsteps=15, M=9):
    """the problem definition of run_EMIx_simulation.py:56-259 (ms / cm / mV units, calibrated initial state, ions K, Cl
    and eliminated Na, membrane models {1: mm_glial, 2: mm_hh} of examples/emix-simulations, synaptic stimulus 5 mS/cm^2
    where x < 3e-4 cm, SolverEMIx.update_ode = the hook of RefSolver) on the synthetic block
    knpemidg.mesh.emix_like_mesh(M, n_cells=100, length=1e-3) - the mesh family bench.py's headline number is quoted on"""
    edir = os.path.join(REF_EX, "emix-simulations")
    e_hh = load_by_path("emix_mm_hh", os.path.join(edir, "mm_hh.py"))
    e_glial = load_by_path("emix_mm_glial", os.path.join(edir, "mm_glial.py"))
    mesh, sub, surf = kmesh.emix_like_mesh(M, n_cells=100, length=1.0e-3)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    dt, C_M, temperature, F, R = 0.1, 2.0, 300e3, 96485e3, 8.314e3
    K = (3.3236967382613933, 102.75563828644862, 124.15397583492471)               # ECS (0), glia (1), neuron (2)
    Na = (100.71925900028181, 12.39731187972181, 12.838513108606818)
    c = {"K": K, "Na": Na, "Cl": tuple(a + b for a, b in zip(K, Na))}
    D = {"K": 1.96e-8, "Cl": 2.03e-8, "Na": 1.33e-8}
    params = namedtuple('params', ('dt', 'n_steps_ODE', 'F', 'psi', 'C_phi', 'C_M', 'R', 'temperature',
                                   'phi_M_init_type', 'rho_sub'))(
        dt, 25, F, F / (R * temperature), C_M / dt, C_M, R, temperature, 'constant', {tg: df.Constant(0) for tg in range(3)})
    ions = [{'c_init_sub': {tg: df.Constant(c[nm_][tg]) for tg in range(3)}, 'c_init_sub_type': 'constant',
             'bdry': df.Constant((0, 0)), 'z': z, 'name': nm_, 'D_sub': {tg: df.Constant(D[nm_]) for tg in range(3)},
             'f_source': df.Constant(0)} for nm_, z in (("K", 1.0), ("Cl", -1.0), ("Na", 1.0))]
    S = RefSolver(params, ions)
    dmesh = df.Mesh(mesh)
    S.setup_domain(dmesh, df.MeshFunction.from_array(dmesh, 3, sub), df.MeshFunction.from_array(dmesh, 2, surf))
    S.setup_parameters()
    S.setup_FEM_spaces()
    S.setup_membrane_model(StimParams(5, {'stim_amplitude': 5}, lambda x: x[0] < 3.0e-4), {1: e_glial, 2: e_hh})
    t = df.Constant(0.0)
    S.solve_system_active(nsteps * dt, t, SolverParams(True, True, 0, 1e-5, 1e-7, 1e-40, 2e-40, 0.9, 0.75))
    mem = np.flatnonzero(np.isin(surf, (1, 2)))
    return dict(nsteps=np.array(nsteps), M=np.array(M), mem_facets=mem.astype(np.int32), mem_tag=surf[mem].astype(np.int32),
                phi_M_trace=np.stack(S.trace)[:, mem], final_phi=S.phi.vector().get_local(),
                final_c=S.c.vector().get_local(), final_c_elim=ions[-1]['c'].vector().get_local(),
                final_E=np.stack([ion['E'].vector().get_local()[mem] for ion in ions]),
                final_states_glial=S.mem_models[0]['ode'].states.copy(), final_states_hh=S.mem_models[1]['ode'].states.copy(),
                t_end=np.array(float(t)))


# ---- BASELINE configs[0]: tests/run_MMS_space.py, tests/run_MMS_time.py ------------------------------------
class RefSolverMMS(Solver):
    """the reference Solver as the MMS scripts construct it (run_MMS_space.py:194-195), recording the tensors the
    first PDE step assembles (solve_emi / solve_knp re-assemble in place, solver.py:477-479, 730-731)"""

    def __init__(self, params, ion_list, mms):
        Solver.__init__(self, params=params, ion_list=ion_list, degree_emi=1, degree_knp=1, mms=mms)
        self.step0 = None

    def solve_for_time_step(self, k, t):
        first = self.step0 is None
        if first:       # the EMI forms as assembled (solve_emi removes the mean of its right-hand side in place, solver.py:489-490)
            self.step0 = dict(zip(("A_emi_row", "A_emi_col", "A_emi_val"), coo(df.assemble(self.a_emi))))
            self.step0["b_emi"] = df.assemble(self.L_emi).get_local().copy()
        Solver.solve_for_time_step(self, k, t)
        if first:
            self.step0.update(zip(("A_knp_row", "A_knp_col", "A_knp_val"), coo(self.AA_knp)))
            self.step0.update(b_knp=self.bb_knp.get_local().copy(), phi1=self.phi.vector().get_local().copy(),
                              c1=self.c.vector().get_local().copy())


def run_mms_case(kind="space", resolution=2, dt=1.0e-10, nsteps=2):
    """one pass of the loop body of tests/run_MMS_space.py:24-246 (kind 'space': dt = 1e-10, Tstop = 2 dt) or of
    tests/run_MMS_time.py (kind 'time': fixed mesh, dt = dt_0 / 2^i, Tstop = 2 dt_0): the reference's own
    setup_mms (tests/mms_space.py / mms_time.py), its Solver(mms=...), solve_system_passive with direct solves, and
    the script's L2 error integrals (quadrature degree 5, phi up to its mean), on the mesh of make_mesh_MMS.py"""
    mms_mod = load_by_path("ref_mms_" + kind, os.path.join("/root/reference/tests", "mms_%s.py" % kind))
    C = df.Constant
    t = C(0.0)
    D_a1, D_a2, D_b1, D_b2, D_c1, D_c2 = C(6), C(5), C(3), C(4), C(1), C(2)
    C_a1, C_a2, C_b1, C_b2, C_c1, C_c2 = C(1), C(2), C(2), C(4), C(3), C(2)
    z_a, z_b, z_c = C(1.0), C(-1.0), C(1.0)
    F, C_M, R, temperature = C(1), C(1.0), C(1), C(1)
    rho_sub = {0: C(0), 1: C(0), 2: C(0)}
    fields = ('D_a1', 'D_a2', 'D_b1', 'D_b2', 'D_c1', 'D_c2', 'C_a1', 'C_a2', 'C_b1', 'C_b2', 'C_c1', 'C_c2',
              'C_phi', 'z_a', 'z_b', 'z_c', 'dt', 'F', 'C_M', 'phi_M_init', 'R', 'temperature', 'phi_M_init_type', 'rho_sub')
    params = namedtuple('params', fields)(D_a1, D_a2, D_b1, D_b2, D_c1, D_c2, C_a1, C_a2, C_b1, C_b2, C_c1, C_c2,
                                          C_M / dt, z_a, z_b, z_c, dt, F, C_M, None, R, temperature, 'expression', rho_sub)
    mesh, sub, surf = kmesh.mms_mesh(resolution)
    sub, surf = np.asarray(sub.array()), np.asarray(surf.array())
    dmesh = df.Mesh(mesh)
    subdomains = df.MeshFunction.from_array(dmesh, 2, sub)
    surfaces = df.MeshFunction.from_array(dmesh, 1, surf)
    mms = mms_mod.setup_mms(params, t, dmesh)
    sol, rhs = mms.solution, mms.rhs
    ions = []
    for key, z, name, Ds, Cs in (("a", z_a, "Na", (D_a1, D_a2), (C_a1, C_a2)), ("b", z_b, "K", (D_b1, D_b2), (C_b1, C_b2)),
                                 ("c", z_c, "Cl", (D_c1, D_c2), (C_c1, C_c2))):
        ions.append({'D_sub': {1: C(Ds[0]), 0: C(Ds[1])}, 'z': z, 'name': name,
                     'c_init_sub': {1: sol['c_%s1_init' % key], 0: sol['c_%s2_init' % key]}, 'c_init_sub_type': 'expression',
                     'f1': rhs['volume_c_%s1' % key], 'f2': rhs['volume_c_%s2' % key],
                     'g_robin_1': rhs['bdry']['u_%s1' % key], 'g_robin_2': rhs['bdry']['u_%s2' % key],
                     'bdry': rhs['bdry']['neumann_' + key], 'C_sub': {1: C(Cs[0]), 0: C(Cs[1])}, 'f_source': C(0)})
    S = RefSolverMMS(params, ions, mms)
    S.setup_domain(dmesh, subdomains, surfaces)
    S.setup_parameters()
    S.setup_FEM_spaces()
    sp = SolverParams(True, True, resolution, 1e-6, 1e-7, 1e-40, 1e-40, 0.9, 7.5)
    uh, uh_cc = S.solve_system_passive(nsteps * dt, t, sp, None)
    dX = df.Measure('dx', domain=dmesh, subdomain_data=subdomains)
    md = {'quadrature_degree': 5}

    def l2(e1, e2, u):
        return np.sqrt(abs(df.assemble(df.inner(e2 - u, e2 - u) * dX(0, metadata=md) + df.inner(e1 - u, e1 - u) * dX(1, metadata=md))))

    errors = [l2(sol['c_a1'], sol['c_a2'], uh[0]), l2(sol['c_b1'], sol['c_b2'], uh[1]), l2(sol['c_c1'], sol['c_c2'], uh_cc)]
    mean_e = df.assemble(sol['phi_1'] * dX(1, metadata=md)) + df.assemble(sol['phi_2'] * dX(0, metadata=md))
    mean_a = df.assemble(uh[2] * dX(1, metadata=md)) + df.assemble(uh[2] * dX(0, metadata=md))
    pm = C(mean_e - mean_a)
    errors.append(l2(sol['phi_1'] - pm, sol['phi_2'] - pm, uh[2]))
    out = dict(S.step0)
    out.update(resolution=np.array(resolution), dt=np.array(dt), nsteps=np.array(nsteps),
               errors=np.array(errors), hmin=np.array(dmesh.hmin()), final_phi=S.phi.vector().get_local(),
               final_c=S.c.vector().get_local(), final_c_elim=uh_cc.vector().get_local(), t_end=np.array(float(t)))
    return out


def run_mms_study():
    """ref_mms.npz: the r = 2 space case with its step-0 tensors, the errors of r = 2..5 (run_MMS_space.py prints
    these and the rates between them) and of the time study i = 2..4 on the r = 3 mesh (run_MMS_time.py, dt_0 = 1e-2)"""
    out = {"space2_" + k: v for k, v in run_mms_case("space", 2).items()}
    out.update({"time_" + k: v for k, v in run_mms_case("time", 3, dt=1.0e-2 / 4, nsteps=8).items()})
    out["space_errors"] = np.stack([out["space2_errors"]] + [run_mms_case("space", r)["errors"] for r in (3, 4, 5)])
    out["space_h"] = np.array([float(np.sqrt(2.0)) / 2 ** r for r in (2, 3, 4, 5)])
    out["time_errors"] = np.stack([run_mms_case("time", 3, dt=1.0e-2 / 2 ** i, nsteps=2 * 2 ** i)["errors"] for i in (2, 3, 4)])
    return out


def main(outdir, only=None):
    os.makedirs(outdir, exist_ok=True)

    def want(name):
        return only is None or name in only

    def save(name, maker):
        if want(name):
            np.savez_compressed(os.path.join(outdir, name + ".npz"), **maker())

    m2, s2, f2 = kmesh.neuron_2d_mesh(1)
    s2, f2 = np.asarray(s2.array()), np.asarray(f2.array())
    save("ref_forms_2d", lambda: forms_case("2d", m2, s2, f2, {1: mm_hh}))
    save("ref_forms_2d_nosplit", lambda: forms_case("2d_nosplit", m2, s2, f2, {1: mm_hh}, splitting=False, seed=1))
    m3, s3, f3 = mesh_3d_two_cells()
    save("ref_forms_3d", lambda: forms_case("3d", m3, s3, f3, {1: mm_hh_no_stim, 2: mm_hh}, D_scale={1: 0.5, 2: 0.7},
                                            rho={0: 0.0, 1: 3.0, 2: -2.0}, f_src=[250.0, -125.0], seed=2))
    save("ref_run_2d", run_case)
    save("ref_run_2d_picard", lambda: run_case(nsteps=12, picard=True))
    save("ref_run_2d_emi", run_emi_case)
    save("ref_run_2d_passive", run_passive_case)
    save("ref_run_astro", run_astro_case)          # ~2.5 min (4 224 LSODA calls through scipy)
    save("ref_mms", run_mms_study)
    save("ref_run_emix", run_emix_case)            # ~4 min (3 600 LSODA calls through scipy)
    save("ref_run_3d", run_3d_case)                # the longest one: 11 776 LSODA calls, 3D direct solves with 124 k unknowns
    print("wrote", sorted(f for f in os.listdir(outdir) if f.endswith(".npz")))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--only=")]
    only = [a[len("--only="):].split(",") for a in sys.argv[1:] if a.startswith("--only=")]
    main(args[0] if args else HERE, only[0] if only else None)
