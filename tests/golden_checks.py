"""Checks against the REFERENCE-EXECUTED golden fixtures tests/golden/ref_*.npz.

The fixtures are produced by tests/golden/make_reference_golden.py, which runs the reference's
own unmodified form code, step updates and time loop (/root/reference/src/knpemidg) on the numeric
dolfin stand-in of oracle/refexec.  Three things are compared with them, all on the stored inputs:
the oracle restatement (oracle/forms.py, oracle/stepper.py), the host-emulation build and the CUDA
library (through the C ABI).

Tolerance for assembled tensors: ENTRYWISE, |a - b| <= 1e-12 |b| + 1e-14 max|row| (common.py).
"""
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from common import _lib, entrywise_failures, rel_err  # noqa: F401
from knpemidg import mesh as kmesh
from oracle import forms

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORM_CASES = ("2d", "2d_nosplit", "3d")


class Golden:
    def __init__(self, name):
        self.g = g = np.load(os.path.join(GOLDEN, f"ref_forms_{name}.npz"))
        self.mesh = kmesh.SimplexMesh(g["coords"], g["cells"])
        self.mesh.init_topology()
        self.tags = g["tags"]
        self.splitting = bool(g["splitting"])
        self.mtags = tuple(int(t) for t in g["membrane_tags"])
        D_sub = [{int(t): g["D"][k][i] for i, t in enumerate(self.tags)} for k in range(3)]
        rho_sub = {int(t): g["rho"][i] for i, t in enumerate(self.tags)}
        self.P = forms.Problem(self.mesh, g["cell_tag"], g["facet_tag"], F=float(g["F"]), R=float(g["R"]),
                               T=float(g["T"]), C_M=float(g["C_M"]), C_phi=float(g["C_M"]) / float(g["dt"]),
                               dt=float(g["dt"]), z=g["z"], D_sub=D_sub, rho_sub=rho_sub, membrane_tags=self.mtags)
        P = self.P
        self.n = P.ndof
        mf = P.mem_facets
        self.phi_M, self.I_ch = g["phi_M"][mf], g["I_ch"][:, mf]
        self.f_source = [float(v) for v in g["f_source"]]

    def ref_matrix(self, key, block=None):
        g, n = self.g, self.n
        N = n if key != "A_knp" else 2 * n
        M = sp.coo_matrix((g[key + "_val"], (g[key + "_row"], g[key + "_col"])), shape=(N, N)).tocsr()
        if block is not None:
            M = M[block * n:(block + 1) * n, block * n:(block + 1) * n]
        return M

    def loads(self):
        """f_source v dx(0) of the solved ions as nodal load vectors (solver.py:599)"""
        P = self.P
        out = []
        for f in self.f_source:
            load = np.zeros((P.nc, P.nd))
            ecs = P.cell_tag == 0
            load[ecs] = (f * P.vol[ecs] / (P.d + 1))[:, None]
            out.append(load)
        return out

    def context(self, lib):
        g, P = self.g, self.P
        ctx = _lib.Context(0, lib)
        region = np.searchsorted(self.tags, g["cell_tag"]).astype(np.int32)
        ctx.set_mesh(self.mesh.coords, self.mesh.cells, region, self.mesh.facet_cells, g["facet_tag"], self.mtags)
        ctx.set_params(F=P.F, R=P.R, T=P.T, C_M=P.C_M, C_phi=P.C_phi, dt=P.dt, tau_emi=P.tau, tau_knp=P.tau,
                       Lp=P.Lp, z=list(g["z"]), D=g["D"], rho=list(g["rho"]), splitting=self.splitting)
        for k in range(3):
            ctx.set_field(_lib.F_C, k, g["c_all"][k])
            ctx.set_field(_lib.F_ICH, k, self.I_ch[k])
        for k in range(2):
            ctx.set_field(_lib.F_CN, k, g["c_n"][k])
            if self.f_source[k] != 0.0:
                ctx.set_field(_lib.F_LOAD_KNP, k, self.loads()[k])
        ctx.set_field(_lib.F_PHI, 0, g["phi"])
        ctx.set_field(_lib.F_PHIM, 0, self.phi_M)
        return ctx


def _assert_entrywise(what, a, b, rtol=1e-12):
    # right-hand sides: the ions' contributions to one entry cancel to ~1e-2 of their size (electroneutral
    # state), so the round-off floor of a vector entry is 1e-13 of the largest entry, not 1e-14
    floor = 1e-14 if sp.issparse(b) else 1e-13
    nfail, worst = entrywise_failures(a, b, rtol=rtol, row_floor=floor)
    assert nfail == 0, f"{what}: {nfail} entries off, worst {worst:.2f} x the bound"


def check_oracle_forms(name):
    """oracle/forms.py against what the reference's own form code assembled"""
    G = Golden(name)
    g, P, n = G.g, G.P, G.n
    A, B, b = forms.assemble_emi(P, g["c_all"], G.phi_M, G.I_ch, splitting=G.splitting)
    _assert_entrywise("A_emi", A, G.ref_matrix("A_emi"))
    _assert_entrywise("B_emi", B, G.ref_matrix("B_emi"))
    _assert_entrywise("b_emi", b, g["b_emi"])
    Ak, bk = forms.assemble_knp(P, g["c_all"], g["c_n"], g["phi"], G.phi_M, G.I_ch, splitting=G.splitting,
                                f_source=G.f_source)
    full = G.ref_matrix("A_knp")
    assert abs(full[:n, n:]).sum() == 0.0 and abs(full[n:, :n]).sum() == 0.0    # the ions are uncoupled
    for k in range(2):
        _assert_entrywise(f"A_knp[{k}]", Ak[k], G.ref_matrix("A_knp", k))
        _assert_entrywise(f"b_knp[{k}]", bk[k], g["b_knp"][k * n:(k + 1) * n], rtol=1e-11)
    mf = P.mem_facets
    for k in range(3):
        assert rel_err(forms.nernst(P, g["c_all"][k], P.z[k]), g["E0"][k][mf]) < 1e-13
    # orientation: n_g points from the lower to the higher cell tag (utils.py:80)
    ng = g["n_g"][mf]
    toward_i = P.X[P.mem_cell_i].mean(axis=1) - P.X[P.mem_cell_e].mean(axis=1)
    assert np.all(np.einsum("fk,fk->f", ng, toward_i) > 0)
    # one PDE step (solver.py:794-847)
    x = _solve_singular(G.ref_matrix("A_emi"), g["b_emi"])
    assert _rel_mod_const(g["step_phi"], x) < 1e-9
    phi1 = g["step_phi"].reshape(P.nc, P.nd)
    assert rel_err(forms.membrane_potential(P, phi1), g["step_phi_M"][mf]) < 1e-12
    c1 = g["step_c"].reshape(2, P.nc, P.nd)
    ce = forms.eliminated_concentration(P, c1)
    assert rel_err(ce.ravel(), g["step_c_elim"]) < 1e-13
    for k in range(3):
        ck = c1[k] if k < 2 else ce
        assert rel_err(forms.nernst(P, ck, P.z[k]), g["step_E"][k][mf]) < 1e-12
    return G


def _solve_singular(A, b):
    n = A.shape[0]
    one = sp.csc_matrix(np.ones((n, 1)))
    K = sp.bmat([[A.tocsc(), one], [one.T, None]], format="csc")
    return spla.spsolve(K, np.concatenate([b - b.mean(), [0.0]]))[:n]


def _rel_mod_const(a, b):
    a = a - a.mean()
    b = b - b.mean()
    return np.abs(a - b).max() / np.abs(b).max()


def check_library_forms(lib, name):
    """the library (emulation or CUDA build, through the C ABI) against the reference-executed tensors"""
    G = Golden(name)
    g, P, n = G.g, G.P, G.n
    ctx = G.context(lib)
    mf = P.mem_facets
    ctx.post_step(_lib.POST_NERNST)
    for k in range(3):
        assert rel_err(ctx.get_field(_lib.F_NERNST, k), g["E0"][k][mf]) < 1e-12
    ctx.assemble_emi()
    _assert_entrywise("A_emi", ctx.matrix(0), G.ref_matrix("A_emi"))
    _assert_entrywise("B_emi", ctx.matrix(1), G.ref_matrix("B_emi"))
    _assert_entrywise("b_emi", ctx.get_field(_lib.F_RHS_EMI), g["b_emi"])
    ctx.assemble_knp()
    for k in range(2):
        _assert_entrywise(f"A_knp[{k}]", ctx.matrix(2 + k), G.ref_matrix("A_knp", k))
        _assert_entrywise(f"b_knp[{k}]", ctx.get_field(_lib.F_RHS_KNP, k), g["b_knp"][k * n:(k + 1) * n], rtol=1e-11)
    # one PDE step with tight Krylov tolerances against the reference's direct solves
    ctx.amg_setup()
    ctx.solver_options(pc=1)
    ctx.solve_emi(rtol=1e-12, atol=1e-40, maxit=2000)
    assert _rel_mod_const(ctx.get_field(_lib.F_PHI), g["step_phi"]) < 1e-8
    ctx.assemble_knp()
    ctx.solve_knp(rtol=1e-13, atol=1e-40, maxit=2000)
    for k in range(2):
        assert rel_err(ctx.get_field(_lib.F_C, k), g["step_c"][k * n:(k + 1) * n]) < 1e-9
    ctx.post_step(_lib.POST_ALL)
    assert rel_err(ctx.get_field(_lib.F_PHIM), g["step_phi_M"][mf]) < 1e-8
    assert rel_err(ctx.get_field(_lib.F_C, 2), g["step_c_elim"]) < 1e-9
    for k in range(3):
        assert rel_err(ctx.get_field(_lib.F_NERNST, k), g["step_E"][k][mf]) < 1e-8
    return ctx


# ---- the reference's time loop (ref_run_2d.npz) -------------------------------------------------
RUN_PHYS = dict(F=96485.0, R=8.314, T=300.0, C_M=0.02, C_phi=0.02 / 1.0e-4, dt=1.0e-4, z=[1.0, -1.0, 1.0],
                D_sub=[{0: 1.96e-9, 1: 1.96e-9}, {0: 2.03e-9, 1: 2.03e-9}, {0: 1.33e-9, 1: 1.33e-9}],
                rho_sub={0: 0.0, 1: 0.0})
NA_I, NA_E, K_I, K_E = 12.838513108648856, 100.71925900027354, 124.15397583491901, 3.3236967382705265
RUN_C_INIT = [{1: K_I, 0: K_E}, {1: NA_I + K_I, 0: NA_E + K_E}, {1: NA_I, 0: NA_E}]


def run_golden():
    return np.load(os.path.join(GOLDEN, "ref_run_2d.npz"))


def trace_deviation(trace, ref):
    """max over steps and facets of |phi_M - ref| relative to the range of the reference trace"""
    return float(np.abs(trace - ref).max() / (ref.max() - ref.min()))


def oracle_run(convention, nsteps):
    from knpemidg.models import mm_hh
    from oracle import stepper
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    P = forms.Problem(mesh, sub.array(), surf.array(), membrane_tags=(1,), **RUN_PHYS)
    c0 = np.stack([np.where((sub.array() == 1)[:, None], ci[1], ci[0]) * np.ones((P.nc, P.nd)) for ci in RUN_C_INIT])
    O = stepper.OracleSolver(P, c0, models={1: mm_hh}, stimulus={"stim_amplitude": 10.0},
                             stimulus_locator=lambda x: x[0] < 20e-6, ion_names=["K", "Cl", "Na"], direct=True,
                             current_convention=convention)
    tr = []
    for _ in range(nsteps):
        O.step()
        tr.append(O.phi_M.copy())
    return np.stack(tr), O


def library_run_picard(lib, nsteps, rtol_emi=1e-10, rtol_knp=1e-11):
    """the run of library_run with Engine.pde_phase_picard (solve_for_time_step_picard, solver.py:850-927)
    as the PDE step; returns the traces and the Picard iteration counts"""
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1,), lib=lib, **RUN_PHYS)
    eng.set_concentrations_by_tag(RUN_C_INIT)
    eng.add_membrane_model(1, mm_hh, ["K", "Cl", "Na"], stimulus={"stim_amplitude": 10.0},
                           stimulus_locator=lambda x: x[0] < 20e-6)
    eng.rtol_emi, eng.rtol_knp = rtol_emi, rtol_knp
    eng.initialize(pc=1)
    tr, its = [], []
    for _ in range(nsteps):
        eng.ode_phase()
        its.append(eng.pde_phase_picard())
        eng.k += 1
        tr.append(eng.phi_M().copy())
    return np.stack(tr), its, eng


def library_run(lib, nsteps, rtol_emi=1e-5, rtol_knp=1e-7):
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1,), lib=lib, **RUN_PHYS)
    eng.set_concentrations_by_tag(RUN_C_INIT)
    eng.add_membrane_model(1, mm_hh, ["K", "Cl", "Na"], stimulus={"stim_amplitude": 10.0},
                           stimulus_locator=lambda x: x[0] < 20e-6)
    eng.rtol_emi, eng.rtol_knp = rtol_emi, rtol_knp
    eng.initialize(pc=1)
    tr = []
    for _ in range(nsteps):
        eng.step()
        tr.append(eng.phi_M().copy())
    return np.stack(tr), eng


# ---- BASELINE configs[2]: run_3D.py on its own resolution-0 mesh (ref_run_3d.npz) ----------------------------
def library_run_3d(lib, nsteps, rtol_emi=1e-10, rtol_knp=1e-11):
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh, mm_hh_no_stim
    mesh, sub, surf = kmesh.bundle_3d_mesh(0)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), lib=lib, **RUN_PHYS)
    eng.set_concentrations_by_tag(RUN_C_INIT)
    for tag, mod in ((1, mm_hh), (2, mm_hh_no_stim)):                                 # run_3D.py:196
        eng.add_membrane_model(tag, mod, ["K", "Cl", "Na"], stimulus={"stim_amplitude": 10.0},
                               stimulus_locator=lambda x: x[0] < 20e-6)
    eng.rtol_emi, eng.rtol_knp = rtol_emi, rtol_knp
    eng.initialize(pc=1)
    tr = []
    for _ in range(nsteps):
        eng.step()
        tr.append(eng.phi_M().copy())
    return np.stack(tr), eng


def check_library_3d(lib):
    """the four-axon bundle of run_3D.py (mm_hh on the first axon, mm_hh_no_stim on the other three) against the
    reference's own solve_system_active on the same mesh"""
    g = np.load(os.path.join(GOLDEN, "ref_run_3d.npz"))
    tr, eng = library_run_3d(lib, int(g["nsteps"]))
    assert np.array_equal(eng.mem["facet"], g["mem_facets"])
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    out = dict(trace=trace_deviation(tr, g["phi_M_trace"]), c=rel_err(cfin, g["final_c"]),
               c_elim=rel_err(eng.concentration(2).reshape(-1), g["final_c_elim"]),
               phi=_rel_mod_const(eng.phi().reshape(-1), g["final_phi"]))
    assert out["trace"] < 1e-6 and out["c"] < 1e-7 and out["c_elim"] < 1e-7 and out["phi"] < 1e-6, out
    return out


# ---- BASELINE configs[3]: run_tortuosity.py on knpemidg.mesh.astro_like_mesh (ref_run_astro.npz) ---------
ASTRO = dict(dt=0.1, C_M=1.0, T=307e3, F=96500e3, R=8.315e3, g_syn=26.0, t_syn=1.2, lambda_i=3.2 * 4, lambda_e=1.6 * 4)
ASTRO_D = [1.96e-8, 1.33e-8, 2.03e-8]                                                # K, Na, Cl (cm^2/ms)
ASTRO_C = [(3.092970607490389, 124.13988964240784, 99.3100014897692),               # K: ECS, neuron, glia
           (144.60625137617149, 12.850454639128186, 15.775818906083778),             # Na
           (133.62525154406637, 5.0, 5.203660274163705)]                             # Cl (eliminated)
ASTRO_NAMES = ["K", "Na", "Cl"]
ASTRO_LINKS = (("K_e", 0, "plus"), ("Na_i", 1, "minus"))                             # run_tortuosity.py:38-49


def astro_golden():
    return np.load(os.path.join(GOLDEN, "ref_run_astro.npz"))


def astro_problem_args():
    A = ASTRO
    lam = [A["lambda_e"], A["lambda_i"], A["lambda_i"]]
    D_sub = [{t: D / lam[t] ** 2 for t in range(3)} for D in ASTRO_D]
    rho_sub = {t: -(ASTRO_C[1][t] + ASTRO_C[0][t] - ASTRO_C[2][t]) for t in range(3)}
    return dict(F=A["F"], R=A["R"], T=A["T"], C_M=A["C_M"], C_phi=A["C_M"] / A["dt"], dt=A["dt"], z=[1.0, 1.0, -1.0],
                D_sub=D_sub, rho_sub=rho_sub)


def astro_source(mesh, sign):
    lo, hi = mesh.source_box

    def f(x, t):
        x = np.asarray(x)
        inside = np.all((x >= lo) & (x <= hi), axis=-1)
        return sign * ASTRO["g_syn"] * inside * (0.2 <= t) * (t <= ASTRO["t_syn"])
    return f


def oracle_run_astro(nsteps, M):
    from knpemidg.models import mm_glial_astro, mm_hh_astro
    from oracle import stepper
    mesh, sub, surf = kmesh.astro_like_mesh(M)
    P = forms.Problem(mesh, sub.array(), surf.array(), membrane_tags=(1, 2, 3), **astro_problem_args())
    tag = sub.array()
    c0 = np.stack([np.choose(tag, ci)[:, None] * np.ones((P.nc, P.nd)) for ci in ASTRO_C])
    O = stepper.OracleSolver(P, c0, models={1: mm_hh_astro, 2: mm_glial_astro, 3: mm_hh_astro},
                             stimulus={"stim_amplitude": 0.0}, stimulus_locator=lambda x: True, ion_names=ASTRO_NAMES,
                             direct=True, f_source=[astro_source(mesh, 1.0), astro_source(mesh, -1.0)],
                             ode_links=list(ASTRO_LINKS), current_convention="last_call")
    tr = []
    for _ in range(nsteps):
        O.step()
        tr.append(O.phi_M.copy())
    return np.stack(tr), O


def library_run_astro(lib, nsteps, M, rtol_emi=1e-5, rtol_knp=1e-7):
    from knpemidg.engine import Engine
    from knpemidg.models import mm_glial_astro, mm_hh_astro
    mesh, sub, surf = kmesh.astro_like_mesh(M)
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=(1, 2, 3), lib=lib, **astro_problem_args())
    eng.set_concentrations_by_tag([{t: ci[t] for t in range(3)} for ci in ASTRO_C])
    for tag_, mod in ((1, mm_hh_astro), (2, mm_glial_astro), (3, mm_hh_astro)):
        eng.add_membrane_model(tag_, mod, ASTRO_NAMES, stimulus={"stim_amplitude": 0.0}, links=ASTRO_LINKS)
    eng.rtol_emi, eng.rtol_knp = rtol_emi, rtol_knp
    src = [astro_source(mesh, 1.0), astro_source(mesh, -1.0)]
    eng.initialize(pc=1)
    tr = []
    for _ in range(nsteps):
        for k in range(2):                              # sources at the OLD time (t.assign comes last, solver.py:845)
            eng.set_source(k, lambda x, f=src[k], t=eng.t: float(f(x, t)))
        eng.step()
        tr.append(eng.phi_M().copy())
    return np.stack(tr), eng


# ---- BASELINE configs[4], the workload of bench.py's headline number (ref_run_emix.npz) ---------------------------
def emix_golden():
    return np.load(os.path.join(GOLDEN, "ref_run_emix.npz"))


def library_run_emix(lib, nsteps, M, rtol_emi=1e-10, rtol_knp=1e-11):
    """bench.py's own engine builder (build_engine_emix) at a small M, stepped with tight Krylov tolerances"""
    import bench
    eng = bench.build_engine_emix(M, 0, lib=lib)
    eng.rtol_emi, eng.rtol_knp = rtol_emi, rtol_knp
    tr = []
    for _ in range(nsteps):
        eng.step()
        tr.append(eng.phi_M().copy())
    return np.stack(tr), eng


def oracle_run_emix(nsteps, M):
    """the oracle loop (oracle/stepper.py, direct solves, scipy LSODA) on the same problem"""
    import bench
    from knpemidg.models import mm_glial_emix, mm_hh_emix
    from oracle import stepper
    mesh, sub, surf = kmesh.emix_like_mesh(M, n_cells=100, length=1.0e-3)
    P = forms.Problem(mesh, sub.array(), surf.array(), membrane_tags=(1, 2), **bench.EMIX_PHYS)
    c0 = np.stack([np.choose(sub.array(), [ci[0], ci[1], ci[2]])[:, None] * np.ones((P.nc, P.nd)) for ci in bench.EMIX_C_INIT])
    O = stepper.OracleSolver(P, c0, models={1: mm_glial_emix, 2: mm_hh_emix}, stimulus={"stim_amplitude": 5.0},
                             stimulus_locator=lambda x: x[0] < 3.0e-4, ion_names=["K", "Cl", "Na"], direct=True)
    tr = []
    for _ in range(nsteps):
        O.step()
        tr.append(O.phi_M.copy())
    return np.stack(tr), O


def check_library_emix(lib):
    """the engine bench.py times, on the block emix_like_mesh(9), against the reference's own solve_system_active on the
    problem of run_EMIx_simulation.py (15 steps: the stimulated neuron fires, -74 -> +48 mV; the glial membrane stays
    within 1 uV of -83.085 mV)"""
    g = emix_golden()
    tr, eng = library_run_emix(lib, int(g["nsteps"]), int(g["M"]))
    assert np.array_equal(eng.mem["facet"], g["mem_facets"])
    dev = trace_deviation(tr, g["phi_M_trace"])
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    out = dict(trace=dev, c=rel_err(cfin, g["final_c"]), c_elim=rel_err(eng.concentration(2).reshape(-1), g["final_c_elim"]),
               phi=_rel_mod_const(eng.phi().reshape(-1), g["final_phi"]))
    assert out["trace"] < 1e-6 and out["c"] < 1e-7 and out["c_elim"] < 1e-7 and out["phi"] < 1e-6, out
    return out


# ---- the reference's manufactured-solution study (ref_mms.npz; tests/run_MMS_space.py, run_MMS_time.py) --------
def mms_golden():
    return np.load(os.path.join(GOLDEN, "ref_mms.npz"))


def _coo(g, prefix, shape):
    return sp.coo_matrix((g[prefix + "_val"], (g[prefix + "_row"], g[prefix + "_col"])), shape=shape).tocsr()


def script_errors(S, L, t=0.0):
    """the four L2 errors as the reference's scripts integrate them (quadrature degree 5, phi modulo its mean;
    run_MMS_space.py:207-246): c_a, c_b, c_c (the eliminated ion), phi"""
    mm, P = L.mm, L.P
    mm.t = t
    uh = S.c.split() + (S.phi,)
    return np.array([mm.l2_error(P, uh[0].nodal(), "c", 0, degree=5), mm.l2_error(P, uh[1].nodal(), "c", 1, degree=5),
                     mm.l2_error(P, S.ion_list[-1]["c"].nodal(), "c", 2, degree=5),
                     mm.l2_error(P, uh[2].nodal(), "phi", mean_free=True, degree=5)])


def _mms_rhs_close(what, P, b, ref):
    """entries of cells at the interface carry C_phi g (1e9 at dt = 1e-10), the others the O(1) volume sources: the two
    groups are compared separately, each to 1e-11 of its own largest entry"""
    at = np.zeros(P.nc, dtype=bool)
    at[P.mem_cell_i] = at[P.mem_cell_e] = True
    d, r = np.abs(b - ref).reshape(P.nc, P.nd), np.abs(ref).reshape(P.nc, P.nd)
    for name, sel in (("interface cells", at), ("other cells", ~at)):
        assert d[sel].max() <= 1e-11 * r[sel].max(), (what, name, d[sel].max(), r[sel].max())


def check_oracle_mms():
    """oracle/mms.py + oracle/forms.py (MMS mode) against what the reference's own MMS code assembled at step 0 of
    the r = 2 space case: matrices entrywise 1e-12 (A_knp on the reference's own phi), right-hand sides 1e-11 with
    the sin/cos sources integrated by the rule of the degree UFL estimates (13), the fields after the two direct
    solves (phi modulo constants: the interface coupling C_phi = 1e10 makes the EMI system ill-conditioned)."""
    from oracle import mms as omms, stepper
    g = mms_golden()
    mesh, sub, surf = kmesh.mms_mesh(2)
    mm = omms.MMS("space", dt=float(g["space2_dt"]), ufl_degree=13)
    P = forms.Problem(mesh, sub.array(), surf.array(), **mm.problem_kwargs())
    n = P.ndof
    c0 = np.stack([mm.exact_field(P, "c", k, t=0.0) for k in range(3)])
    O = stepper.OracleSolver(P, c0, mms=mm, splitting=False)
    mm.t = 0.0
    A, B, b = O.assemble_emi()
    _assert_entrywise("A_emi (mms)", A, _coo(g, "space2_A_emi", (n, n)))
    _mms_rhs_close("b_emi", P, b, g["space2_b_emi"])
    O.solve_emi()
    assert _rel_mod_const(O.phi.ravel(), g["space2_phi1"]) < 1e-6
    O.phi = g["space2_phi1"].reshape(P.nc, P.nd).copy()          # the drift terms of A_knp on the reference's own phi
    Ak, bk = O.assemble_knp()
    full = _coo(g, "space2_A_knp", (2 * n, 2 * n))
    assert abs(full[:n, n:]).sum() == 0.0 and abs(full[n:, :n]).sum() == 0.0
    for k in range(2):
        _assert_entrywise(f"A_knp[{k}] (mms)", Ak[k], full[k * n:(k + 1) * n, k * n:(k + 1) * n])
        _mms_rhs_close(f"b_knp[{k}]", P, bk[k], g["space2_b_knp"][k * n:(k + 1) * n])
    O.solve_knp()
    assert rel_err(O.c[:2].ravel(), g["space2_c1"]) < 1e-12


def check_library_mms(lib, resolutions=(2, 3, 4), time_levels=(2, 3)):
    """the product path (Solver(mms=...) -> engine -> C ABI) against the reference-executed study: the step-0 matrices
    of the r = 2 case entrywise, and the L2 errors the reference's scripts print, for the space study at
    `resolutions` and the time study at dt_0 / 2^i, i in `time_levels` (r = 3 mesh)"""
    import solver_checks as sc
    g = mms_golden()
    out = {}
    errs, S, L = sc.run_mms(lib, 2, nsteps=1, ufl_degree=13)               # the matrices of the first step stay in the context
    n = L.P.ndof
    ctx = S.engine.ctx
    _assert_entrywise("A_emi (mms)", ctx.matrix(0), _coo(g, "space2_A_emi", (n, n)))
    full = _coo(g, "space2_A_knp", (2 * n, 2 * n))
    for k in range(2):
        _assert_entrywise(f"A_knp[{k}] (mms)", ctx.matrix(2 + k), full[k * n:(k + 1) * n, k * n:(k + 1) * n])
    out["phi1"] = _rel_mod_const(S.phi.nodal().ravel(), g["space2_phi1"])
    out["c1"] = rel_err(np.concatenate([f.nodal().ravel() for f in S.c.split()]), g["space2_c1"])
    ref_r = {2: 0, 3: 1, 4: 2, 5: 3}
    assert out["phi1"] < 1e-6 and out["c1"] < 1e-12, out
    ref_r = {2: 0, 3: 1, 4: 2, 5: 3}
    out["space"], out["time"] = [], []
    for r in resolutions:
        _, S, L = sc.run_mms(lib, r, ufl_degree=13)
        out["space"].append(float(np.abs(script_errors(S, L) / g["space_errors"][ref_r[r]] - 1.0).max()))
    for i in time_levels:
        _, S, L = sc.run_mms_time(lib, i, r=3, script_initial_data=True)
        out["time"].append(float(np.abs(script_errors(S, L, t=2.0e-2) / g["time_errors"][i - 2] - 1.0).max()))
    # measured (emulation and B200): space 5e-8 .. 1.5e-7, time 8e-9 .. 2e-8
    assert max(out["space"] + out["time"]) < 1e-6, out
    return out
