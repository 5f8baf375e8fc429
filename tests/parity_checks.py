"""Parity checks shared by the CPU (host-emulation) and GPU test modules: every
function takes the loaded library, builds a seeded case, runs the hot path
through the C ABI and compares with the oracle (oracle/).

Tolerances: assembled matrix entries ENTRY BY ENTRY, |a - b| <= 1e-12 |b| + 1e-14 max|row|
(north_star: "within 1e-12 relative (fp64)"; the row-relative floor is the round-off of the
sum of contributions that cancel), right-hand sides 1e-12 |b| + 1e-13 max|b|;
linear solutions against a sparse direct solve at the level the KSP tolerance
allows; ODE states against a 1e-10 LSODA solve at 1e-6 (north_star trace
tolerance).
"""
import importlib

import numpy as np
import scipy.sparse.linalg as spla

from common import Case, rel_err, entrywise_failures, _lib
from oracle import forms, ode as oracle_ode
from oracle.stepper import solve_singular_direct

TOL_ASM = 1e-12


def assert_entrywise(what, a, b, rtol=TOL_ASM):
    import scipy.sparse as sp
    nfail, worst = entrywise_failures(a, b, rtol=rtol, row_floor=1e-14 if sp.issparse(b) else 1e-13)
    assert nfail == 0, f"{what}: {nfail} entries outside the entrywise bound, worst {worst:.2f} x"


def check_assembly(lib, name, splitting=True, D_scale=(1.0, 1.0), seed=0):
    cs = Case(name, lib, seed=seed, splitting=splitting, D_scale=D_scale)
    P, ctx = cs.P, cs.ctx
    A, B, b = forms.assemble_emi(P, cs.c_all, cs.phi_M, cs.I_ch, splitting=splitting)
    ctx.assemble_emi()
    Ag, Bg = ctx.matrix(0), ctx.matrix(1)
    assert rel_err(Ag, A) < TOL_ASM
    assert rel_err(Bg, B) < TOL_ASM
    assert rel_err(ctx.get_field(_lib.F_RHS_EMI), b) < TOL_ASM
    assert_entrywise("A_emi", Ag, A)
    assert_entrywise("B_emi", Bg, B)
    assert_entrywise("b_emi", ctx.get_field(_lib.F_RHS_EMI), b)
    # structure: symmetric, constants in the null space (solver.py:465-466)
    assert abs(Ag - Ag.T).max() < 1e-12 * abs(Ag).max()
    assert np.abs(Ag @ np.ones(P.ndof)).max() < 1e-10 * abs(Ag).max()
    As, bs = forms.assemble_knp(P, cs.c_all, cs.c_n, cs.phi, cs.phi_M, cs.I_ch, splitting=splitting)
    ctx.assemble_knp()
    # membrane part of the KNP rhs is ~1e-4 of the mass part: check it separately
    _, bs0 = forms.assemble_knp(P, cs.c_all, cs.c_n, 0 * cs.phi, 0 * cs.phi_M, 0 * cs.I_ch,
                                splitting=splitting)
    for k in range(P.N_ions):
        assert rel_err(ctx.matrix(2 + k), As[k]) < TOL_ASM
        bg = ctx.get_field(_lib.F_RHS_KNP, k)
        assert rel_err(bg, bs[k]) < TOL_ASM
        assert_entrywise(f"A_knp[{k}]", ctx.matrix(2 + k), As[k])
        assert_entrywise(f"b_knp[{k}]", bg, bs[k], rtol=1e-11)
        mem = bs[k] - bs0[k]
        assert np.abs((bg - bs0[k]) - mem).max() < max(1e-8 * np.abs(mem).max(), 1e-13 * np.abs(bs[k]).max())
    # SpMV kernel against the exported matrix
    x = np.random.default_rng(seed + 1).standard_normal(P.ndof)
    for which in (0, 1, 3):                             # error relative to |M||x| (cancellation-safe)
        M = ctx.matrix(which)
        scale = (abs(M) @ np.abs(x)).max()
        assert np.abs(ctx.spmv(which, x) - M @ x).max() < 1e-14 * scale
    return cs


def check_post_step(lib, name):
    cs = Case(name, lib)
    P, ctx = cs.P, cs.ctx
    ctx.post_step()
    assert rel_err(ctx.get_field(_lib.F_PHIM), forms.membrane_potential(P, cs.phi)) < 1e-12
    ce = forms.eliminated_concentration(P, cs.c_all[:2])
    assert rel_err(ctx.get_field(_lib.F_C, 2), ce.ravel()) < 1e-14
    call = cs.c_all.copy()
    call[2] = ce
    for k in range(3):
        assert rel_err(ctx.get_field(_lib.F_NERNST, k), forms.nernst(P, call[k], P.z[k])) < 1e-12
    assert rel_err(ctx.facet_trace(_lib.F_C, 0, 0), forms.facet_mean_trace(P, cs.c_all[0], "plus")) < 1e-13
    assert rel_err(ctx.facet_trace(_lib.F_C, 1, 1), forms.facet_mean_trace(P, cs.c_all[1], "minus")) < 1e-13


def check_solvers(lib, name, pcs=(0, 1), emi_tol=1e-6, knp_tol=1e-7, max_emi_it=None):
    cs = Case(name, lib)
    P, ctx = cs.P, cs.ctx
    ctx.assemble_emi()
    A, B, b = forms.assemble_emi(P, cs.c_all, cs.phi_M, cs.I_ch)
    xs = solve_singular_direct(A, b)
    out = {}
    for pc in pcs:
        ctx.set_field(_lib.F_PHI, 0, cs.phi)
        if pc == 1:
            ctx.amg_setup()
            rows, nnz = ctx.amg_info()
            assert rows[0] == P.ndof and rows[-1] <= 1024 and all(a > b_ for a, b_ in zip(rows, rows[1:]))
        ctx.solver_options(pc=pc)
        it, res = ctx.solve_emi(rtol=1e-5, atol=1e-40, maxit=4000)   # reference tolerance (run_2D.py:187)
        out[("emi", pc)] = it
        ctx.solve_emi(rtol=1e-11, atol=1e-40, maxit=4000)            # continue to a tight solve
        x = ctx.get_field(_lib.F_PHI)
        # the EMI potential is defined up to a constant (pure Neumann)
        assert np.abs((x - x.mean()) - xs).max() < emi_tol * np.abs(xs).max()
    if max_emi_it is not None:
        assert out[("emi", 1)] <= max_emi_it
    ctx.assemble_knp()
    phi = ctx.get_field(_lib.F_PHI).reshape(P.nc, P.nd)
    As, bs = forms.assemble_knp(P, cs.c_all, cs.c_n, phi, cs.phi_M, cs.I_ch)
    xd = [spla.spsolve(As[k].tocsc(), bs[k]) for k in range(2)]
    for pc in pcs:
        for k in range(2):
            ctx.set_field(_lib.F_C, k, cs.c_all[k])
        ctx.solver_options(pc=pc)
        it, res = ctx.solve_knp(rtol=1e-7, atol=1e-40, maxit=4000)   # reference tolerance
        assert it >= 5                                   # ksp_min_it (solver.py:686)
        ctx.solve_knp(rtol=1e-12, atol=1e-40, maxit=4000)
        for k in range(2):
            assert rel_err(ctx.get_field(_lib.F_C, k), xd[k]) < knp_tol
        out[("knp", pc)] = it
    return out


def _model_setup(mod, nm, rng):
    S = np.tile(mod.init_state_values(), (nm, 1))
    Pm = np.tile(mod.init_parameter_values(), (nm, 1))
    names = mod.PARAMETER_NAMES
    si = mod.state_indices("V")
    mv_units = abs(S[0, si]) > 1.0
    S[:, si] *= 1 + 0.05 * rng.uniform(-1, 1, nm)

    def setp(key, val):
        if key in names:
            Pm[:, mod.parameter_indices(key)] = val

    sc = 1e3 if mv_units else 1.0
    setp("Cm", 1.0 if mv_units else 0.02)
    setp("E_Na", 0.054 * sc); setp("E_K", -0.088 * sc); setp("E_Cl", -0.07 * sc)
    setp("K_e", 4.0); setp("Na_i", 12.0)
    setp("stim_amplitude", 0.5 if mv_units else 10.0)
    return S, Pm, si, (0.1 if mv_units else 1e-4)


def check_ode(lib, model_names, nsteps=4):
    cs = Case("2d", lib)
    P, ctx = cs.P, cs.ctx
    nm = P.nm
    models = lib.models()
    for name in model_names:
        mod = importlib.import_module("knpemidg.models." + name)
        mid, ns, npar = models[name]
        assert ns == len(mod.init_state_values()) and npar == len(mod.init_parameter_values())
        S, Pm, si, dt = _model_setup(mod, nm, np.random.default_rng(3))
        h = ctx.membrane_register(mid, np.arange(nm), S, Pm)
        ich = [mod.parameter_indices("I_ch_" + n) for n in ("K", "Cl", "Na")]
        ctx.membrane_outputs(h, si, ich)
        S2, P2 = S.copy(), Pm.copy()
        t = 0.0
        for _ in range(nsteps):
            ctx.ode_step(h, t, dt, rtol=1e-8, atol=0.0, set_v=False)
            oracle_ode.step_rows(mod, S2, P2, t, dt, rtol=1e-10, atol=1e-14)
            t += dt
        Sg = ctx.membrane_get(h, "states", (nm, ns))
        Pg = ctx.membrane_get(h, "params", (nm, npar))
        for j in range(ns):
            assert rel_err(Sg[:, j], S2[:, j]) < 1e-6
        for kk, col in enumerate(ich):
            assert np.abs(Pg[:, col] - P2[:, col]).max() <= 1e-6 * max(np.abs(P2[:, ich]).max(), 1e-300)
            assert np.array_equal(ctx.get_field(_lib.F_ICH, kk), Pg[:, col])     # ODE -> PDE scatter
        assert np.array_equal(ctx.get_field(_lib.F_PHIM), Sg[:, si])


def check_ode_links(lib):
    """PDE -> ODE gathers (solver.py:1094-1101) and the stimulus mask (membrane.py:102-104)."""
    cs = Case("2d", lib)
    P, ctx = cs.P, cs.ctx
    mod = importlib.import_module("knpemidg.models.mm_hh")
    mid, ns, npar = lib.models()["mm_hh"]
    nm = P.nm
    rows = np.arange(0, nm, 2)
    S = np.tile(mod.init_state_values(), (len(rows), 1))
    Pm = np.tile(mod.init_parameter_values(), (len(rows), 1))
    Pm[:, mod.parameter_indices("Cm")] = 0.02
    h = ctx.membrane_register(mid, rows, S, Pm)
    ctx.post_step()                                   # fills phi_M, E_k on the device
    E = [ctx.get_field(_lib.F_NERNST, k) for k in range(3)]
    phiM = ctx.get_field(_lib.F_PHIM)
    for k, nme in enumerate(("K", "Cl", "Na")):
        ctx.membrane_link(h, mod.parameter_indices("E_" + nme), 0, _lib.F_NERNST, k)
    ctx.membrane_link(h, mod.parameter_indices("K_e"), 1, _lib.F_C, 0, 0)
    ctx.membrane_link(h, mod.parameter_indices("Na_i"), 1, _lib.F_C, 2, 1)
    ctx.membrane_outputs(h, mod.state_indices("V"), [mod.parameter_indices("I_ch_" + n) for n in ("K", "Cl", "Na")])
    mask = np.zeros(len(rows), dtype=np.uint8)
    mask[::3] = 1
    ctx.membrane_stimulus(h, mask, [mod.parameter_indices("stim_amplitude")], [10.0])
    Ke = ctx.facet_trace(_lib.F_C, 0, 0)
    Nai = ctx.facet_trace(_lib.F_C, 2, 1)
    ctx.ode_step(h, 0.0, 1e-4, set_v=True)
    Pg = ctx.membrane_get(h, "params", (len(rows), npar))
    Sg = ctx.membrane_get(h, "states", (len(rows), ns))
    for k, nme in enumerate(("K", "Cl", "Na")):
        assert np.array_equal(Pg[:, mod.parameter_indices("E_" + nme)], E[k][rows])
    assert np.array_equal(Pg[:, mod.parameter_indices("K_e")], Ke[rows])
    assert np.array_equal(Pg[:, mod.parameter_indices("Na_i")], Nai[rows])
    assert np.array_equal(Pg[:, mod.parameter_indices("stim_amplitude")], 10.0 * mask)
    # oracle step from the same gathered inputs
    S2, P2 = S.copy(), Pm.copy()
    S2[:, mod.state_indices("V")] = phiM[rows]
    for k, nme in enumerate(("K", "Cl", "Na")):
        P2[:, mod.parameter_indices("E_" + nme)] = E[k][rows]
    P2[:, mod.parameter_indices("K_e")] = Ke[rows]
    P2[:, mod.parameter_indices("Na_i")] = Nai[rows]
    oracle_ode.step_rows(mod, S2, P2, 0.0, 1e-4, stim_mask=mask.astype(bool),
                         stimulus={"stim_amplitude": 10.0}, rtol=1e-10, atol=1e-14)
    assert rel_err(Sg, S2) < 1e-6
    # rows that carry no ODE point keep their phi_M
    others = np.setdiff1d(np.arange(nm), rows)
    assert np.array_equal(ctx.get_field(_lib.F_PHIM)[others], phiM[others])


def check_calibration_kat(lib, nsteps=100000):
    """KNOWN ANSWER from the reference's own run: examples/emix-simulations/run_calibration.py
    integrates mm_calibration for 100 000 steps of 0.1 ms (LSODA, rtol 1e-8) and its outcome is
    hard-coded as the initial state of examples/emix-simulations/mm_hh.py:11-14 (m, h, n, phi_M),
    restated in knpemidg.models.mm_hh_emix and pinned to the reference file by
    tests/golden/ode_rhs_golden.json.  The same run through the library's ODE kernel (a
    free-standing MembraneModel, one ODE point) must land on those numbers."""
    from knpemidg import mesh as kmesh
    from knpemidg.membrane import MembraneModel
    from knpemidg.models import mm_calibration, mm_hh_emix

    class Space:                                  # any facet space of the mesh; carries the library under test
        pass
    V = Space()
    V.lib = lib
    mesh = kmesh.rectangle_mesh((0.0, 0.0), (1.0, 1.0), 1, 1)
    facet_f = kmesh.MeshFunction(mesh, 1, 7)
    facet_f.array()[0] = 0
    m = MembraneModel(mm_calibration, facet_f=facet_f, tag=0, V=V)
    assert m.nodes == 1
    m.step_lsoda(dt=0.1, stimulus={"stim_amplitude": 0})
    ctx = m.engine.ctx
    for k in range(1, nsteps):
        ctx.ode_step(m.handle, 0.1 * k, 0.1, 1e-8, 0.0, False)
    got = m.states[0]
    want = mm_hh_emix.init_state_values()                        # m, h, n, V
    assert rel_err(got[:4], want) < 1e-11, (got[:4], want)
    assert np.abs(got[:4] / want - 1.0).max() < 1e-10            # each component, not only the norm
    # the remaining steady-state values as run_calibration.py prints them (this build, 12 digits)
    assert abs(got[4] + 83.08244665955735) < 1e-8                # phi_M glia
    assert np.abs(got[5:] / np.array([3.3236743806989133, 124.1417480555213, 102.74303871910962,
                                      100.70633053050948, 12.8382384023078, 12.396954489036373]) - 1.0).max() < 1e-8
    return got
