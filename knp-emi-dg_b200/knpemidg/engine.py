"""Engine: the per-time-step loop of the reference's Solver on one device context.

It owns the order of operations of `Solver.solve_system_active` /
`solve_for_time_step` (src/knpemidg/solver.py:1072-1127, 794-847):

    ODE phase   per membrane model: phi_M, Nernst potentials and update_ode links
                -> parameter columns, step, V and I_ch back         (:1077-1113)
    EMI         assemble A, B, L; CG + AMG(B)                        (:470-529)
    KNP         assemble A_k, L_k; GMRES + AMG(A_k)                  (:723-789)
    updates     phi_M, Nernst, eliminated ion                        (:809-842)

All arithmetic happens in libknpemi.so; this class only sequences C-ABI calls
and keeps the small amount of host state (step counter, times, statistics).
`knpemidg.Solver` (the reference-facing API) is a thin layer over it.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from ._lib import F_C, F_CN, F_ICH, F_NERNST, F_PHI, F_PHIM  # noqa: F401


class MembraneHandle:
    """One membrane tag = one ODE model instance on the device."""

    def __init__(self, engine, tag, module, rows, model_id, ns, npar):
        self.engine, self.tag, self.module = engine, tag, module
        self.rows = rows
        self.model_id, self.ns, self.np = model_id, ns, npar
        self.time = 0.0
        self.handle = None


def _simplex_rule(dim, n):
    """collapsed tensor Gauss rule on the reference simplex (barycentric points, weights
    summing to 1), exact to degree 2n-1-(dim-1)"""
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    if dim == 2:
        U, V = np.meshgrid(x, x, indexing="ij")
        WU, WV = np.meshgrid(w, w, indexing="ij")
        l1, l2 = U.ravel(), (V * (1 - U)).ravel()
        return np.column_stack([1 - l1 - l2, l1, l2]), (WU * WV * (1 - U)).ravel() * 2.0
    U, V, W = np.meshgrid(x, x, x, indexing="ij")
    WU, WV, WW = np.meshgrid(w, w, w, indexing="ij")
    l1, l2, l3 = U.ravel(), (V * (1 - U)).ravel(), (W * (1 - U) * (1 - V)).ravel()
    return (np.column_stack([1 - l1 - l2 - l3, l1, l2, l3]),
            (WU * WV * WW * (1 - U) ** 2 * (1 - V)).ravel() * 6.0)


def find_compiled_model(lib, ode):
    """name of the model compiled into `lib` that implements the ODE module `ode`: same table
    sizes, same defaults and the same right-hand side on probe inputs (two different example
    directories of the reference ship different modules called mm_hh); None if there is none"""
    import importlib
    from .models import BUNDLED
    s0 = np.asarray(ode.init_state_values(), dtype=float)
    p0 = np.asarray(ode.init_parameter_values(), dtype=float)
    f = getattr(ode, "rhs_numba", None)
    user_rhs = None
    for attr in ("py_func", "_pyfunc"):
        if f is not None and hasattr(f, attr):
            user_rhs = getattr(f, attr)
    rng = np.random.default_rng(7)
    probes = [(0.01 * i, s0 * (1 + 0.1 * rng.uniform(-1, 1, s0.size)), p0 + 0.1 * rng.uniform(0.5, 1, p0.size))
              for i in range(3)]
    compiled = lib.models()
    for name in BUNDLED:
        if name not in compiled:
            continue
        mod = importlib.import_module("knpemidg.models." + name)
        if mod is ode:
            return name
        if len(mod.init_state_values()) != s0.size or len(mod.init_parameter_values()) != p0.size:
            continue
        if not (np.array_equal(mod.init_state_values(), s0) and np.array_equal(mod.init_parameter_values(), p0)):
            continue
        if user_rhs is None:
            return name
        ok = True
        for t, y, p in probes:
            d1, d2, p1, p2 = np.zeros_like(y), np.zeros_like(y), p.copy(), p.copy()
            user_rhs(t, y, d1, p1)
            mod.rhs_numba.py_func(t, y, d2, p2)
            ok &= np.allclose(d1, d2, rtol=1e-12, atol=0) and np.allclose(p1, p2, rtol=1e-12, atol=0)
        if ok:
            return name
    return None


class Engine:
    def __init__(self, mesh, cell_tags, facet_tags, *, F, R, T, C_M, C_phi, dt, z, D_sub, rho_sub=None,
                 membrane_tags=(), degree=1, splitting=True, mms=False, C_sub=None, device=0, lib=None,
                 transport=None, part=None):
        """`transport` (knpemidg.partition.TorchTransport of a torch.distributed job with
        more than one rank): this engine then holds ONE part of the cell-partitioned mesh
        (`part[cell]` = owning rank, default partition_cells) and every rank runs the same
        sequence of calls; `mesh`, the tags and all set_* inputs stay GLOBAL arrays."""
        mesh.init_topology()
        self.global_mesh = mesh
        self.transport = transport if (transport is not None and transport.world > 1) else None
        self.local = None
        cell_tags = np.asarray(cell_tags)
        facet_tags = np.asarray(facet_tags)
        self.tags = np.unique(cell_tags)
        self.nc_global = mesh.cells.shape[0]
        self._global_cell_tags = cell_tags
        self.global_membrane_facets = np.flatnonzero(
            (mesh.facet_cells[:, 1] >= 0) & np.isin(facet_tags, [int(t) for t in membrane_tags]))
        if self.transport is not None:
            from . import partition
            if part is None:
                part = partition.partition_cells(mesh, self.transport.world)
            self.local = partition.LocalPart(mesh, cell_tags, facet_tags, part, self.transport.rank)
            mesh, cell_tags, facet_tags = self.local.mesh, self.local.cell_tags, self.local.facet_tags
        self.mesh = mesh
        self.ctx = _lib.Context(device, lib)
        ctx = self.ctx
        region = np.searchsorted(self.tags, cell_tags).astype(np.int32)
        ctx.set_mesh(mesh.coords, mesh.cells, region, mesh.facet_cells, np.asarray(facet_tags),
                     tuple(int(t) for t in membrane_tags))
        if self.local is not None:
            ctx.set_dist(**self.local.dist_args())
            self.transport.attach(ctx)
        self.d, self.nd, self.nc, self.n, self.nm = ctx.d, ctx.nd, ctx.nc, ctx.n, ctx.nm
        self.N = len(z)
        self.dt = float(dt)
        tau = 20.0 * self.d * degree                                   # solver.py:110-111
        ext = mesh.coords.max(axis=0) - mesh.coords.min(axis=0)
        Lp = float(ext.max())                                          # solver.py:383-391
        # The EMI preconditioner matrix B = A + kappa/Lp^2 (u, v) (solver.py:393-395) over-weights the
        # constants of compact intracellular regions when the box is much larger than the cells: on the
        # reference's EMIx mesh the kappa/Lp^2 |cell| of a cell is ~70x its membrane coupling C_phi |membrane|,
        # which leaves one outlying eigenvalue per cell in M^-1 A (a 16-iteration CG plateau).
        # KNP_EMI_LP_SCALE=10 divides the shift by 100: 47 -> 14 CG iterations there, same solution (B
        # only preconditions).  Default 1 = the reference's B; "auto" picks the scale from the geometry at
        # initialize() (see _auto_lp_scale).
        self._lp_mode = os.environ.get("KNP_EMI_LP_SCALE", "auto")
        if self._lp_mode != "auto":
            Lp *= float(self._lp_mode)

        def table(sub):
            return [float(sub[int(t)]) if int(t) in sub else 0.0 for t in self.tags]

        D = np.array([table(Dk) for Dk in D_sub])
        rho = np.zeros(len(self.tags)) if rho_sub is None else np.array(table(rho_sub))
        Cs = None if C_sub is None else np.array([table(Ck) for Ck in C_sub])
        self._params = dict(F=F, R=R, T=T, C_M=C_M, C_phi=C_phi, dt=dt, tau_emi=tau, tau_knp=tau, Lp=Lp, z=z,
                            D=D, rho=rho, C_sub=Cs, mms=mms)
        self.splitting = bool(splitting)
        ctx.set_params(splitting=splitting, **self._params)
        self.C_M = float(C_M)
        self.cell_tags = cell_tags
        self.mem = ctx.membrane_table()
        self.members = []
        self.user_models = {}            # ODE module -> model name in a library variant (_lib.variant_with)
        self.k = 0
        self.t = 0.0
        self.amg_ready = False
        self.stats = {"emi_niter": [], "knp_niter": []}
        self.rtol_emi, self.atol_emi = 1e-5, 1e-40
        self.rtol_knp, self.atol_knp = 1e-7, 1e-40
        self.max_it = 1000                                             # ksp_max_it (solver.py:429, 687)
        self.ode_rtol, self.ode_atol = 1e-8, 0.0                       # membrane.py:112
        self.phi_M_init_type = "constant"

    # -- initial data --------------------------------------------------------
    def set_concentrations_by_tag(self, c_init_sub):
        """c_init_sub[k]: {cell tag: value} for all N ions (solver.py:179-206, 'constant')."""
        for k, sub in enumerate(c_init_sub):
            vals = np.zeros(self.nc)
            for t in self.tags:
                vals[self.cell_tags == t] = float(sub[int(t)])
            self.ctx.set_field(F_C, k, np.repeat(vals, self.nd))

    def localize(self, nodal):
        """global per-dof array [nc_global * nd] -> this rank's cells (owned + ghost)"""
        if self.local is None:
            return nodal
        return np.asarray(nodal).reshape(self.nc_global, -1)[self.local.l2g]

    def set_concentration(self, k, nodal):
        self.ctx.set_field(F_C, k, self.localize(nodal))

    def membrane_midpoints(self):
        return self.mesh.facet_midpoints()[self.mem["facet"]]

    def set_splitting(self, splitting):
        """splitting scheme on/off (solver.py:958, 1042) without touching the fields"""
        self.splitting = bool(splitting)
        self.ctx.set_params(splitting=splitting, **self._params)

    def set_source(self, k, f):
        """f_source of solved ion k as a load vector: int_{Omega_0} f lambda_i dx over the
        ECS cells (cell tag 0) only (solver.py:599).  f: number or callable(x)."""
        d, nd = self.d, self.nd
        ecs = np.flatnonzero(self.cell_tags == 0)
        X = self.mesh.coords[self.mesh.cells[ecs]]                       # [ne, nd, d]
        vol = self.mesh.cell_volume()[ecs]
        load = np.zeros((self.nc, nd))
        if callable(f):
            bq, wq = _simplex_rule(d, 3)                                   # degree-5 collapsed Gauss rule
            xq = np.einsum("qa,cak->cqk", bq, X)
            fv = np.array([[float(f(x)) for x in cell] for cell in xq])
            load[ecs] = np.einsum("q,c,cq,qi->ci", wq, vol, fv, bq)
        else:
            load[ecs] = (float(f) * vol / (d + 1))[:, None]
        self.ctx.set_field(_lib.F_LOAD_KNP, k, load)

    def resolve_model(self, ode):
        """name of the compiled model that implements the user's ODE module"""
        name = self.user_models.get(ode) or find_compiled_model(self.ctx.lib, ode)
        if name is None:
            raise _lib.KnpError(f"membrane model '{ode.__name__}' has no compiled counterpart in the library "
                                f"(available: {sorted(self.ctx.lib.models())}); add it to knpemidg/models and "
                                "rebuild, or let Solver.setup_membrane_model build a library variant for it")
        return name

    # -- membrane models -----------------------------------------------------
    def add_membrane_model(self, tag, module, ion_names, stimulus=None, stimulus_locator=None,
                           links=(("K_e", 0, "plus"), ("Na_i", -1, "minus"))):
        """setup_membrane_model for one tag (solver.py:228-267).  `links`: the
        update_ode hook as data: (parameter name, ion index, side)."""
        lib = self.ctx.lib
        models = lib.models()
        # The compiled model is found by WHAT the module computes (table sizes, defaults, right-hand
        # side on probe inputs), never by its bare name: the reference ships different modules that
        # are all called mm_hh (examples/idealized-geometries, emix-simulations,
        # local-astrocyte-depolarization).
        try:
            name = self.resolve_model(module)
        except AttributeError:
            raise _lib.KnpError(f"membrane model '{module.__name__}' is not compiled into libknpemi.so and does not follow "
                                "the mm_*.py protocol (init_state_values, init_parameter_values, state_indices, "
                                f"parameter_indices); available: {sorted(models)}") from None
        mid, ns, npar = models[name]
        rows = np.flatnonzero(self.mem["tag"] == tag).astype(np.int32)   # ascending facet index
        m = MembraneHandle(self, tag, module, rows, mid, ns, npar)
        s0 = np.asarray(module.init_state_values(), dtype=float)
        p0 = np.asarray(module.init_parameter_values(), dtype=float)
        if s0.size != ns or p0.size != npar:
            raise _lib.KnpError(f"membrane model '{module.__name__}': tables of {s0.size} states / {p0.size} parameters, "
                                f"the compiled model '{name}' has {ns} / {npar}")
        states = np.tile(s0, (len(rows), 1))
        params = np.tile(p0, (len(rows), 1))
        params[:, module.parameter_indices("Cm")] = self.C_M              # solver.py:248
        m.handle = self.ctx.membrane_register(mid, rows, states, params)
        ich = [module.parameter_indices("I_ch_" + nme) for nme in ion_names]
        self.ctx.membrane_outputs(m.handle, module.state_indices("V"), ich)
        for k, nme in enumerate(ion_names):                               # solver.py:1097-1098
            self.ctx.membrane_link(m.handle, module.parameter_indices("E_" + nme), 0, F_NERNST, k)
        for pname, ion, side in links:                                    # update_ode hook
            self.ctx.membrane_link(m.handle, module.parameter_indices(pname), 1, F_C, ion % self.N,
                                   0 if side == "plus" else 1)
        if stimulus:
            mid_pts = self.membrane_midpoints()[rows]
            if stimulus_locator is None:
                mask = np.ones(len(rows), dtype=np.uint8)
            else:
                mask = np.fromiter((bool(stimulus_locator(x)) for x in mid_pts), dtype=np.uint8,
                                   count=len(rows))                       # membrane.py:92
            cols = [module.parameter_indices(key) for key in stimulus]
            self.ctx.membrane_stimulus(m.handle, mask, cols, [float(v) for v in stimulus.values()])
        # I_ch_k functions start from the ODE parameter table (solver.py:251-259)
        for k in range(len(ion_names)):
            cur = self.ctx.get_field(F_ICH, k)
            cur[rows] = params[:, ich[k]]
            self.ctx.set_field(F_ICH, k, cur)
        self.members.append(m)
        return m

    # -- one step --------------------------------------------------------------
    def initialize(self, pc=1, amg_theta=0.08):
        """What the reference does before the loop: initial Nernst potentials
        (setup_varform_emi, solver.py:299) and the first assembly (setup_solver_*,
        :452-453, 710); here the first assembly also fixes the AMG plan."""
        if self._lp_mode == "auto":
            scale = self._auto_lp_scale()
            if scale > 1.0:
                self._params["Lp"] *= scale
                self.ctx.set_params(splitting=self.splitting, **self._params)
        self.ctx.post_step(_lib.POST_NERNST)
        self.ctx.assemble_emi()
        if pc == 1:
            self.ctx.amg_setup(theta=amg_theta)
            self.amg_ready = True
        self.ctx.solver_options(pc=pc)
        self._initialized = True

    def _auto_lp_scale(self):
        """KNP_EMI_LP_SCALE=auto: the factor on Lp that brings the mass shift kappa |region| / Lp^2 of the
        preconditioner matrix B down to the membrane coupling C_phi |membrane| of every intracellular
        region (1 where it already is below, e.g. the long thin axons of the bundle: B then is the
        reference's).  Whole-mesh quantities from the initial state; setup-time host code."""
        mesh, tags = self.global_mesh, self._global_cell_tags
        vol = mesh.cell_volume()
        X = mesh.coords[mesh.facet_verts[self.global_membrane_facets]]
        if mesh.gdim == 3:
            area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
        else:
            area = np.linalg.norm(X[:, 1] - X[:, 0], axis=1)
        fc = mesh.facet_cells[self.global_membrane_facets]
        side = np.maximum(tags[fc[:, 0]], tags[fc[:, 1]])                # the intracellular side
        P = self._params
        psi = P["F"] / (P["R"] * P["T"])
        c = [self.concentration(k, gather=True).mean(axis=1) for k in range(self.N)]
        worst = 0.0
        for ti, t in enumerate(self.tags):
            sel = side == t
            if not sel.any():
                continue
            cells = tags == t
            kappa = P["F"] * psi * sum(P["z"][k] ** 2 * P["D"][k][ti] * float(c[k][cells].mean()) for k in range(self.N))
            worst = max(worst, kappa * vol[cells].sum() / (P["Lp"] ** 2 * P["C_phi"] * area[sel].sum()))
        return float(np.sqrt(worst)) if worst > 1.0 else 1.0

    def ode_phase(self):
        for m in self.members:
            set_v = not (self.phi_M_init_type == "constant" and self.k == 0)   # solver.py:1086-1094
            self.ctx.ode_step(m.handle, m.time, self.dt, self.ode_rtol, self.ode_atol, set_v)
            m.time += self.dt                                                  # membrane.py:115

    def pde_phase(self):
        ctx = self.ctx
        ctx.assemble_emi()
        it, _ = ctx.solve_emi(self.rtol_emi, self.atol_emi, self.max_it)
        self.stats["emi_niter"].append(it)
        ctx.assemble_knp()
        it, _ = ctx.solve_knp(self.rtol_knp, self.atol_knp, self.max_it)
        self.stats["knp_niter"].append(it)
        ctx.post_step(_lib.POST_ALL)
        self.t += self.dt

    def pde_phase_picard(self, tol=1.0e-4, max_iter=25):
        """solve_for_time_step_picard (solver.py:850-927; present but not called in the reference):
        EMI and KNP are re-solved with the conductivities / fractions of the latest Picard iterate
        c_prev_k until max |c_prev_k - c| <= tol, the time derivative always refers to c_prev_n.
        Returns the number of Picard iterations."""
        ctx = self.ctx
        nsolved = self.N - 1
        c_n = [ctx.get_field(F_C, k) for k in range(nsolved)]
        for k in range(nsolved):
            ctx.set_field(F_CN, k, c_n[k])                 # c_prev_n stays fixed during the iteration
        prev = c_n
        it = 0
        eps = 2.0 * tol
        while eps > tol:
            it += 1
            ctx.assemble_emi()
            n_emi, _ = ctx.solve_emi(self.rtol_emi, self.atol_emi, self.max_it)
            ctx.assemble_knp()
            n_knp, _ = ctx.solve_knp(self.rtol_knp, self.atol_knp, self.max_it)
            self.stats["emi_niter"].append(n_emi)
            self.stats["knp_niter"].append(n_knp)
            cur = [ctx.get_field(F_C, k) for k in range(nsolved)]
            eps = max(float(np.abs(a - b).max()) for a, b in zip(prev, cur))   # inf-norm, solver.py:883-884
            prev = cur
            ctx.post_step(_lib.POST_ELIMINATED | _lib.POST_NERNST)              # solver.py:892-913
            if it > max_iter:
                raise _lib.KnpError("Picard solver diverged (solver.py:916-918)")
        for k in range(nsolved):
            ctx.set_field(F_CN, k, prev[k])                                     # c_prev_n <- c_prev_k (:921)
        ctx.post_step(_lib.POST_PHIM)                                           # :924-925
        self.t += self.dt
        self.picard_iterations = it
        return it

    def step(self):
        if not getattr(self, "_initialized", False):
            self.initialize()
        if self.members:
            self.ode_phase()
        self.pde_phase()
        self.k += 1

    # -- observables ---------------------------------------------------------
    # Single part: the arrays of the whole mesh.  Partitioned: this rank's cells (owned first,
    # then ghosts) unless gather=True, which assembles the global array on every rank
    # (collective; tests and output only, never inside the time loop).
    def _cells(self, local, gather):
        local = local.reshape(self.nc, self.nd)
        if self.local is None or not gather:
            return local
        no = self.local.nc_owned
        return self.transport.gather_owned(local[:no], self.local.owned_global_cells(), self.nc_global)

    def phi(self, gather=False):
        return self._cells(self.ctx.get_field(F_PHI), gather)

    def concentration(self, k, gather=False):
        return self._cells(self.ctx.get_field(F_C, k), gather)

    def phi_M(self, gather=False):
        v = self.ctx.get_field(F_PHIM)
        if self.local is None or not gather:
            return v
        mine = self.mem["cell_i"] < self.local.nc_owned           # the copy that counts: ICS cell owner
        gfacet = self.local.facets[self.mem["facet"]]
        rows = np.searchsorted(self.global_membrane_facets, gfacet)
        return self.transport.gather_owned(v[mine], rows[mine], self.global_membrane_facets.size)

    def dofs(self):
        """V_emi.dim() + V_knp.dim() (solver.py:1163-1164) of the whole mesh."""
        return self.N * self.nd * self.nc_global
