"""Solver with the reference's interface (src/knpemidg/solver.py:62-1298) on top of the
B200 library.  The run scripts of the reference (examples/*/run_*.py) drive exactly this
surface: constructor, setup_domain / setup_parameters / setup_FEM_spaces /
setup_membrane_model, solve_system_active / solve_system_passive and the update_ode hook.

What differs, deliberately:
  * `mesh`, `subdomains`, `surfaces` are knpemidg.mesh.SimplexMesh / MeshFunction objects
    (or dolfin objects, adapted once through knpemidg.dolfin_adapter when dolfin exists);
  * fields are device handles (knpemidg.frontend), not dolfin Functions;
  * linear algebra, assembly and the ODE step run in libknpemi.so; the PETSc option
    `threshold_*` (BoomerAMG's classical strength threshold, run_3D.py:174) has no counterpart in
    the aggregation hierarchy used here (fixed strength parameter 0.08): a value other than None
    is reported once on stderr and otherwise ignored - it only ever affected iteration counts;
  * `save_fields` writes results.h5 in the reference's HDF5 layout (solver.py:1214-1242) through
    knpemidg.h5lite (no libhdf5 here); solver statistics keep the reference's text format.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import _lib
from .engine import Engine
from .frontend import (CellField, Constant, FacetField, FunctionSpace, InterfaceNormal, MixedCellField,
                       as_float)
from .membrane import MembraneModel
from .utils import interface_normal


def _adapt_mesh(mesh, subdomains, surfaces):
    from .mesh import SimplexMesh
    if isinstance(mesh, SimplexMesh):
        return mesh, np.asarray(subdomains.array()), np.asarray(surfaces.array())
    from .dolfin_adapter import from_dolfin          # dolfin objects (only where dolfin exists)
    return from_dolfin(mesh, subdomains, surfaces)


class Solver:
    def __init__(self, params, ion_list, degree_emi=1, degree_knp=1, mms=None, sf=1, device=0, lib=None):
        if degree_emi != 1 or degree_knp != 1:
            raise NotImplementedError("the B200 path implements DG-P1 (the degree every example uses)")
        self.ion_list = ion_list
        self.N_ions = len(ion_list[:-1])            # solver.py:69
        self.degree_emi, self.degree_knp = degree_emi, degree_knp
        self.mms = mms
        self.params = params
        self.sf = sf
        self._device, self._lib = device, lib
        self.ode_solve_timer = self.emi_solve_timer = self.knp_solve_timer = 0
        self.emi_ass_timer = self.knp_ass_timer = 0
        self.engine = None
        self.mem_models = []

    # -- setup ---------------------------------------------------------------------
    def setup_domain(self, mesh, subdomains, surfaces):
        self.mesh, self._cell_tags, self._facet_tags = _adapt_mesh(mesh, subdomains, surfaces)
        self.subdomains, self.surfaces = subdomains, surfaces
        self.gdim = self.mesh.gdim
        self.tau_emi = Constant(20 * self.gdim * self.degree_emi)     # solver.py:110-111
        self.tau_knp = Constant(20 * self.gdim * self.degree_knp)
        self.n_g = interface_normal(subdomains, mesh)
        if self.mms is not None:
            self.lm_tags = [1, 2, 3, 4]                                # solver.py:118-119

    def setup_parameters(self):
        p = self.params
        self.C_phi, self.C_M, self.dt = Constant(as_float(p.C_phi)), Constant(as_float(p.C_M)), Constant(as_float(p.dt))
        self.F, self.R = Constant(as_float(p.F)), Constant(as_float(p.R))
        self.temperature = Constant(as_float(p.temperature))
        self.psi = float(self.F) / (float(self.R) * float(self.temperature))
        self.phi_M_init_type = p.phi_M_init_type
        for ion in self.ion_list:
            ion["D"] = {int(k): as_float(v) for k, v in ion["D_sub"].items()}      # make_global, :1244-1258
            if self.mms is not None:
                ion["C"] = {int(k): as_float(v) for k, v in ion["C_sub"].items()}
        self.rho = {int(k): as_float(v) for k, v in p.rho_sub.items()}

    def _create_engine(self, membrane_tags, splitting):
        C_sub = [ion["C"] for ion in self.ion_list[:-1]] if self.mms is not None else None
        eng = Engine(self.mesh, self._cell_tags, self._facet_tags, F=float(self.F), R=float(self.R),
                     T=float(self.temperature), C_M=float(self.C_M), C_phi=float(self.C_phi), dt=float(self.dt),
                     z=[as_float(ion["z"]) for ion in self.ion_list], D_sub=[ion["D"] for ion in self.ion_list],
                     rho_sub=self.rho, membrane_tags=membrane_tags, degree=self.degree_emi, splitting=splitting,
                     mms=self.mms is not None, C_sub=C_sub, device=self._device, lib=self._lib)
        eng.phi_M_init_type = self.phi_M_init_type
        return eng

    def setup_FEM_spaces(self):
        """The device context needs the membrane tags, which the reference only learns in
        setup_membrane_model; the engine is therefore created lazily (first of
        setup_membrane_model / solve_system_*).  Initial data are recorded here."""
        self._c_init = []
        for ion in self.ion_list:
            kind = ion["c_init_sub_type"]
            if kind == "constant":
                self._c_init.append(("constant", {int(k): as_float(v) for k, v in ion["c_init_sub"].items()}))
            elif kind == "expression":
                self._c_init.append(("expression", ion["c_init_sub"]))
            elif kind == "function":
                self._c_init.append(("function", ion["c_init_sub"]))
            else:
                print(f'Type of initial condition "{kind}" not recognized - please spesify whether initial '
                      'condition is "constant", "expression" or "function"')
                sys.exit(0)

    def _ensure_engine(self, membrane_tags, splitting):
        if self.engine is not None:
            return
        eng = self.engine = self._create_engine(membrane_tags, splitting)
        X = self.mesh.coords[self.mesh.cells]                      # [nc, nd, d] nodal coordinates
        for k, (kind, data) in enumerate(self._c_init):
            if kind == "constant":
                vals = np.zeros(eng.nc)
                for t in eng.tags:
                    vals[eng.cell_tags == t] = data[int(t)]
                nodal = np.repeat(vals, eng.nd)
            elif kind == "expression":                             # {tag: callable(x) -> value}
                nodal = np.zeros((eng.nc, eng.nd))
                for t in eng.tags:
                    m = eng.cell_tags == t
                    f = data[int(t)]
                    nodal[m] = np.array([[f(x) for x in cell] for cell in X[m]]) if callable(f) else as_float(f)
                nodal = nodal.ravel()
            else:
                nodal = data.vector().get_local() if hasattr(data, "vector") else np.asarray(data, dtype=float).ravel()
            eng.set_concentration(k, nodal)
        self.V_emi = FunctionSpace(eng, "DG1", self.mesh)
        self.V_knp = FunctionSpace(eng, "DG1", self.mesh)
        self.Q = FunctionSpace(eng, "DLT0", self.mesh)
        self.phi = CellField(eng, _lib.F_PHI, 0, "phi")
        self.c = MixedCellField(eng, _lib.F_C)
        self.c_prev_k = MixedCellField(eng, _lib.F_C)
        self.c_prev_n = MixedCellField(eng, _lib.F_CN)
        self.ion_list[-1]["c"] = CellField(eng, _lib.F_C, eng.N - 1, "c_elim")
        self.phi_M_prev_PDE = FacetField(eng, _lib.F_PHIM)
        for k, ion in enumerate(self.ion_list):
            ion["E"] = FacetField(eng, _lib.F_NERNST, k)

    def setup_membrane_model(self, stim_params, odes):
        self.stimulus = stim_params.stimulus
        self.stimulus_locator = stim_params.stimulus_locator
        user_names = {}
        if self.engine is None:
            # ODE modules without a compiled counterpart: build (once, cached) a library variant that
            # carries their right-hand sides as device code, before the device context is created
            from .engine import find_compiled_model
            lib = self._lib or _lib.get()
            missing = [ode for ode in odes.values() if find_compiled_model(lib, ode) is None]
            if missing:
                print(f"knpemidg: compiling {len(missing)} user membrane model(s) into a library variant ...")
                self._lib, user_names = _lib.variant_with(lib, missing)
        self._ensure_engine(tuple(int(t) for t in odes), splitting=True)
        eng = self.engine
        eng.user_models.update(user_names)
        self.mem_models = []
        for tag, ode in odes.items():
            ode_model = MembraneModel(ode, facet_f=self.surfaces, tag=int(tag), V=self.Q)
            ode_model.set_parameter_values({"Cm": lambda x: as_float(self.params.C_M)})       # solver.py:248
            names = [ion["name"] for ion in self.ion_list]
            ich = [ode.parameter_indices("I_ch_" + nme) for nme in names]
            eng.ctx.membrane_outputs(ode_model.handle, ode.state_indices("V"), ich)
            I_ch_k = {}
            table = ode_model.parameters
            for k, nme in enumerate(names):                                                   # solver.py:251-259
                cur = eng.ctx.get_field(_lib.F_ICH, k)
                cur[ode_model.indices] = table[:, ich[k]]
                eng.ctx.set_field(_lib.F_ICH, k, cur)
                I_ch_k[nme] = FacetField(eng, _lib.F_ICH, k)
            self.mem_models.append({"ode": ode_model, "I_ch_k": I_ch_k})

    # -- time stepping -------------------------------------------------------------
    def _configure(self, solver_params, splitting):
        self.solver_params = solver_params
        self.direct_emi = solver_params.direct_emi
        self.direct_knp = solver_params.direct_knp
        # a "direct" solve is emulated by iterating to (near) machine precision
        self.rtol_emi = 1e-11 if self.direct_emi else solver_params.rtol_emi
        self.atol_emi = 1e-300 if self.direct_emi else solver_params.atol_emi
        self.rtol_knp = 1e-13 if self.direct_knp else solver_params.rtol_knp
        self.atol_knp = 1e-300 if self.direct_knp else solver_params.atol_knp
        self.splitting_scheme = splitting
        for name in ("threshold_emi", "threshold_knp"):
            if getattr(solver_params, name, None) is not None and not getattr(Solver, "_threshold_noted", False):
                print(f"knpemidg (B200): solver_params.{name} = {getattr(solver_params, name)} is a BoomerAMG option; the "
                      "aggregation AMG of libknpemi.so has no such parameter, it is ignored", file=sys.stderr)
                Solver._threshold_noted = True
        self._ensure_engine((1, 2, 3, 4) if self.mms is not None else (), splitting)
        eng = self.engine
        eng.rtol_emi, eng.atol_emi = self.rtol_emi, self.atol_emi
        eng.rtol_knp, eng.atol_knp = self.rtol_knp, self.atol_knp
        eng.max_it = 50000 if (self.direct_emi or self.direct_knp) else 1000
        if bool(eng.splitting) != bool(splitting):
            eng.set_splitting(splitting)
        self._update_loads()
        eng.initialize(pc=1)

    def _update_loads(self):
        """f_source (solver.py:599) and MMS data (:365-374, 645-657) as load vectors,
        evaluated at the OLD time (t.assign comes last, :845)."""
        eng = self.engine
        if self.mms is not None:
            loads = self.mms
            if not hasattr(loads, "load_emi"):
                # the reference's own MMSData (tests/mms_space.py, mms_time.py): symbolic sources
                if getattr(self, "_mms_loads", None) is None:
                    from .mms_loads import ReferenceMMSLoads
                    self._mms_loads = ReferenceMMSLoads(self)
                loads = self._mms_loads
            eng.ctx.set_field(_lib.F_LOAD_EMI, 0, loads.load_emi(float(self._t)))
            for k in range(self.N_ions):
                eng.ctx.set_field(_lib.F_LOAD_KNP, k, loads.load_knp(k, float(self._t)))
            return
        for k, ion in enumerate(self.ion_list[:-1]):
            f = ion.get("f_source", None)
            if f is None:
                continue
            if callable(f):
                eng.set_source(k, lambda x, f=f: f(x, float(self._t)))
            elif as_float(f) != 0.0:
                eng.set_source(k, as_float(f))

    def update_ode(self, ode_model):
        """hook for PDE -> ODE updates specific to a membrane model (solver.py:1137-1144)"""
        raise NotImplementedError("Subclass must implement update_ode (e.g. K_e and Na_i traces)")

    def solve_for_time_step(self, k, t):
        """one global PDE step (solver.py:794-847)"""
        eng = self.engine
        self._t = t
        if self.mms is not None or any(callable(ion.get("f_source", None)) for ion in self.ion_list[:-1]):
            self._update_loads()
        eng.pde_phase()
        t.assign(float(t) + float(self.dt))

    def solve_for_time_step_picard(self, k, t):
        """one global PDE step with Picard iterations (solver.py:850-927; the reference keeps the
        call commented out at :996, 1123); note the reference advances t BEFORE the solves here"""
        eng = self.engine
        t.assign(float(t) + float(self.dt))
        self._t = t
        if self.mms is not None or any(callable(ion.get("f_source", None)) for ion in self.ion_list[:-1]):
            self._update_loads()
        try:
            eng.pde_phase_picard()
        except _lib.KnpError as e:
            if "Picard" in str(e):
                print("Picard solver diverged")
                sys.exit(2)
            raise

    def _ode_phase(self, k):
        eng = self.engine
        for mem_model in self.mem_models:                                  # solver.py:1077-1113
            ode_model = mem_model["ode"]
            if not (self.phi_M_init_type == "constant" and k == 0):
                ode_model.set_membrane_potential(self.phi_M_prev_PDE)
            for ion in self.ion_list:
                ode_model.set_parameter("E_" + ion["name"], ion["E"])
            self.update_ode(ode_model)
            ode_model.step_lsoda(dt=float(self.dt), stimulus=self.stimulus,
                                 stimulus_locator=self.stimulus_locator)
            # V -> phi_M_prev_PDE and I_ch_k -> PDE source functions are scattered by the kernel

    def _loop(self, Tstop, t, active, filename, save_fields, save_solver_stats):
        self.filename, self.save_fields, self.save_solver_stats = filename, save_fields, save_solver_stats
        if filename is None and (save_solver_stats or save_fields):
            print("Please specify filename when initiating the Solver.solve_system_* method")
            sys.exit(0)
        eng = self.engine
        if save_fields:
            self.init_h5_savefile(filename + "results.h5")
        if save_solver_stats:
            self.init_solver_stats(filename + "solver/")
        nsteps = int(round(Tstop / float(self.dt)))
        for k in range(nsteps):
            eng.ctx.timers(reset=True)
            if active:
                self._ode_phase(k)
            self.solve_for_time_step(k, t)
            tm = eng.ctx.timers()
            self.ode_solve_timer += tm["ode"]
            self.emi_ass_timer += tm["emi_assemble"]; self.emi_solve_timer += tm["emi_solve"]
            self.knp_ass_timer += tm["knp_assemble"]; self.knp_solve_timer += tm["knp_solve"]
            if save_solver_stats:
                self._write_stats(tm, eng.stats["emi_niter"][-1], eng.stats["knp_niter"][-1])
            if (k % self.sf) == 0 and save_fields:
                self.save_h5()
            eng.k += 1
        if save_fields:
            self.close_h5()
        if save_solver_stats:
            self.close_solver_stats()

    def solve_system_active(self, Tstop, t, solver_params, filename=None, save_fields=False,
                            save_solver_stats=False):
        """ODE step, then PDE step, per time step (solver.py:1014-1135)"""
        self._t = t
        self._configure(solver_params, splitting=True)
        self._loop(Tstop, t, True, filename, save_fields, save_solver_stats)

    def solve_system_passive(self, Tstop, t, solver_params, membrane_params=None, filename=None,
                             save_fields=False, save_solver_stats=False):
        """PDE steps only (solver.py:930-1011); returns (uh, c_elim)"""
        self._t = t
        self._configure(solver_params, splitting=False)
        self._loop(Tstop, t, False, filename, save_fields, save_solver_stats)
        uh = self.c.split() + (self.phi,)
        return uh, self.ion_list[-1]["c"]

    # -- output (solver.py:1146-1242) ---------------------------------------------
    def init_solver_stats(self, path):
        os.makedirs(path, exist_ok=True)
        r = getattr(self.solver_params, "resolution", 0)
        eng = self.engine
        dofs = eng.dofs()
        self._stat_files = {}
        for sysname in ("emi", "knp"):
            for what in ("solve", "assem", "niter"):
                f = open(os.path.join(path, f"{sysname}_{what}_{r}.txt"), "w")
                f.write(f"num cells: {eng.nc}\n")
                f.write(f"dofs: {dofs}\n")
                self._stat_files[(sysname, what)] = f

    def _write_stats(self, tm, it_emi, it_knp):
        sf = self._stat_files
        sf[("emi", "assem")].write("ass_time: %.4f \n" % tm["emi_assemble"])
        sf[("emi", "solve")].write("solve_time: %.4f \n" % tm["emi_solve"])
        sf[("emi", "niter")].write("niter: %d \n" % it_emi)
        sf[("knp", "assem")].write("ass_time: %.4f \n" % tm["knp_assemble"])
        sf[("knp", "solve")].write("solve_time: %.4f \n" % tm["knp_solve"])
        sf[("knp", "niter")].write("niter: %d \n" % it_knp)

    def close_solver_stats(self):
        for f in self._stat_files.values():
            f.close()

    def init_h5_savefile(self, filename):
        """HDF5 time series in the reference's layout (solver.py:1214-1227): /mesh, /subdomains,
        /surfaces, then /concentrations/vector_i, /elim_concentration/vector_i, /potential/vector_i
        for i = 0 (the state before the first step), 1, 2, ... - written incrementally through
        knpemidg.h5lite.Writer (no libhdf5 in this image).  dolfin's per-function dof-map datasets
        (cell_dofs, x_cell_dofs, cells) are written next to the vectors: dof = nd*cell + local
        vertex, the mixed concentration space stacks the ions."""
        from . import h5lite
        os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
        eng = self.engine
        mesh = self.mesh
        w = self._h5 = h5lite.Writer(filename)
        self.h5_idx = 0
        nc, nd = mesh.cells.shape
        w.write("/mesh/coordinates", mesh.coords)
        w.write("/mesh/topology", mesh.cells.astype(np.int64))
        w.write("/mesh/cell_indices", np.arange(nc, dtype=np.int64))
        for name, tags, dim in (("subdomains", self._cell_tags, mesh.gdim), ("surfaces", self._facet_tags, mesh.gdim - 1)):
            w.write(f"/{name}/coordinates", mesh.coords)
            w.write(f"/{name}/topology", (mesh.cells if dim == mesh.gdim else mesh.facet_verts).astype(np.int64))
            w.write(f"/{name}/values", np.asarray(tags).astype(np.int64))
        cells = np.arange(nc, dtype=np.int64)
        scalar = (nd * cells[:, None] + np.arange(nd)[None, :])
        for group, ncomp in (("concentrations", self.N_ions), ("elim_concentration", 1), ("potential", 1)):
            dofs = np.concatenate([k * nc * nd + scalar for k in range(ncomp)], axis=1)
            w.write(f"/{group}/cell_dofs", dofs.reshape(-1))
            w.write(f"/{group}/x_cell_dofs", np.arange(nc + 1, dtype=np.int64) * (nd * ncomp))
            w.write(f"/{group}/cells", cells)
        self._write_h5_fields()

    def _write_h5_fields(self):
        eng, w, i = self.engine, self._h5, self.h5_idx
        w.write(f"/concentrations/vector_{i}",
                np.concatenate([eng.concentration(k, gather=True).reshape(-1) for k in range(self.N_ions)]))
        w.write(f"/elim_concentration/vector_{i}", eng.concentration(eng.N - 1, gather=True).reshape(-1))
        w.write(f"/potential/vector_{i}", eng.phi(gather=True).reshape(-1))

    def save_h5(self):
        self.h5_idx += 1
        self._write_h5_fields()

    def close_h5(self):
        self._h5.close()
