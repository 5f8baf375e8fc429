"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the reference's time loop:
  solve_system_active    src/knpemidg/solver.py:1072-1127 (ODE phase :1077-1113)
  solve_system_passive   src/knpemidg/solver.py:985-1000
  solve_for_time_step    src/knpemidg/solver.py:794-847
  setup_membrane_model   src/knpemidg/solver.py:228-267
with scipy sparse direct solves (the reference's MMS tests use MUMPS LU,
tests/run_MMS_space.py:202, 208) or scipy Krylov solvers with the reference's
tolerances (solver.py:425-444, 684-701).

Pinned by the reference's own loop executed on oracle/refexec: 40 steps of the 2D neuron and 16 steps
of the astrocyte problem (tests/golden/ref_run_*.npz), membrane-potential traces within 1e-6.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import forms, ode, amg


def solve_singular_direct(A, b):
    """Direct solve of the pure-Neumann EMI system: the reference removes the
    constant null space from b and lets MUMPS handle the null pivot
    (solver.py:415-416, 489-490).  Here: bordered system, zero-mean solution."""
    n = A.shape[0]
    one = np.ones((n, 1))
    K = sp.bmat([[A, sp.csr_matrix(one)], [sp.csr_matrix(one.T), None]], format="csc")
    rhs = np.concatenate([b - b.mean(), [0.0]])
    return spla.spsolve(K, rhs)[:n]


class OracleSolver:
    def __init__(self, P, c_init, models=None, stimulus=None, stimulus_locator=None,
                 phi_M_init_type="constant", mms=None, splitting=True, f_source=None,
                 ode_links=None, direct=True, rtol_emi=1e-5, rtol_knp=1e-7,
                 current_convention="end_state", ion_names=None):
        self.P = P
        self.mms = mms
        self.splitting = splitting
        self.f_source = f_source
        self.direct = direct
        self.rtol_emi, self.rtol_knp = rtol_emi, rtol_knp
        self.current_convention = current_convention
        self.ion_names = ion_names or [f"ion{k}" for k in range(P.N)]
        c_init = np.asarray(c_init, dtype=float)
        self.c = c_init[:P.N_ions].copy()            # c_prev_k == c_prev_n between steps
        self.c_elim = c_init[-1].copy()
        self.phi = np.zeros((P.nc, P.nd))
        self.phi_M = np.zeros(P.nm)                  # phi_M_prev_PDE starts as zero (solver.py:211-214)
        self.I_ch = np.zeros((P.N, P.nm))
        self.E = np.zeros((P.N, P.nm))
        self.phi_M_init_type = phi_M_init_type
        self.t = 0.0
        self.k = 0
        self.niter = {"emi": [], "knp": []}
        self._plan = None
        self.timers = {"emi_assemble": 0.0, "emi_solve": 0.0, "knp_assemble": 0.0, "knp_solve": 0.0, "ode": 0.0}
        # default update_ode hook of the idealized examples (run_2D.py:38-50)
        self.ode_links = ode_links if ode_links is not None else \
            [("K_e", 0, "plus"), ("Na_i", P.N - 1, "minus")]
        self.models = []
        if models:
            mid = P.mesh.facet_midpoints()[P.mem_facets]
            for tag, module in models.items():
                rows = np.flatnonzero(P.mem_tag == tag)
                m = {"tag": tag, "module": module, "rows": rows,
                     # (reshape: a tag without facets gives empty tables, not 1-D arrays)
                     "states": np.array([module.init_state_values() for _ in rows]).reshape(
                         len(rows), len(module.init_state_values())),
                     "parameters": np.array([module.init_parameter_values() for _ in rows]).reshape(
                         len(rows), len(module.init_parameter_values())),
                     "time": 0.0}
                m["parameters"][:, module.parameter_indices("Cm")] = P.C_M      # solver.py:248
                if stimulus_locator is None:
                    m["mask"] = np.ones(len(rows), dtype=bool)
                else:
                    m["mask"] = np.fromiter((bool(stimulus_locator(x)) for x in mid[rows]),
                                            dtype=bool, count=len(rows))
                for k, name in enumerate(self.ion_names):                       # solver.py:253-259
                    self.I_ch[k, rows] = m["parameters"][:, module.parameter_indices("I_ch_" + name)]
                self.models.append(m)
        self.stimulus = stimulus
        self._update_nernst()

    # -- pieces ---------------------------------------------------------------
    def c_all(self):
        return np.concatenate([self.c, self.c_elim[None]], axis=0)

    def _update_nernst(self):
        if self.P.nm == 0 or self.mms is not None:   # MMS: E unused, c may vanish
            return
        ca = self.c_all()
        for k in range(self.P.N):
            self.E[k] = forms.nernst(self.P, ca[k], self.P.z[k])

    def ode_phase(self):
        P = self.P
        ca = self.c_all()
        for m in self.models:
            mod, rows = m["module"], m["rows"]
            if not (self.phi_M_init_type == "constant" and self.k == 0):        # solver.py:1086-1094
                m["states"][:, mod.state_indices("V")] = self.phi_M[rows]
            for k, name in enumerate(self.ion_names):                           # solver.py:1097-1098
                m["parameters"][:, mod.parameter_indices("E_" + name)] = self.E[k, rows]
            for pname, ion, side in self.ode_links:                             # update_ode hook
                tr = forms.facet_mean_trace(P, ca[ion], side)
                m["parameters"][:, mod.parameter_indices(pname)] = tr[rows]
            ode.step_rows(mod, m["states"], m["parameters"], m["time"], P.dt,
                          stim_mask=m["mask"], stimulus=self.stimulus,
                          current_convention=self.current_convention)
            m["time"] += P.dt
            self.phi_M[rows] = m["states"][:, mod.state_indices("V")]           # solver.py:1108
            for k, name in enumerate(self.ion_names):                           # solver.py:1111-1113
                self.I_ch[k, rows] = m["parameters"][:, mod.parameter_indices("I_ch_" + name)]

    def assemble_emi(self):
        return forms.assemble_emi(self.P, self.c_all(), self.phi_M, self.I_ch,
                                  splitting=self.splitting, mms=self.mms)

    def assemble_knp(self):
        fs = None
        if self.f_source is not None:
            fs = [(lambda x, f=f: f(x, self.t)) if callable(f) else f for f in self.f_source]
        return forms.assemble_knp(self.P, self.c_all(), self.c, self.phi, self.phi_M, self.I_ch,
                                  splitting=self.splitting, f_source=fs, mms=self.mms)

    def solve_emi(self):
        A, B, b = self.assemble_emi()
        if self.direct:
            x = solve_singular_direct(A, b)
        else:
            if self._plan is None:                      # aggregates are kept across steps
                self._plan = amg.Plan(self.P, B)
            H = amg.Hierarchy(self._plan, B)
            M = spla.LinearOperator(A.shape, matvec=H.apply)
            it = [0]
            x, info = spla.cg(A, b, x0=self.phi.ravel().copy(), rtol=self.rtol_emi, atol=0.0,
                              maxiter=1000, M=M, callback=lambda _: it.__setitem__(0, it[0] + 1))
            assert info == 0
            self.niter["emi"].append(it[0])
        self.phi = x.reshape(self.P.nc, self.P.nd)

    def solve_knp(self):
        As, bs = self.assemble_knp()
        new = np.empty_like(self.c)
        for k, (A, b) in enumerate(zip(As, bs)):
            if self.direct:
                x = spla.spsolve(A.tocsc(), b)
            else:
                if self._plan is None:
                    self._plan = amg.Plan(self.P, A)
                H = amg.Hierarchy(self._plan, A)
                M = spla.LinearOperator(A.shape, matvec=H.apply)
                it = [0]
                x, info = spla.gmres(A, b, x0=self.c[k].ravel().copy(), rtol=self.rtol_knp, atol=0.0,
                                     restart=30, maxiter=1000, M=M, callback_type="pr_norm",
                                     callback=lambda _: it.__setitem__(0, it[0] + 1))
                assert info == 0
                self.niter["knp"].append(it[0])
            new[k] = x.reshape(self.P.nc, self.P.nd)
        self.c = new

    def pde_phase(self):
        """solve_for_time_step (solver.py:794-847)."""
        if self.mms is not None:
            self.mms.t = self.t          # sources are evaluated before t.assign (solver.py:845)
        self.solve_emi()
        self.solve_knp()
        if self.P.nm:
            self.phi_M = forms.membrane_potential(self.P, self.phi)
        self.c_elim = forms.eliminated_concentration(self.P, self.c)
        self._update_nernst()
        self.t += self.P.dt

    def step(self):
        if self.models:
            self.ode_phase()
        self.pde_phase()
        self.k += 1

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()
        return self
