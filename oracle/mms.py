"""ORACLE (test infrastructure only - never imported by the product path).

Manufactured-solution data of the reference's convergence tests, re-derived
with sympy (the reference builds the same expressions with UFL):

  space test: exact solutions  tests/mms_space.py:31-39, sources :64-74,
              interface data   :77-138, parameters tests/run_MMS_space.py:16-58
  time test:  exact solutions  tests/mms_time.py:28-43, parameters
              tests/run_MMS_time.py:16-58

and the MMS-only right-hand-side terms of the solver:
  EMI: src/knpemidg/solver.py:359, 365-366, 369, 372-374
  KNP: src/knpemidg/solver.py:645-646, 653-654, 657

Pinned by tests/golden/ref_mms.npz: the reference's own setup_mms + Solver(mms=...) +
solve_system_passive executed on oracle/refexec (the generator differentiates the UFL
expressions symbolically as UFL's apply_derivatives does): step-0 tensors of the r = 2 case,
the L2 errors the scripts print for r = 2..5 (space) and dt_0/4..dt_0/16 (time).  The reference's
scripts themselves assert nothing; the expected rates (space ~2, time ~1) are the second pin.
"""
import numpy as np
import sympy as sy

from . import quadrature as quad

X, Y, Tt = sy.symbols("x y t")

# normals from ICS (1) to ECS (2) on interface tags 1..4 (mms_space.py:77)
MMS_NORMALS = {1: (-1.0, 0.0), 2: (0.0, -1.0), 3: (1.0, 0.0), 4: (0.0, 1.0)}


class MMS:
    def __init__(self, kind="space", dt=1e-10, ufl_degree=None):
        """ufl_degree: integrate the MMS loads with the rules of that degree - 13 is what UFL estimates for the
        sin/cos data of mms_space.py (four trigonometric factors of estimated degree 3 each, times the test
        function; oracle/refexec/ufl_numeric.py) - instead of the fixed rules below"""
        self.kind = kind
        self.ufl_degree = ufl_degree
        self.t = 0.0
        # parameters (run_MMS_space.py:31-43 == run_MMS_time.py:47-56)
        self.D1 = [6.0, 3.0, 1.0]      # ICS  D_a1, D_b1, D_c1
        self.D2 = [5.0, 4.0, 2.0]      # ECS  D_a2, D_b2, D_c2
        self.C1 = [1.0, 2.0, 3.0]
        self.C2 = [2.0, 4.0, 2.0]
        self.z = [1.0, -1.0, 1.0]
        self.F = self.R = self.T = self.C_M = 1.0
        self.dt = dt
        self.C_phi = self.C_M / dt
        self.psi = self.F / (self.R * self.T)
        za, zb, zc = self.z
        pi = sy.pi
        if kind == "space":
            ka1 = 0.3 + 0.2 * sy.sin(2 * pi * X) * sy.sin(2 * pi * Y)
            kb1 = 0.9 + 0.3 * sy.cos(2 * pi * X) * sy.sin(2 * pi * Y)
            phi1 = sy.cos(2 * pi * X) * sy.cos(2 * pi * Y)
            ka2 = 0.3 + 0.2 * sy.cos(2 * pi * X) * sy.cos(2 * pi * Y)
            kb2 = 0.8 + 0.3 * sy.sin(2 * pi * X) * sy.cos(2 * pi * Y)
            phi2 = sy.sin(2 * pi * X) * sy.sin(2 * pi * Y)
        else:
            ka1 = 1 + (X + Y) + 0.2 * sy.cos(2 * pi * Tt)
            kb1 = 1 + (X + Y) + 0.3 * sy.cos(2 * pi * Tt)
            phi1 = (1 + X + Y) * (1 + Tt ** 2)
            ka2 = 1 + (X + Y) + 0.5 * sy.sin(2 * pi * Tt)
            kb2 = 1 + (X + Y) + 0.6 * sy.sin(2 * pi * Tt)
            phi2 = (1 + X - Y) * (1 + Tt ** 2)
        kc1 = -1 / zc * (za * ka1 + zb * kb1)
        kc2 = -1 / zc * (za * ka2 + zb * kb2)
        self.sol = {"c1": [ka1, kb1, kc1], "c2": [ka2, kb2, kc2], "phi1": phi1, "phi2": phi2}

        def grad(f):
            return sy.Matrix([sy.diff(f, X), sy.diff(f, Y)])

        def div(v):
            return sy.diff(v[0], X) + sy.diff(v[1], Y)

        J1 = [-D * grad(k) - z * D * self.psi * k * grad(phi1)
              for D, k, z in zip(self.D1, self.sol["c1"], self.z)]
        J2 = [-D * grad(k) - z * D * self.psi * k * grad(phi2)
              for D, k, z in zip(self.D2, self.sol["c2"], self.z)]
        self.J1, self.J2 = J1, J2
        f1 = [sy.diff(k, Tt) + div(J) for k, J in zip(self.sol["c1"], J1)]
        f2 = [sy.diff(k, Tt) + div(J) for k, J in zip(self.sol["c2"], J2)]
        fphi1 = self.F * sum(z * div(J) for z, J in zip(self.z, J1))
        fphi2 = self.F * sum(z * div(J) for z, J in zip(self.z, J2))
        lam = lambda e: sy.lambdify((X, Y, Tt), e, "numpy")
        self._f1 = [lam(e) for e in f1]
        self._f2 = [lam(e) for e in f2]
        self._fphi1, self._fphi2 = lam(fphi1), lam(fphi2)
        self._J2 = [(lam(J[0]), lam(J[1])) for J in J2]
        self._g1, self._g2, self._gphi, self._gJ = {}, {}, {}, {}
        for tag, n in MMS_NORMALS.items():
            nn = sy.Matrix(n)
            dot = lambda J: (J.T * nn)[0]
            self._g1[tag] = [lam(phi1 - phi2 - (1 / C) * dot(J)) for C, J in zip(self.C1, J1)]
            self._g2[tag] = [lam(phi1 - phi2 - (1 / C) * dot(J)) for C, J in zip(self.C2, J2)]
            self._gphi[tag] = lam(phi1 - phi2 - (1 / self.C_phi) * self.F
                                  * sum(z * dot(J) for z, J in zip(self.z, J1)))
            self._gJ[tag] = lam(-self.F * sum(z * (dot(Ja) - dot(Jb))
                                              for z, Ja, Jb in zip(self.z, J1, J2)))
        self._sol = {"c1": [lam(e) for e in self.sol["c1"]], "c2": [lam(e) for e in self.sol["c2"]],
                     "phi1": lam(phi1), "phi2": lam(phi2)}

    # -- evaluation ----------------------------------------------------------
    def _ev(self, f, x, t=None):
        t = self.t if t is None else t
        v = f(x[..., 0], x[..., 1], t)
        return np.broadcast_to(np.asarray(v, dtype=float), x.shape[:-1]).copy()

    def exact_field(self, P, which, k=None, t=None):
        """nodal interpolant ([nc, nd]) of the exact solution, side by cell tag."""
        f1 = self._sol[which + "1"] if k is None else self._sol[which + "1"][k]
        f2 = self._sol[which + "2"] if k is None else self._sol[which + "2"][k]
        ics = (P.cell_tag == 1)[:, None]
        return np.where(ics, self._ev(f1, P.X, t), self._ev(f2, P.X, t))

    def problem_kwargs(self):
        return dict(F=self.F, R=self.R, T=self.T, C_M=self.C_M, C_phi=self.C_phi, dt=self.dt,
                    z=self.z,
                    D_sub=[{1: a, 0: b} for a, b in zip(self.D1, self.D2)],
                    C_sub=[{1: a, 0: b} for a, b in zip(self.C1, self.C2)],
                    rho_sub={0: 0.0, 1: 0.0}, membrane_tags=(1, 2, 3, 4))

    # -- right-hand sides ----------------------------------------------------
    def _volume(self, P, f_ics, f_ecs):
        bq, wq = quad.duffy_rule(P.d, 5) if self.ufl_degree is None else quad.cell_rule(P.d, self.ufl_degree)
        xq = np.einsum("qa,cak->cqk", bq, P.X)
        ics = (P.cell_tag == 1)[:, None]
        fv = np.where(ics, self._ev(f_ics, xq), self._ev(f_ecs, xq))
        b = np.zeros(P.ndof)
        np.add.at(b, P.dofs(np.arange(P.nc)), np.einsum("q,c,cq,qi->ci", wq, P.vol, fv, bq))
        return b

    def _interface(self, P, b, g_by_tag, side, scale=1.0):
        """scale * int g[tag] * trace_side(v) dS(tag)."""
        bf, wf = quad.interval_rule(9 if self.ufl_degree is None else self.ufl_degree)
        for tag, g in g_by_tag.items():
            sel = np.flatnonzero(P.mem_tag == tag)
            if len(sel) == 0:
                continue
            fm = P.mem_facets[sel]
            cells = (P.mem_cell_i if side == "minus" else P.mem_cell_e)[sel]
            x = P.facet_points(fm, bf)
            L = P.basis_at(cells, x)
            W = wf[None, :] * P.farea[fm, None]
            np.add.at(b, P.dofs(cells), scale * np.einsum("fq,fq,fqa->fa", W, self._ev(g, x), L))

    def _neumann(self, P, b, Jxy, scale):
        """scale * int dot(J, n) v ds over exterior facets."""
        mesh = P.mesh
        ext = mesh.exterior_facets()
        bf, wf = quad.interval_rule(9 if self.ufl_degree is None else self.ufl_degree)
        cells = mesh.facet_cells[ext, 0]
        x = P.facet_points(ext, bf)
        L = P.basis_at(cells, x)
        n = P.fnormal[ext]
        W = wf[None, :] * P.farea[ext, None]
        Jn = self._ev(Jxy[0], x) * n[:, None, 0] + self._ev(Jxy[1], x) * n[:, None, 1]
        np.add.at(b, P.dofs(cells), scale * np.einsum("fq,fq,fqa->fa", W, Jn, L))

    def emi_rhs(self, P):
        b = self._volume(P, self._fphi1, self._fphi2)                       # solver.py:365-366
        # C_phi g_phi JUMP(v) dS(tag)                                         solver.py:359
        self._interface(P, b, self._gphi, "minus", P.C_phi)
        self._interface(P, b, self._gphi, "plus", -P.C_phi)
        self._interface(P, b, self._gJ, "plus", 1.0)                        # solver.py:369
        for k in range(P.N):                                                # solver.py:372-374
            self._neumann(P, b, self._J2[k], -P.F * P.z[k])
        return b

    def knp_rhs(self, P, k):
        b = self._volume(P, self._f1[k], self._f2[k])                       # solver.py:645-646
        g1 = {tag: g[k] for tag, g in self._g1.items()}
        g2 = {tag: g[k] for tag, g in self._g2.items()}
        self._interface(P, b, g1, "minus", self.C1[k])                      # solver.py:653
        self._interface(P, b, g2, "plus", -self.C2[k])                      # solver.py:654
        self._neumann(P, b, self._J2[k], -1.0)                              # solver.py:657
        return b

    # -- error norms (run_MMS_space.py:231-264) -------------------------------
    def l2_error(self, P, uh, which, k=None, mean_free=False, degree=None):
        """degree=5: the rule the reference's scripts ask for (run_MMS_space.py:213-246, metadata quadrature_degree 5)"""
        bq, wq = quad.duffy_rule(P.d, 6) if degree is None else quad.cell_rule(P.d, degree)
        xq = np.einsum("qa,cak->cqk", bq, P.X)
        uq = np.einsum("qa,ca->cq", bq, uh)
        f1 = self._sol[which + "1"] if k is None else self._sol[which + "1"][k]
        f2 = self._sol[which + "2"] if k is None else self._sol[which + "2"][k]
        ics = (P.cell_tag == 1)[:, None]
        ue = np.where(ics, self._ev(f1, xq), self._ev(f2, xq))
        W = wq[None, :] * P.vol[:, None]
        if mean_free:
            ue = ue - ((W * ue).sum() - (W * uq).sum())     # phi compared modulo the mean
        return float(np.sqrt(np.abs((W * (ue - uq) ** 2).sum())))
