"""Load vectors of the manufactured-solution terms of the reference's forms, computed from the
reference's OWN MMS data object (tests/mms_space.py / mms_time.py `setup_mms` -> MMSData with
symbolic `solution` / `rhs`, plus the `f1, f2, g_robin_1, g_robin_2, bdry, C_sub` entries of the
ion dictionaries built in tests/run_MMS_*.py).

Terms restated (src/knpemidg/solver.py):
  EMI  :365-366  int f_phi1 v dx(1) + int f_phi2 v dx(0)
       :359      sum_tag C_phi int g_phi[tag] (v_i - v_e) dS(tag)        (JUMP(v, n_g))
       :369      sum_tag int g_flux_cont[tag] v_e dS(tag)                (plus(v, n_g) = ECS side)
       :372-374  - F z_k int (bdry_k . n) v ds   for EVERY ion of ion_list
  KNP  :645-646  int f1 v dx(1) + int f2 v dx(0)
       :653-654  sum_tag int C_1 g_robin_1[tag] v_i dS(tag) - int C_2 g_robin_2[tag] v_e dS(tag)
       :657      - int (bdry . n) v ds
(the jump(phi) coupling terms :649-650 live in the KNP assembly kernel's mms mode).

The expressions are evaluated with knpemidg.symbolic (sympy; Constants such as the time `t`
are looked up when the vector is built), integrated with Gauss rules of degree >= 9 on facets
and 6 on cells - the integrands are smooth, the result is rule-insensitive far below the
discretisation error.  2D only: the reference's MMS data (interface tags 1..4, constant 2D
normals) are two-dimensional.  Host-side numpy; the vectors go to the device with
knp_field_set(KNP_F_LOAD_EMI / KNP_F_LOAD_KNP).
"""
from __future__ import annotations

import numpy as np

from . import symbolic as S
from .frontend import as_float


class ReferenceMMSLoads:
    def __init__(self, solver):
        self.solver = solver
        eng = solver.engine
        mesh = eng.mesh
        if mesh.gdim != 2:
            raise NotImplementedError("the reference's manufactured solutions are two-dimensional")
        mesh.init_topology()
        self.mesh, self.eng = mesh, eng
        self.nd = eng.nd
        self.X = mesh.coords[mesh.cells]                       # [nc, nd, d]
        self.vol = mesh.cell_volume()
        self.cell_tags = np.asarray(eng.cell_tags)
        self.bq, self.wq = S.simplex_rule(2, 5)                # cell rule, degree 8
        self.bf, self.wf = S.simplex_rule(1, 6)                # facet rule, degree 11
        mem = eng.mem
        self.mem_facet, self.mem_ci, self.mem_ce, self.mem_tag = mem["facet"], mem["cell_i"], mem["cell_e"], mem["tag"]
        fv = mesh.coords[mesh.facet_verts]                      # [nf, 2, d]
        self.fverts = fv
        self.farea = np.linalg.norm(fv[:, 1] - fv[:, 0], axis=1)
        ext = mesh.exterior_facets()
        self.ext = ext
        tang = fv[ext, 1] - fv[ext, 0]
        n = np.stack([tang[:, 1], -tang[:, 0]], axis=1) / self.farea[ext, None]
        cells = mesh.facet_cells[ext, 0]
        outward = mesh.coords[mesh.facet_verts[ext]].mean(axis=1) - self.X[cells].mean(axis=1)
        flip = (n * outward).sum(axis=1) < 0
        n[flip] *= -1.0
        self.ext_normal, self.ext_cells = n, cells

    # -- helpers -------------------------------------------------------------------------
    def _facet_points(self, facets):
        return np.einsum("qa,fak->fqk", self.bf, self.fverts[facets])          # [nf, nq, d]

    def _basis_at(self, cells, x):
        """P1 basis functions of `cells` at points x[f, q, :] -> [f, q, nd]"""
        Xc = self.X[cells]                                                      # [f, nd, d]
        T = np.stack([Xc[:, 1] - Xc[:, 0], Xc[:, 2] - Xc[:, 0]], axis=2)         # [f, d, 2]
        rel = x - Xc[:, None, 0, :]
        lam = np.linalg.solve(T[:, None, :, :], rel[..., None])[..., 0]         # [f, q, 2]
        return np.concatenate([1.0 - lam.sum(axis=-1, keepdims=True), lam], axis=-1)

    def _volume(self, b, f_ics, f_ecs):
        xq = np.einsum("qa,cak->cqk", self.bq, self.X)
        ics = (self.cell_tags == 1)[:, None]
        fv = np.where(ics, S.evaluate(f_ics, xq), S.evaluate(f_ecs, xq))
        b += np.einsum("q,c,cq,qi->ci", self.wq, self.vol, fv, self.bq)

    def _interface(self, b, g_by_tag, side, scale):
        for tag, g in g_by_tag.items():
            sel = np.flatnonzero(self.mem_tag == int(tag))
            if sel.size == 0:
                continue
            facets = self.mem_facet[sel]
            cells = (self.mem_ci if side == "minus" else self.mem_ce)[sel]
            x = self._facet_points(facets)
            L = self._basis_at(cells, x)
            W = self.wf[None, :] * self.farea[facets, None]
            np.add.at(b, cells, scale * np.einsum("fq,fq,fqa->fa", W, S.evaluate(g, x), L))

    def _neumann(self, b, J, scale):
        J = S.as_vec(J)
        x = self._facet_points(self.ext)
        L = self._basis_at(self.ext_cells, x)
        W = self.wf[None, :] * self.farea[self.ext, None]
        Jn = S.evaluate(J[0], x) * self.ext_normal[:, None, 0] + S.evaluate(J[1], x) * self.ext_normal[:, None, 1]
        np.add.at(b, self.ext_cells, scale * np.einsum("fq,fq,fqa->fa", W, Jn, L))

    # -- the two load vectors (`t` is ignored: the expressions carry the solver's time Constant) --
    def load_emi(self, t=None):
        s, mms = self.solver, self.solver.mms
        b = np.zeros((self.eng.nc, self.nd))
        self._volume(b, mms.rhs["volume_phi_1"], mms.rhs["volume_phi_2"])
        C_phi, F = float(s.C_phi), float(s.F)
        g = mms.rhs["bdry"]["u_phi"]
        self._interface(b, g, "minus", C_phi)
        self._interface(b, g, "plus", -C_phi)
        self._interface(b, mms.rhs["bdry"]["stress"], "plus", 1.0)
        for ion in s.ion_list:
            self._neumann(b, ion["bdry"], -F * as_float(ion["z"]))
        return b.ravel()

    def load_knp(self, k, t=None):
        ion = self.solver.ion_list[k]
        b = np.zeros((self.eng.nc, self.nd))
        self._volume(b, ion["f1"], ion["f2"])
        C1, C2 = as_float(ion["C_sub"][1]), as_float(ion["C_sub"][0])
        self._interface(b, ion["g_robin_1"], "minus", C1)
        self._interface(b, ion["g_robin_2"], "plus", -C2)
        self._neumann(b, ion["bdry"], -1.0)
        return b.ravel()
