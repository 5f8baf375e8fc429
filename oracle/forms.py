"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement of the discrete operators of the reference's KNP-EMI DG-P1
splitting scheme, evaluated *literally*: every UFL integrand of
src/knpemidg/solver.py is evaluated at quadrature points with explicit basis
function traces ('+'/'-' restrictions, jump, avg, FacetNormal), and scattered
into scipy sparse matrices the way dolfin's assembler adds macro-element
tensors.  The CUDA kernels use closed-form P1 integrals instead, so the two are
independent derivations of the same forms.

Reference lines restated here:
  EMI bilinear form a          src/knpemidg/solver.py:325-328, 346 (362 for MMS)
  EMI rhs L                    src/knpemidg/solver.py:309-310, 334-344 (359-374 MMS)
  EMI preconditioner form B    src/knpemidg/solver.py:377-395
  KNP bilinear form            src/knpemidg/solver.py:583-594
  KNP rhs                      src/knpemidg/solver.py:597-629 (645-657 MMS)
  post-step updates            src/knpemidg/solver.py:809-842
  n_g / plus / minus / facet-mean projection   src/knpemidg/utils.py:61-124

PINNED by the reference itself: oracle/refexec executes the reference's own, unmodified
form code (solver.py:270-403, 534-663) on a numeric dolfin stand-in, and every entry
assembled here agrees with what those forms assembled (tests/golden/ref_forms_*.npz,
tests/test_reference_golden.py: 1e-12 |entry| + 1e-14 max|row|, 0 entries fail).
Further pins: the MMS convergence study of tests/run_MMS_space.py (rate ~2), symmetry /
constant null space of the EMI operator, the rest-state known answer.

Conventions: a DG-P1 field is an array [nc, nd] of nodal values at the cell's
vertices; global dof = nd*cell + local vertex.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import quadrature as quad


class Problem:
    """Everything the forms need that does not change in time."""

    def __init__(self, mesh, cell_tag, facet_tag, *, F, R, T, C_M, C_phi, dt,
                 z, D_sub, rho_sub=None, membrane_tags=(), degree=1,
                 C_sub=None):
        mesh.init_topology()
        self.mesh = mesh
        self.d = mesh.gdim
        self.nd = mesh.nd
        self.nc = mesh.num_cells()
        self.cell_tag = np.asarray(cell_tag, dtype=np.int64)
        self.facet_tag = np.asarray(facet_tag, dtype=np.int64)
        self.F, self.R, self.T = float(F), float(R), float(T)
        self.C_M, self.C_phi, self.dt = float(C_M), float(C_phi), float(dt)
        self.psi = self.F / (self.R * self.T)            # solver.py:139
        self.z = np.asarray(z, dtype=np.float64)         # all N ions
        self.N = len(self.z)
        self.N_ions = self.N - 1                         # solver.py:69
        self.tau = 20.0 * self.d * degree                # solver.py:110-111
        # make_global (solver.py:1244-1258): DG0 by cell tag
        self.D = np.stack([self._by_tag(Dk) for Dk in D_sub])          # [N, nc]
        self.rho = self._by_tag(rho_sub) if rho_sub is not None else np.zeros(self.nc)
        self.C_tag = None if C_sub is None else np.stack([self._by_tag(Ck) for Ck in C_sub])
        self.membrane_tags = tuple(int(t) for t in membrane_tags)
        ext = mesh.coords.max(axis=0) - mesh.coords.min(axis=0)
        self.Lp = float(ext.max())                       # solver.py:383-391
        self._geometry()
        self._membrane_table()

    def _by_tag(self, table):
        out = np.zeros(self.nc)
        seen = np.zeros(self.nc, dtype=bool)
        for tag, val in table.items():
            m = self.cell_tag == int(tag)
            out[m] = float(val)
            seen |= m
        return out

    # -- geometry computed from vertex coordinates ------------------------
    def _geometry(self):
        mesh, d, nd = self.mesh, self.d, self.nd
        X = mesh.coords[mesh.cells]                      # [nc, nd, d]
        T = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)  # columns = edges
        Tinv = np.linalg.inv(T)                          # rows = grad lambda_1..d
        g = np.empty((self.nc, nd, d))
        g[:, 1:, :] = Tinv
        g[:, 0, :] = -Tinv.sum(axis=1)
        self.grad = g
        self.vol = mesh.cell_volume()
        self.h = mesh.cell_diameter()
        self.X = X
        # facets
        FX = mesh.coords[mesh.facet_verts]               # [nf, d, d]
        if d == 2:
            t = FX[:, 1] - FX[:, 0]
            area = np.linalg.norm(t, axis=1)
            n = np.column_stack([t[:, 1], -t[:, 0]]) / area[:, None]
        else:
            cr = np.cross(FX[:, 1] - FX[:, 0], FX[:, 2] - FX[:, 0])
            nrm = np.linalg.norm(cr, axis=1)
            area = 0.5 * nrm
            n = cr / nrm[:, None]
        c0 = mesh.facet_cells[:, 0]
        l0 = mesh.facet_local[:, 0]
        opp = X[c0, l0]                                  # vertex of cell0 opposite the facet
        sgn = np.sign(np.einsum("fk,fk->f", FX.mean(axis=1) - opp, n))
        self.fnormal = n * sgn[:, None]                  # outward from facet_cells[:,0] ('+')
        self.farea = area
        self.FX = FX

    def _membrane_table(self):
        """Membrane facets = interior facets whose tag is a membrane-model tag,
        ascending facet index (dlt_dof_extraction.py:34).  ECS side = lower cell
        tag, ICS side = higher (utils.py:80; README.md:67-72)."""
        mesh = self.mesh
        if len(self.membrane_tags) == 0:
            ids = np.zeros(0, dtype=np.int64)
        else:
            ids = np.flatnonzero(np.isin(self.facet_tag, self.membrane_tags)
                                 & (mesh.facet_cells[:, 1] >= 0))
        c0 = mesh.facet_cells[ids, 0]
        c1 = mesh.facet_cells[ids, 1]
        hi0 = self.cell_tag[c0] >= self.cell_tag[c1]     # chi('+') >= chi('-')
        self.mem_facets = ids
        self.mem_cell_i = np.where(hi0, c0, c1)          # 'minus' side of n_g
        self.mem_cell_e = np.where(hi0, c1, c0)          # 'plus' side of n_g
        self.mem_tag = self.facet_tag[ids]
        self.nm = len(ids)

    # -- helpers -----------------------------------------------------------
    def basis_at(self, cells, x):
        """P1 basis of `cells` [n] evaluated at physical points x [n, nq, d]."""
        g = self.grad[cells]                             # [n, nd, d]
        xv = self.X[cells]                               # [n, nd, d]
        return 1.0 + np.einsum("nak,nqak->nqa", g, x[:, :, None, :] - xv[:, None, :, :])

    def facet_points(self, facets, bary):
        return np.einsum("qa,fak->fqk", bary, self.FX[facets])

    def dofs(self, cells):
        return self.nd * np.asarray(cells)[:, None] + np.arange(self.nd)[None, :]

    @property
    def ndof(self):
        return self.nd * self.nc


def _scatter(n, rows, cols, vals):
    A = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(n, n))
    return A.tocsr()


def _macro_dofs(P, c0, c1):
    return np.concatenate([P.dofs(c0), P.dofs(c1)], axis=1)


def kappa_nodal(P, c_all):
    """kappa = F psi sum_k z_k^2 D_k c_k over ALL ions (solver.py:306)."""
    k = np.zeros((P.nc, P.nd))
    for i in range(P.N):
        k += P.F * P.z[i] ** 2 * P.D[i][:, None] * P.psi * c_all[i]
    return k


# ---------------------------------------------------------------------------
# EMI
# ---------------------------------------------------------------------------
def assemble_emi(P, c_all, phi_M, I_ch=None, splitting=True, mms=None):
    """Returns (A, B, b).  `c_all` [N, nc, nd]: concentrations of all ions
    (solved ones from c_prev_k, last = eliminated); phi_M [nm]; I_ch [N, nm]."""
    mesh, d, nd, nc = P.mesh, P.d, P.nd, P.nc
    n = P.ndof
    kap = kappa_nodal(P, c_all)
    rows, cols, vals = [], [], []
    b = np.zeros(n)

    # ---- cell integrals: inner(kappa grad u, grad v) dx  (solver.py:325) --
    bq, wq = quad.cell_rule(d, 3)
    kq = np.einsum("qm,cm->cq", bq, kap)
    GG = np.einsum("cik,cjk->cij", P.grad, P.grad)
    Ac = np.einsum("q,cq,c,cij->cij", wq, kq, P.vol, GG)
    dd = P.dofs(np.arange(nc))
    rows.append(np.repeat(dd[:, :, None], nd, 2)); cols.append(np.repeat(dd[:, None, :], nd, 1)); vals.append(Ac)
    # mass for B: kappa/Lp^2 u v dx (solver.py:393)
    Mk = np.einsum("q,cq,c,qi,qj->cij", wq, kq, P.vol, bq, bq) / P.Lp ** 2
    # rhs: -F z_k inner(D grad c_k, grad v) dx  (solver.py:309)
    for k in range(P.N):
        gc = np.einsum("cm,cmk->ck", c_all[k], P.grad)
        b_c = -P.F * P.z[k] * P.D[k][:, None] * P.vol[:, None] * np.einsum("ck,cik->ci", gc, P.grad)
        np.add.at(b, dd, b_c)

    # ---- interior facets tagged 0 (solver.py:326-328, 310) ----------------
    f0 = np.flatnonzero((mesh.facet_cells[:, 1] >= 0) & (P.facet_tag == 0))
    if len(f0):
        c0, c1 = mesh.facet_cells[f0, 0], mesh.facet_cells[f0, 1]
        bf, wf = quad.facet_rule(d, 3) if d == 2 else quad.duffy_rule(2, 4)
        x = P.facet_points(f0, bf)
        W = wf[None, :] * P.farea[f0, None]
        Lp_, Lm_ = P.basis_at(c0, x), P.basis_at(c1, x)
        nplus = P.fnormal[f0]
        kp = np.einsum("fqm,fm->fq", Lp_, kap[c0])
        km = np.einsum("fqm,fm->fq", Lm_, kap[c1])
        JV = np.concatenate([Lp_, -Lm_], axis=2)                   # jump(v)
        gnp = np.einsum("fak,fk->fa", P.grad[c0], nplus)
        gnm = np.einsum("fak,fk->fa", P.grad[c1], nplus)
        AG = np.concatenate([0.5 * kp[:, :, None] * gnp[:, None, :],
                             0.5 * km[:, :, None] * gnm[:, None, :]], axis=2)  # avg(kappa grad u).n+
        pen = P.tau / (0.5 * (P.h[c0] + P.h[c1]))
        Am = (-np.einsum("fq,fqb,fqa->fab", W, AG, JV)
              - np.einsum("fq,fqa,fqb->fab", W, AG, JV)
              + np.einsum("f,fq,fq,fqa,fqb->fab", pen, W, 0.5 * (kp + km), JV, JV))
        md = _macro_dofs(P, c0, c1)
        rows.append(np.repeat(md[:, :, None], 2 * nd, 2)); cols.append(np.repeat(md[:, None, :], 2 * nd, 1)); vals.append(Am)
        for k in range(P.N):
            gcp = np.einsum("fm,fmk->fk", c_all[k][c0], P.grad[c0]) * P.D[k][c0, None]
            gcm = np.einsum("fm,fmk->fk", c_all[k][c1], P.grad[c1]) * P.D[k][c1, None]
            flux = np.einsum("fk,fk->f", 0.5 * (gcp + gcm), nplus)
            bm = P.F * P.z[k] * np.einsum("fq,f,fqa->fa", W, flux, JV)
            np.add.at(b, md, bm)

    # ---- membrane facets (solver.py:334-346; MMS 359-362) -----------------
    if P.nm:
        fm = P.mem_facets
        ci, ce = P.mem_cell_i, P.mem_cell_e
        bf, wf = quad.facet_rule(d, 2) if d == 2 else quad.duffy_rule(2, 3)
        x = P.facet_points(fm, bf)
        W = wf[None, :] * P.farea[fm, None]
        Li, Le = P.basis_at(ci, x), P.basis_at(ce, x)
        JV = np.concatenate([Li, -Le], axis=2)           # JUMP(v, n_g) = v_i - v_e
        Am = P.C_phi * np.einsum("fq,fqa,fqb->fab", W, JV, JV)   # jump(u) jump(v), orientation free
        md = _macro_dofs(P, ci, ce)
        rows.append(np.repeat(md[:, :, None], 2 * nd, 2)); cols.append(np.repeat(md[:, None, :], 2 * nd, 1)); vals.append(Am)
        if mms is None:
            g = np.asarray(phi_M, dtype=float).copy()
            if not splitting:
                g = g - I_ch.sum(axis=0) / P.C_phi       # solver.py:337
            bm = P.C_phi * np.einsum("fq,f,fqa->fa", W, g, JV)
            np.add.at(b, md, bm)

    A = _scatter(n, np.concatenate([r.ravel() for r in rows]),
                 np.concatenate([c.ravel() for c in cols]),
                 np.concatenate([v.ravel() for v in vals]))
    dd3 = P.dofs(np.arange(nc))
    Bm = _scatter(n, np.repeat(dd3[:, :, None], nd, 2), np.repeat(dd3[:, None, :], nd, 1), Mk)
    B = (A + Bm).tocsr()
    if mms is not None:
        b += mms.emi_rhs(P)
    return A, B, b


# ---------------------------------------------------------------------------
# KNP
# ---------------------------------------------------------------------------
def alpha_sides(P, c_all, cells, L):
    """alpha_k = D_k z_k^2 c_k / sum_j D_j z_j^2 c_j on one side of membrane
    facets at quadrature points (solver.py:303, 603).  L: basis values [f,q,nd]."""
    num = []
    for k in range(P.N):
        ck = np.einsum("fqm,fm->fq", L, c_all[k][cells])
        num.append(P.D[k][cells, None] * P.z[k] ** 2 * ck)
    tot = sum(num)
    return [nk / tot for nk in num]


def assemble_knp(P, c_all, c_n, phi, phi_M, I_ch, splitting=True, f_source=None, mms=None):
    """Returns ([A_k], [b_k]) for the N-1 solved ions.  c_all [N,nc,nd] = c_prev_k
    plus eliminated ion; c_n [N_ions,nc,nd] = c_prev_n; phi [nc,nd] = potential
    just computed by the EMI step."""
    mesh, d, nd, nc = P.mesh, P.d, P.nd, P.nc
    n = P.ndof
    dd = P.dofs(np.arange(nc))
    bq, wq = quad.cell_rule(d, 3)
    Mref = np.einsum("q,qi,qj->ij", wq, bq, bq)
    GG = np.einsum("cik,cjk->cij", P.grad, P.grad)
    gphi = np.einsum("cm,cmk->ck", phi, P.grad)                  # grad(phi) per cell
    f0 = np.flatnonzero((mesh.facet_cells[:, 1] >= 0) & (P.facet_tag == 0))
    As, bs = [], []
    if len(f0):
        c0, c1 = mesh.facet_cells[f0, 0], mesh.facet_cells[f0, 1]
        bf, wf = quad.facet_rule(d, 3) if d == 2 else quad.duffy_rule(2, 4)
        x0 = P.facet_points(f0, bf)
        W0 = wf[None, :] * P.farea[f0, None]
        Lp_, Lm_ = P.basis_at(c0, x0), P.basis_at(c1, x0)
        nplus = P.fnormal[f0]
        JV0 = np.concatenate([Lp_, -Lm_], axis=2)
        gnp = np.einsum("fak,fk->fa", P.grad[c0], nplus)
        gnm = np.einsum("fak,fk->fa", P.grad[c1], nplus)
        pen = P.tau / (0.5 * (P.h[c0] + P.h[c1]))
        md0 = _macro_dofs(P, c0, c1)
    if P.nm:
        fm = P.mem_facets
        ci, ce = P.mem_cell_i, P.mem_cell_e
        bfm, wfm = quad.facet_rule(d, 5)
        xm = P.facet_points(fm, bfm)
        Wm = wfm[None, :] * P.farea[fm, None]
        Li, Le = P.basis_at(ci, xm), P.basis_at(ce, xm)
        mdm = _macro_dofs(P, ci, ce)
        phi_i = np.einsum("fqm,fm->fq", Li, phi[ci])
        phi_e = np.einsum("fqm,fm->fq", Le, phi[ce])
        if mms is None:
            al_i = alpha_sides(P, c_all, ci, Li)
            al_e = alpha_sides(P, c_all, ce, Le)
            I_tot = I_ch.sum(axis=0)                              # solver.py:315-322
    for k in range(P.N_ions):
        z, D = P.z[k], P.D[k]
        rows, cols, vals = [], [], []
        b = np.zeros(n)
        # cell terms (solver.py:586-587, 593, 597)
        drift = np.einsum("ck,cik->ci", gphi, P.grad)            # grad(phi).grad(v_i)
        lam_int = np.einsum("q,qj->j", wq, bq)                   # int lambda_j / |K|
        Ac = (P.vol[:, None, None] * Mref[None] / P.dt
              + (D * P.vol)[:, None, None] * GG
              + z * P.psi * (D * P.vol)[:, None, None] * drift[:, :, None] * lam_int[None, None, :])
        rows.append(np.repeat(dd[:, :, None], nd, 2)); cols.append(np.repeat(dd[:, None, :], nd, 1)); vals.append(Ac)
        b_c = np.einsum("c,ij,cj->ci", P.vol, Mref, c_n[k]) / P.dt
        np.add.at(b, dd, b_c)
        if f_source is not None and f_source[k] is not None:
            # f_source * v * dx(0)  (solver.py:599) - ECS cells only
            ecs = np.flatnonzero(P.cell_tag == 0)
            xq = np.einsum("qa,cak->cqk", bq, P.X[ecs])
            fv = f_source[k](xq) if callable(f_source[k]) else float(f_source[k]) * np.ones(xq.shape[:2])
            np.add.at(b, dd[ecs], np.einsum("q,c,cq,qi->ci", wq, P.vol[ecs], fv, bq))
        # interior facets tag 0 (solver.py:583, 588-590, 594)
        if len(f0):
            AG = np.concatenate([0.5 * D[c0, None, None] * gnp[:, None, :] * np.ones_like(Lp_),
                                 0.5 * D[c1, None, None] * gnm[:, None, :] * np.ones_like(Lm_)], axis=2)
            JDU = np.concatenate([D[c0, None, None] * Lp_, -D[c1, None, None] * Lm_], axis=2)   # jump(D u)
            un_p = np.maximum(D[c0] * np.einsum("fk,fk->f", gphi[c0], nplus), 0.0)
            un_m = np.maximum(D[c1] * np.einsum("fk,fk->f", gphi[c1], -nplus), 0.0)
            JUN = np.concatenate([un_p[:, None, None] * Lp_, -un_m[:, None, None] * Lm_], axis=2)  # jump(un u)
            Am = (-np.einsum("fq,fqb,fqa->fab", W0, AG, JV0)
                  - np.einsum("fq,fqa,fqb->fab", W0, AG, JV0)
                  + np.einsum("f,fq,fqa,fqb->fab", pen, W0, JV0, JDU)
                  - z * P.psi * np.einsum("fq,fqa,fqb->fab", W0, JV0, JUN))
            rows.append(np.repeat(md0[:, :, None], 2 * nd, 2)); cols.append(np.repeat(md0[:, None, :], 2 * nd, 1)); vals.append(Am)
        # membrane facets
        if P.nm:
            dphi = phi_i - phi_e                                  # orientation-free product of jumps below
            if mms is None:
                Ci = al_i[k] * P.C_M / (P.F * z * P.dt)           # solver.py:606
                Ce = al_e[k] * P.C_M / (P.F * z * P.dt)
                if splitting:
                    gi = phi_M[:, None] - P.dt / (P.C_M * al_i[k]) * I_ch[k][:, None] + (P.dt / P.C_M) * I_tot[:, None]
                    ge = phi_M[:, None] - P.dt / (P.C_M * al_e[k]) * I_ch[k][:, None] + (P.dt / P.C_M) * I_tot[:, None]
                else:
                    gi = phi_M[:, None] - P.dt / (P.C_M * al_i[k]) * I_ch[k][:, None]
                    ge = phi_M[:, None] - P.dt / (P.C_M * al_e[k]) * I_ch[k][:, None]
                # JUMP(C g v, n_g) dS(tag)   (solver.py:625)
                bi = np.einsum("fq,fq,fqa->fa", Wm, Ci * gi, Li)
                be = -np.einsum("fq,fq,fqa->fa", Wm, Ce * ge, Le)
            else:
                Ci = P.C_tag[k][ci][:, None] * np.ones_like(phi_i)
                Ce = P.C_tag[k][ce][:, None] * np.ones_like(phi_e)
                bi = np.zeros((P.nm, nd)); be = np.zeros((P.nm, nd))
            # - jump(phi) jump(C) avg(v) - jump(phi) avg(C) jump(v)  (solver.py:628-629)
            bi += -np.einsum("fq,fq,fqa->fa", Wm, dphi * (Ci - Ce) * 0.5, Li) \
                  - np.einsum("fq,fq,fqa->fa", Wm, dphi * 0.5 * (Ci + Ce), Li)
            be += -np.einsum("fq,fq,fqa->fa", Wm, dphi * (Ci - Ce) * 0.5, Le) \
                  + np.einsum("fq,fq,fqa->fa", Wm, dphi * 0.5 * (Ci + Ce), Le)
            np.add.at(b, mdm, np.concatenate([bi, be], axis=1))
        A = _scatter(n, np.concatenate([r.ravel() for r in rows]),
                     np.concatenate([c.ravel() for c in cols]),
                     np.concatenate([v.ravel() for v in vals]))
        if mms is not None:
            b += mms.knp_rhs(P, k)
        As.append(A); bs.append(b)
    return As, bs


# ---------------------------------------------------------------------------
# post-step (solver.py:809-842) and facet traces (utils.py:87-124)
# ---------------------------------------------------------------------------
def facet_mean_trace(P, field, side, degree=1):
    """pcws_constant_project(plus|minus(field, n_g), Q) restricted to membrane
    facets: facet mean of the one-sided trace.  side 'plus' = ECS (lower tag)."""
    cells = P.mem_cell_e if side == "plus" else P.mem_cell_i
    bf, wf = quad.facet_rule(P.d, degree)
    x = P.facet_points(P.mem_facets, bf)
    L = P.basis_at(cells, x)
    return np.einsum("q,fqm,fm->f", wf, L, field[cells])


def membrane_potential(P, phi):
    """phi_M = facet mean of JUMP(phi, n_g) = phi_i - phi_e (solver.py:813-814)."""
    return facet_mean_trace(P, phi, "minus") - facet_mean_trace(P, phi, "plus")


def nernst(P, c_k, z_k):
    """E = RT/(F z) ln(plus(c)/minus(c)), facet mean with the degree-4 rule
    (solver.py:299, 827)."""
    bf, wf = quad.facet_rule(P.d, 4)
    x = P.facet_points(P.mem_facets, bf)
    ce = np.einsum("fqm,fm->fq", P.basis_at(P.mem_cell_e, x), c_k[P.mem_cell_e])
    ci = np.einsum("fqm,fm->fq", P.basis_at(P.mem_cell_i, x), c_k[P.mem_cell_i])
    return P.R * P.T / (P.F * z_k) * np.einsum("q,fq->f", wf, np.log(ce / ci))


def eliminated_concentration(P, c_solved):
    """c_N = -(sum_k z_k c_k + rho)/z_N (solver.py:831-838; the reference L2
    projects this DG1 expression onto DG1, which is the identity)."""
    s = np.zeros((P.nc, P.nd))
    for k in range(P.N_ions):
        s += P.z[k] * c_solved[k]
    return -(s + P.rho[:, None]) / P.z[-1]
