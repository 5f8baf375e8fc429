"""ORACLE (test infrastructure only - never imported by the product path).

CPU stand-in for the reference's `pc_type hypre` (BoomerAMG) preconditioner
(src/knpemidg/solver.py:433-444, 688-701): hypre is not in /root/reference nor
in this image, so the CPU baseline uses an aggregation AMG V-cycle written with
scipy.sparse (same hierarchy idea as the CUDA library: DG1 -> region-wise
continuous P1 -> plain aggregation -> dense), rebuilt at every solve as the
reference rebuilds BoomerAMG at every `setOperators`.

parity unpinned: Krylov iteration counts are not comparable with hypre's; the
solutions are (both converge to the reference's KSP tolerances).
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def vertex_injection(P):
    """DG dof -> region-wise continuous vertex (dofs glued across tag-0 facets)."""
    mesh, nd = P.mesh, P.nd
    f0 = np.flatnonzero((mesh.facet_cells[:, 1] >= 0) & (P.facet_tag == 0))
    c0, c1 = mesh.facet_cells[f0, 0], mesh.facet_cells[f0, 1]
    v0, v1 = mesh.cells[c0], mesh.cells[c1]
    rows, cols = [], []
    for a in range(nd):
        for b in range(nd):
            m = v0[:, a] == v1[:, b]
            rows.append(nd * c0[m] + a)
            cols.append(nd * c1[m] + b)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    n = P.ndof
    G = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    ncomp, lab = connected_components(G, directed=False)
    return sp.csr_matrix((np.ones(n), (np.arange(n), lab)), shape=(n, ncomp))


def greedy_aggregate(A, theta=0.08):
    A = A.tocsr()
    n = A.shape[0]
    d = np.abs(A.diagonal())
    C = A.tocoo()
    strong = (C.row != C.col) & (np.abs(C.data) >= theta * np.sqrt(d[C.row] * d[C.col])) & (C.data != 0)
    S = sp.csr_matrix((np.ones(strong.sum()), (C.row[strong], C.col[strong])), shape=(n, n))
    ip, ix = S.indptr, S.indices
    agg = -np.ones(n, dtype=np.int64)
    na = 0
    for i in range(n):
        if agg[i] >= 0 or ip[i] == ip[i + 1]:
            continue
        nb = ix[ip[i]:ip[i + 1]]
        if np.all(agg[nb] < 0):
            agg[i] = na
            agg[nb] = na
            na += 1
    agg2 = agg.copy()
    for i in range(n):
        if agg[i] >= 0:
            continue
        nb = ix[ip[i]:ip[i + 1]]
        nb = nb[agg[nb] >= 0]
        if len(nb):
            agg2[i] = agg[nb[0]]
    agg = agg2
    for i in range(n):
        if agg[i] >= 0:
            continue
        agg[i] = na
        nb = ix[ip[i]:ip[i + 1]]
        agg[nb[agg[nb] < 0]] = na
        na += 1
    return sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, na))


class Plan:
    """Transfer operators; built once per problem (aggregates are reused across steps)."""

    def __init__(self, P, B, theta=0.08, coarse_size=200):
        self.nd = P.nd
        self.Ps = [vertex_injection(P)]
        A = (self.Ps[0].T @ B @ self.Ps[0]).tocsr()
        while A.shape[0] > coarse_size:
            T = greedy_aggregate(A, theta)
            if T.shape[1] >= 0.9 * A.shape[0]:
                break
            self.Ps.append(T)
            A = (T.T @ A @ T).tocsr()


class Hierarchy:
    """Numeric part, rebuilt for every matrix (Galerkin products, smoother data)."""

    def __init__(self, plan, A, omega=0.7):
        self.plan, self.omega = plan, omega
        nd = plan.nd
        self.As = [A.tocsr()]
        for T in plan.Ps:
            self.As.append((T.T @ self.As[-1] @ T).tocsr())
        Ab = self.As[0].tobsr((nd, nd))
        nb = A.shape[0] // nd
        D = np.zeros((nb, nd, nd))
        rows = np.repeat(np.arange(nb), np.diff(Ab.indptr))
        diag = rows == Ab.indices
        D[rows[diag]] = Ab.data[diag]
        self.Dinv = np.linalg.inv(D)
        self.l1 = [None] + [1.0 / np.asarray(abs(M).sum(axis=1)).ravel() for M in self.As[1:]]
        self.coarse = np.linalg.inv(self.As[-1].toarray())

    def _smooth0(self, x, b):
        r = b if x is None else b - self.As[0] @ x
        dx = self.omega * np.einsum("bij,bj->bi", self.Dinv, r.reshape(-1, self.plan.nd)).ravel()
        return dx if x is None else x + dx

    def _cycle(self, lev, b):
        if lev == len(self.As) - 1:
            return self.coarse @ b
        A = self.As[lev]
        if lev == 0:
            x = self._smooth0(None, b)
        else:
            x = self.l1[lev] * b
        T = self.plan.Ps[lev]
        x = x + T @ self._cycle(lev + 1, T.T @ (b - A @ x))
        if lev == 0:
            return self._smooth0(x, b)
        return x + self.l1[lev] * (b - A @ x)

    def apply(self, r):
        return self._cycle(0, r)
