"""ORACLE (test infrastructure only - never imported by the product path).

Stand-in for the few petsc4py objects the reference touches (src/knpemidg/solver.py:406-468,
502-529, 665-720, 755-789): Options, KSP, PC, NullSpace.  `KSP.solve` is a sparse DIRECT solve
(scipy SuperLU) whatever Krylov method and preconditioner the options name: the reference's
iterative answers agree with it to their KSP tolerances (rtol 1e-5 CG / 1e-7 GMRES), its own
direct path (MUMPS, solver.py:412-422, 671-681) to rounding.  A matrix with an attached (near)
null space of constants - the pure Neumann EMI operator, solver.py:465-466, 487 - is solved in
the bordered form (zero-mean solution; MUMPS null-pivot detection in the reference).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Options(dict):
    def __init__(self, prefix=None):
        super().__init__()
        self.prefix = prefix

    def setValue(self, k, v):
        self[k] = v


class PC:
    def setType(self, t):
        self.type = t

    def setFactorSolverType(self, t):
        self.factor = t


class NullSpace:
    def create(self, vectors=None, constant=False, comm=None):
        self.vectors = list(vectors or [])
        return self

    def remove(self, vec):
        for z in self.vectors:
            zz = z.array / np.linalg.norm(z.array)
            vec.array[:] -= zz * (zz @ vec.array)


class KSP:
    def create(self, comm=None):
        self.pc = PC()
        self.its = 0
        return self

    def setOptionsPrefix(self, p):
        self.prefix = p

    def setFromOptions(self):
        pass

    def setConvergenceHistory(self):
        pass

    def getPC(self):
        return self.pc

    def setOperators(self, A, P=None):
        self.A = A

    def getIterationNumber(self):
        return self.its

    def solve(self, b, x):
        A = self.A.A.tocsc()
        n = A.shape[0]
        if getattr(self.A, "nullspace", None) is not None:
            one = sp.csc_matrix(np.ones((n, 1)))
            K = sp.bmat([[A, one], [one.T, None]], format="csc")
            rhs = np.concatenate([b.array - b.array.mean(), [0.0]])
            x.array[:] = spla.spsolve(K, rhs)[:n]
        else:
            x.array[:] = spla.spsolve(A, b.array)
        self.its = 1


class _Enum:
    INSERT = 1
    FORWARD = 1
    AIJ = "aij"


class Mat:
    Type = _Enum


InsertMode = _Enum
ScatterMode = _Enum
