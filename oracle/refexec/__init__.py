"""ORACLE (test infrastructure only - never imported by the product path).

refexec = "execute the reference".  adajel/KNP-EMI-DG is pure Python on top of dolfin/UFL,
petsc4py and numbalsoda, none of which can be installed here.  This package supplies numeric
stand-ins for exactly those three imports so that the reference's OWN modules -
/root/reference/src/knpemidg/{solver,utils,membrane,dlt_dof_extraction}.py, unmodified, imported
from where they lie - run in this container: its form definitions, its orientation helpers, its
PDE<->ODE transfers and its time loop are executed, not restated.  What the stand-ins restate is
listed in ufl_numeric.py (degree estimation, quadrature rules, assembler), fake_petsc.py (direct
solves in place of CG/GMRES+BoomerAMG) and fake_numbalsoda.py (scipy LSODA).

    install()   registers the stand-ins as `dolfin`, `petsc4py`, `numbalsoda` and puts the
                reference's src/ first on sys.path.  The reference package is called `knpemidg`
                like the product's host mirror, so this must happen in a process of its own:
                tests/golden/make_reference_golden.py is that process; it writes the golden
                fixtures tests/golden/ref_*.npz that the parity tests (oracle and CUDA path)
                compare with.  /root/reference does not exist on the GPU box; the fixtures travel.
"""
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available():
    return os.path.exists(os.path.join(REFERENCE_SRC, "knpemidg", "solver.py"))


def install():
    from . import fake_dolfin, fake_petsc, fake_numbalsoda
    import numpy as np
    for old, new in (("float_", np.float64), ("Inf", np.inf)):     # NumPy 1.x names the reference still uses
        if not hasattr(np, old):
            setattr(np, old, new)
    if "knpemidg" in sys.modules and not sys.modules["knpemidg"].__file__.startswith(REFERENCE_SRC):
        raise RuntimeError("the product's knpemidg is already imported in this process")
    sys.modules["dolfin"] = fake_dolfin
    petsc4py = types.ModuleType("petsc4py")
    petsc4py.PETSc = fake_petsc
    sys.modules["petsc4py"] = petsc4py
    sys.modules["petsc4py.PETSc"] = fake_petsc
    sys.modules["numbalsoda"] = fake_numbalsoda
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import knpemidg                      # the reference's package
    assert knpemidg.__file__.startswith(REFERENCE_SRC), knpemidg.__file__
    return knpemidg
