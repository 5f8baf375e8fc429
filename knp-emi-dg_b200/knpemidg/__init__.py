"""knpemidg - B200-native drop-in for the per-time-step hot path of
adajel/KNP-EMI-DG (reference: src/knpemidg/__init__.py:1-17 for the exported
names).  Host code is Python over a C-ABI CUDA library (include/knpemi.h);
there is no CPU fallback: using Solver/MembraneModel without the built
library raises."""
from knpemidg.mesh import SimplexMesh, MeshFunction  # noqa: F401

_LAZY = {
    "Solver": "knpemidg.solver", "SolverEMI": "knpemidg.solver_emi", "MembraneModel": "knpemidg.membrane",
    "interface_normal": "knpemidg.utils", "plus": "knpemidg.utils", "minus": "knpemidg.utils",
    "pcws_constant_project": "knpemidg.utils", "subdomain_marking_foo": "knpemidg.utils",
    "Constant": "knpemidg.frontend", "Expression": "knpemidg.symbolic",
}

__all__ = ["Solver", "SolverEMI", "MembraneModel", "subdomain_marking_foo", "interface_normal", "plus",
           "minus", "pcws_constant_project", "SimplexMesh", "MeshFunction"]


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(name)
