"""Edge cases and error behaviour of the library (shared by the emulation and GPU test modules):
meshes without membranes, tagged facets without a model, contract violations, Krylov
non-convergence (the reference raises: ksp_error_if_not_converged, solver.py:428)."""
import numpy as np
import pytest

from common import kmesh
from knpemidg import _lib
from knpemidg.engine import Engine
from knpemidg.models import mm_hh

PHYS = dict(F=96485.0, R=8.314, T=300.0, C_M=0.02, C_phi=0.02 / 1e-4, dt=1e-4, z=[1.0, -1.0, 1.0])


def _engine(lib, mesh, cell_tags, facet_tags, mtags, tags=(0, 1)):
    D = [{int(t): d for t in tags} for d in (1.96e-9, 2.03e-9, 1.33e-9)]
    return Engine(mesh, cell_tags, facet_tags, membrane_tags=mtags, lib=lib, D_sub=D,
                  rho_sub={int(t): 0.0 for t in tags}, **PHYS)


def check_no_membrane(lib):
    """a mesh that is all ECS (no interface, nm = 0): the EMI operator annihilates constants,
    a uniform state is a fixed point of the full step"""
    mesh = kmesh.rectangle_mesh((0.0, 0.0), (4e-6, 2e-6), 8, 4, "crossed")
    mesh.init_topology()
    nc, nf = mesh.num_cells(), mesh.num_facets()
    eng = _engine(lib, mesh, np.zeros(nc, dtype=np.int64), np.zeros(nf, dtype=np.int64), (), tags=(0,))
    assert eng.nm == 0
    eng.set_concentrations_by_tag([{0: 4.0}, {0: 104.0}, {0: 100.0}])
    eng.initialize(pc=1)
    A = eng.ctx.matrix(0)
    assert np.abs(A @ np.ones(A.shape[1])).max() < 1e-12 * np.abs(A).max()
    assert abs(A - A.T).max() < 1e-12 * np.abs(A).max()
    for _ in range(2):
        eng.step()
    assert np.abs(eng.phi() - eng.phi().mean()).max() < 1e-12
    for k, c0 in enumerate((4.0, 104.0, 100.0)):
        assert np.abs(eng.concentration(k) - c0).max() < 1e-9 * c0


def check_tagged_facets_without_model(lib):
    """interior facets whose tag is neither 0 nor a membrane-model tag carry NO terms
    (SURVEY.md 8a): the two sides decouple completely"""
    mesh, sub, surf = kmesh.neuron_2d_mesh(1)       # (resolution 0 does not resolve the cell: its tags are inconsistent)
    fc = mesh.facet_cells
    inter = fc[:, 1] >= 0
    iface = inter & (sub.array()[fc[:, 0]] != sub.array()[np.maximum(fc[:, 1], 0)])
    assert iface.sum() > 0 and np.all(surf.array()[iface] == 1)
    eng = _engine(lib, mesh, sub.array(), surf.array(), ())          # tag 1 facets exist, no model given
    assert eng.nm == 0
    eng.set_concentrations_by_tag([{0: 4.0, 1: 125.0}, {0: 104.0, 1: 137.0}, {0: 100.0, 1: 12.0}])
    eng.initialize(pc=0)
    A = eng.ctx.matrix(0).tocsr()
    nd = eng.nd
    ics = np.repeat(sub.array() == 1, nd)
    assert A[ics][:, ~ics].nnz == 0 or np.abs(A[ics][:, ~ics]).max() == 0.0


def check_contract_violations(lib):
    mesh, sub, surf = kmesh.neuron_2d_mesh(0)
    eng = _engine(lib, mesh, sub.array(), surf.array(), (1,))
    ctx = eng.ctx
    with pytest.raises(_lib.KnpError, match="wrong element count"):
        ctx._call("knp_field_set", _lib.F_PHI, 0, np.zeros(3).ctypes.data_as(_lib._dp), 3)
    with pytest.raises(_lib.KnpError, match="out of range"):
        ctx.set_field(_lib.F_C, 7, np.zeros(ctx.n))
    with pytest.raises(_lib.KnpError, match="assemble first"):
        ctx.solve_emi()
    with pytest.raises(_lib.KnpError, match="row out of range"):
        ctx.membrane_register(0, [ctx.nm + 5], np.zeros((1, 4)), np.zeros((1, 17)))

    class NotAModel:
        __name__ = "mm_does_not_exist"
    with pytest.raises(_lib.KnpError, match="not compiled"):
        eng.add_membrane_model(1, NotAModel, ["K", "Cl", "Na"])
    bad = kmesh.rectangle_mesh((0.0, 0.0), (1.0, 1.0), 2, 2)
    bad.init_topology()
    fc = bad.facet_cells.copy()
    fc[fc[:, 1] >= 0, 1] = 0                               # claims cell 0 neighbours everything
    c2 = _lib.Context(0, lib)
    with pytest.raises(_lib.KnpError):
        c2.set_mesh(bad.coords, bad.cells, np.zeros(bad.num_cells(), dtype=np.int32), fc,
                    np.zeros(len(fc), dtype=np.int32), ())


def check_nonconvergence_raises(lib):
    mesh, sub, surf = kmesh.neuron_2d_mesh(0)
    eng = _engine(lib, mesh, sub.array(), surf.array(), (1,))
    eng.set_concentrations_by_tag([{0: 4.0, 1: 125.0}, {0: 104.0, 1: 137.0}, {0: 100.0, 1: 12.0}])
    eng.add_membrane_model(1, mm_hh, ["K", "Cl", "Na"], stimulus={"stim_amplitude": 10.0})
    eng.initialize(pc=0)
    eng.ode_phase()
    eng.ctx.assemble_emi()
    with pytest.raises(_lib.KnpError, match="did not converge"):
        eng.ctx.solve_emi(1e-14, 1e-300, 2)
    eng.ctx.assemble_knp()
    with pytest.raises(_lib.KnpError, match="did not converge"):
        eng.ctx.solve_knp(1e-15, 1e-300, 1)


def check_model_without_facets(lib):
    """a membrane model registered for a tag that no facet carries (an EMIx-like block in which only glia
    were placed): empty ODE tables on both sides, the run goes on and agrees with the oracle"""
    import solver_checks as sc
    from common import rel_err
    eng, O = sc.run_emix_block(lib, 8, 2)
    tags = np.bincount(eng.mem["tag"], minlength=3)
    assert tags[1] > 0 and tags[2] == 0
    assert [m.rows.size for m in eng.members] == [int(tags[1]), 0]
    assert rel_err(eng.phi_M(), O.phi_M) < 1e-6
