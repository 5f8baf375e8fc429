"""GPU: edge cases / error behaviour (same checks as tests/test_edge_cases.py on the CUDA
library) and size-independent properties at the full bench workload."""
import numpy as np
import pytest

import edge_checks as ec

pytestmark = pytest.mark.gpu


def test_mesh_without_membranes(gpu_lib):
    ec.check_no_membrane(gpu_lib)


def test_tagged_facets_without_model_carry_no_terms(gpu_lib):
    ec.check_tagged_facets_without_model(gpu_lib)


def test_contract_violations_raise(gpu_lib):
    ec.check_contract_violations(gpu_lib)


def test_krylov_nonconvergence_raises(gpu_lib):
    ec.check_nonconvergence_raises(gpu_lib)


def test_properties_at_bench_size(gpu_lib):
    """BASELINE configs[2] at full size (419 904 cells, 5.04 M DOFs), where the oracle is too slow:
    the EMI operator is symmetric and annihilates constants, the SpMV is linear, the total amount
    of every ion barely moves in three steps (zero-flux boundary; only the membrane currents and
    the alpha-weighted capacitive terms, solver.py:603-629, move ions), the eliminated ion keeps
    the medium electroneutral."""
    import bench
    eng = bench.build_engine(bench.WORKLOAD_DIMS, 0)
    ctx = eng.ctx
    rng = np.random.default_rng(0)
    x, y = rng.uniform(-1, 1, ctx.n), rng.uniform(-1, 1, ctx.n)
    Ax, Ay = ctx.spmv(0, x), ctx.spmv(0, y)
    scale = np.abs(Ax).max()
    assert abs(y @ Ax - x @ Ay) < 1e-10 * abs(y @ Ax)
    assert np.abs(ctx.spmv(0, np.ones(ctx.n))).max() < 1e-10 * scale
    assert np.abs(ctx.spmv(0, 2.0 * x - 3.0 * y) - (2.0 * Ax - 3.0 * Ay)).max() < 1e-12 * scale
    vol = eng.mesh.cell_volume()

    def total(k):
        return float((eng.concentration(k).mean(axis=1) * vol).sum())
    m0 = [total(k) for k in range(3)]
    for _ in range(3):
        eng.step()
    assert max(eng.stats["knp_niter"]) <= 30 and max(eng.stats["emi_niter"]) <= 30
    for k in range(3):
        assert abs(total(k) - m0[k]) < 1e-4 * abs(m0[k])
    z = bench.PHYS["z"]
    charge = sum(z[k] * eng.concentration(k) for k in range(3))
    assert np.abs(charge).max() < 1e-9 * np.abs(eng.concentration(1)).max()
    assert np.isfinite(eng.phi_M()).all() and eng.phi_M().max() > -0.0744     # the stimulated axon depolarises


def test_properties_at_headline_size(gpu_lib):
    """BASELINE configs[4] at bench.py's headline size (EMIx-like block, 6 749 184 cells, 81 M DOFs):
    A_emi symmetric with the constants in its null space, the SpMV linear, glia at rest and the
    stimulated neurons depolarising, mass of every ion conserved up to the membrane exchange,
    electroneutrality of the eliminated ion, Krylov iteration counts in the expected range."""
    import bench
    eng = bench.build_engine_emix(bench.EMIX_M, 0)
    ctx = eng.ctx
    assert eng.dofs() >= 20_000_000
    rng = np.random.default_rng(0)
    x, y = rng.uniform(-1, 1, ctx.n), rng.uniform(-1, 1, ctx.n)
    Ax, Ay = ctx.spmv(0, x), ctx.spmv(0, y)
    scale = np.abs(Ax).max()
    assert abs(y @ Ax - x @ Ay) < 1e-10 * abs(y @ Ax)
    assert np.abs(ctx.spmv(0, np.ones(ctx.n))).max() < 1e-10 * scale
    assert np.abs(ctx.spmv(0, 2.0 * x - 3.0 * y) - (2.0 * Ax - 3.0 * Ay)).max() < 1e-12 * scale
    vol = eng.mesh.cell_volume()

    def total(k):
        return float((eng.concentration(k).mean(axis=1) * vol).sum())
    m0 = [total(k) for k in range(3)]
    for _ in range(5):
        eng.step()
    assert max(eng.stats["knp_niter"]) <= 12 and max(eng.stats["emi_niter"]) <= 30
    for k in range(3):
        assert abs(total(k) - m0[k]) < 1e-4 * abs(m0[k])
    z = bench.EMIX_PHYS["z"]
    charge = sum(z[k] * eng.concentration(k) for k in range(3))
    assert np.abs(charge).max() < 1e-9 * np.abs(eng.concentration(1)).max()
    pm, tag = eng.phi_M(), eng.mem["tag"]
    assert np.isfinite(pm).all()
    assert np.abs(pm[tag == 1] + 83.085).max() < 1.0            # glia stay at rest (mV)
    assert pm[tag == 2].max() > -74.0                            # the stimulated neurons depolarise
