"""The reference EXECUTED (tests/golden/ref_*.npz, made by tests/golden/make_reference_golden.py from
the unmodified /root/reference/src/knpemidg on the numeric dolfin stand-in oracle/refexec) against
(i) the oracle restatement and (ii) the host-emulation build of the library.  The CUDA library is
compared with the same fixtures in tests/test_gpu_reference_golden.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_checks as gc
from oracle import refexec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


FAST = "ref_forms_2d,ref_forms_2d_nosplit,ref_forms_3d,ref_run_2d,ref_run_2d_picard,ref_run_2d_emi,ref_run_2d_passive,ref_mms"


@pytest.mark.skipif(not refexec.available(), reason="/root/reference is not present (GPU box)")
@pytest.mark.parametrize("only", [FAST] + [pytest.param(name, marks=pytest.mark.skipif(
    os.environ.get("KNP_SLOW_TESTS") != "1", reason="2.5 - 4 min; set KNP_SLOW_TESTS=1")) for name in ("ref_run_astro", "ref_run_emix")] + [pytest.param("ref_run_3d", marks=pytest.mark.skipif(
        os.environ.get("KNP_SLOW_TESTS") != "3d", reason="20 min (3D direct solves, 11 776 LSODA calls); set KNP_SLOW_TESTS=3d"))])
def test_fixtures_regenerate_from_the_reference(tmp_path, only):
    """the committed fixtures are what the reference produces here, today"""
    subprocess.run([sys.executable, os.path.join(gc.GOLDEN, "make_reference_golden.py"), str(tmp_path), "--only=" + only],
                   check=True, stdout=subprocess.DEVNULL, timeout=1800)
    for name in sorted(os.listdir(tmp_path)):
        new, old = np.load(tmp_path / name), np.load(os.path.join(gc.GOLDEN, name))
        assert sorted(new.files) == sorted(old.files), name
        for key in new.files:
            a, b = new[key], old[key]
            assert a.shape == b.shape, (name, key)
            scale = max(float(np.abs(b).max()), 1e-300) if b.size else 1.0
            assert np.abs(a.astype(float) - b.astype(float)).max(initial=0.0) <= 1e-11 * scale, (name, key)


@pytest.mark.skipif(not refexec.available(), reason="/root/reference is not present (GPU box)")
def test_reference_forms_do_not_depend_on_the_plus_side(tmp_path):
    """dolfin's choice of the '+' cell of an interior facet is internal; the reference's forms are
    written to be independent of it (utils.py:80, 90, 97).  Executing them with the two cells of
    every facet swapped must give the same tensors."""
    code = (
        "import sys, numpy as np\n"
        f"sys.argv = ['x', {str(tmp_path)!r}]\n"
        f"sys.path.insert(0, {gc.GOLDEN!r})\n"
        "import make_reference_golden as m\n"
        "mesh, sub, surf = m.kmesh.neuron_2d_mesh(1)\n"
        "mesh.init_topology()\n"
        "inner = mesh.facet_cells[:, 1] >= 0\n"
        "mesh.facet_cells[inner] = mesh.facet_cells[inner][:, ::-1]\n"
        "mesh.facet_local[inner] = mesh.facet_local[inner][:, ::-1]\n"
        "out = m.forms_case('2d', mesh, np.asarray(sub.array()), np.asarray(surf.array()), {1: m.mm_hh})\n"
        f"np.savez({str(tmp_path / 'swapped.npz')!r}, **out)\n")
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL, timeout=600)
    new, old = np.load(tmp_path / "swapped.npz"), np.load(os.path.join(gc.GOLDEN, "ref_forms_2d.npz"))
    import scipy.sparse as sp
    n = old["b_emi"].size
    for key, N in (("A_emi", n), ("B_emi", n), ("A_knp", 2 * n)):
        A = sp.coo_matrix((new[key + "_val"], (new[key + "_row"], new[key + "_col"])), shape=(N, N)).tocsr()
        B = sp.coo_matrix((old[key + "_val"], (old[key + "_row"], old[key + "_col"])), shape=(N, N)).tocsr()
        assert gc.entrywise_failures(A, B)[0] == 0, key
    for key in ("b_emi", "b_knp", "step_c", "step_c_elim"):
        assert gc.rel_err(new[key], old[key]) < 1e-11, key
    mem = np.isin(old["facet_tag"], old["membrane_tags"])     # (elsewhere n_g follows the '+' side: utils.py:80)
    for key in ("E0", "step_phi_M", "step_E"):
        assert gc.rel_err(new[key][..., mem], old[key][..., mem]) < 1e-11, key


@pytest.mark.parametrize("name", gc.FORM_CASES)
def test_oracle_forms_match_the_reference(name):
    gc.check_oracle_forms(name)


@pytest.mark.parametrize("name", gc.FORM_CASES)
def test_emulation_forms_match_the_reference(emu_lib, name):
    gc.check_library_forms(emu_lib, name)


def test_oracle_run_matches_the_reference_and_current_convention_deviation():
    """40 steps of the reference's solve_system_active (direct solves, LSODA; currents = whatever
    the integrator's LAST right-hand-side call left, membrane.py:108-114) against the oracle loop in
    both conventions.  Measured here: 4.1e-8 of the trace's range with 'last_call', 6.6e-8 with
    'end_state' (the GPU kernel's convention, I(y(t+dt))) - the convention costs < 1e-7, well inside
    the 1e-6 trace tolerance of north_star."""
    g = gc.run_golden()
    ref, n = g["phi_M_trace"], int(g["nsteps"])
    dev = {}
    for conv in ("last_call", "end_state"):
        tr, O = gc.oracle_run(conv, n)
        dev[conv] = gc.trace_deviation(tr, ref)
        assert dev[conv] < 1e-6, dev
        assert gc.rel_err(O.c.reshape(-1), g["final_c"]) < 1e-7
    assert abs(dev["end_state"] - dev["last_call"]) < 5e-7


def test_emulation_run_matches_the_reference(emu_lib):
    """the library's time loop (Krylov solves at 1e-10 / 1e-11: the reference fixture was made with
    direct solves) against the reference-executed traces: 1e-6 of the trace's range"""
    g = gc.run_golden()
    ref, n = g["phi_M_trace"], int(g["nsteps"])
    tr, eng = gc.library_run(emu_lib, n, 1e-10, 1e-11)
    assert gc.trace_deviation(tr, ref) < 1e-6
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    assert gc.rel_err(cfin, g["final_c"]) < 1e-7
    # at the reference's own tolerances (CG 1e-5, GMRES 1e-7) the traces agree to what those allow
    tr2, _ = gc.library_run(emu_lib, n)
    assert gc.trace_deviation(tr2, ref) < 1e-3


@pytest.mark.skipif(os.environ.get("KNP_SLOW_TESTS") != "1", reason="1.5 min (scipy LSODA in Python); set KNP_SLOW_TESTS=1")
def test_oracle_astro_run_matches_the_reference():
    """BASELINE configs[3] (run_tortuosity.py: three membrane tags, glial + neuronal models, rho != 0,
    tortuosity, the time-windowed K+/Na+ source) - the reference's own loop against the oracle loop"""
    g = gc.astro_golden()
    tr, O = gc.oracle_run_astro(int(g["nsteps"]), int(g["M"]))
    assert gc.trace_deviation(tr, g["phi_M_trace"]) < 1e-6
    assert gc.rel_err(O.c.reshape(-1), g["final_c"]) < 1e-7
    assert gc.rel_err(O.c_elim.reshape(-1), g["final_c_elim"]) < 1e-7


def test_emulation_astro_run_matches_the_reference(emu_lib):
    g = gc.astro_golden()
    tr, eng = gc.library_run_astro(emu_lib, int(g["nsteps"]), int(g["M"]), 1e-10, 1e-11)
    assert gc.trace_deviation(tr, g["phi_M_trace"]) < 1e-6
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    assert gc.rel_err(cfin, g["final_c"]) < 1e-7
    assert gc.rel_err(eng.concentration(2).reshape(-1), g["final_c_elim"]) < 1e-7


def test_emulation_picard_run_matches_the_reference(emu_lib):
    """Engine.pde_phase_picard against the reference's own solve_for_time_step_picard (solver.py:850-927,
    executed on oracle/refexec): 12 steps, membrane-potential traces and final concentrations"""
    g = np.load(os.path.join(gc.GOLDEN, "ref_run_2d_picard.npz"))
    tr, its, eng = gc.library_run_picard(emu_lib, int(g["nsteps"]))
    assert max(its) <= 4
    assert gc.trace_deviation(tr, g["phi_M_trace"]) < 1e-6
    cfin = np.concatenate([eng.concentration(k).reshape(-1) for k in range(2)])
    assert gc.rel_err(cfin, g["final_c"]) < 1e-7
    assert gc.rel_err(eng.concentration(2).reshape(-1), g["final_c_elim"]) < 1e-7


def test_emulation_solver_emi_matches_the_reference(emu_lib):
    """knpemidg.SolverEMI (run-script flow) against the reference's own solver_emi.py executed on oracle/refexec"""
    import solver_checks as sc
    sc.check_solver_emi_against_reference(emu_lib)


def test_emulation_passive_run_matches_the_reference(emu_lib):
    """Solver.solve_system_passive (non-splitting forms) against the reference's own passive loop"""
    import solver_checks as sc
    sc.check_passive_run_against_reference(emu_lib)


def test_oracle_mms_matches_the_reference():
    """BASELINE configs[0]: the MMS-mode forms (solver.py:349-374, 632-657) and data (tests/mms_space.py)"""
    gc.check_oracle_mms()


def test_reference_mms_study_converges_at_the_expected_rates():
    """what the reference's own scripts print when executed (they assert nothing): second order in space, first in time"""
    g = gc.mms_golden()
    e, h = g["space_errors"], g["space_h"]
    rates = np.log(e[1:] / e[:-1]) / np.log(h[1:] / h[:-1])[:, None]
    assert np.all(rates[-1] > 1.9) and np.all(rates[-1] < 2.05), rates
    t = g["time_errors"]
    assert np.all(np.log2(t[:-1] / t[1:])[-1] > 0.95), t


def test_emulation_mms_matches_the_reference(emu_lib):
    """the product path reproduces the L2 errors of the reference-executed MMS study (space r = 2, 3, 4; time dt_0/4, dt_0/8)
    to 1e-6 and the step-0 matrices entrywise"""
    gc.check_library_mms(emu_lib)


def test_emulation_emix_run_matches_the_reference(emu_lib):
    """BASELINE configs[4] = the workload bench.py times: its own engine builder on emix_like_mesh(9) against the reference's
    solve_system_active on the problem of run_EMIx_simulation.py (measured: trace 1.1e-8, concentrations 1e-9)"""
    gc.check_library_emix(emu_lib)


@pytest.mark.skipif(os.environ.get("KNP_SLOW_TESTS") != "1", reason="~2 min (scipy LSODA in Python); set KNP_SLOW_TESTS=1")
def test_oracle_emix_run_matches_the_reference():
    g = gc.emix_golden()
    tr, O = gc.oracle_run_emix(int(g["nsteps"]), int(g["M"]))
    assert gc.trace_deviation(tr, g["phi_M_trace"]) < 1e-6
    assert gc.rel_err(O.c.reshape(-1), g["final_c"]) < 1e-7
    assert gc.rel_err(O.c_elim.reshape(-1), g["final_c_elim"]) < 1e-7


def test_emulation_3d_bundle_run_matches_the_reference(emu_lib):
    """BASELINE configs[2]: the four-axon bundle of run_3D.py on its own resolution-0 mesh (15 552 tetrahedra), mm_hh +
    mm_hh_no_stim, 8 steps, against the reference's own loop (measured: trace 3e-8, concentrations 3e-9 .. 8e-9)"""
    gc.check_library_3d(emu_lib)
