"""Edge cases and error behaviour on the host-emulation build (CPU)."""
import edge_checks as ec


def test_mesh_without_membranes(emu_lib):
    ec.check_no_membrane(emu_lib)


def test_tagged_facets_without_model_carry_no_terms(emu_lib):
    ec.check_tagged_facets_without_model(emu_lib)


def test_contract_violations_raise(emu_lib):
    ec.check_contract_violations(emu_lib)


def test_krylov_nonconvergence_raises(emu_lib):
    ec.check_nonconvergence_raises(emu_lib)


def test_membrane_model_without_facets(emu_lib):
    ec.check_model_without_facets(emu_lib)
