"""The C-ABI shared library loads and exports every symbol include/knpemi.h declares;
without a GPU the product library must refuse to create a context (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "knpemi.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(knp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product_lib():
    subprocess.run([sys.executable, os.path.join(ROOT, "knp-emi-dg_b200", "build.py")], check=True,
                   stdout=subprocess.DEVNULL)
    from knpemidg import _lib
    return _lib


def test_header_declares_the_bound_symbols(product_lib):
    assert set(declared_symbols()) == set(product_lib.SYMBOLS)


def test_product_library_exports_all_symbols(product_lib):
    dll = ctypes.CDLL(product_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(dll, name), name
    lib = product_lib.Lib(product_lib.LIB_PATH)
    assert lib.is_cuda()
    assert set(lib.models()) >= {"mm_hh", "mm_hh_no_stim", "mm_leak", "mm_glial_emix"}


def test_no_cpu_fallback(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = product_lib.Lib(product_lib.LIB_PATH)
    with pytest.raises(product_lib.KnpError, match="no CUDA device"):
        product_lib.Context(0, lib)


def test_emulation_build_is_not_the_product(emu_lib, product_lib):
    assert not emu_lib.is_cuda()
    assert os.path.dirname(emu_lib.path) != os.path.dirname(product_lib.LIB_PATH)
