"""Host-side helpers that need neither the library nor the oracle."""
import pytest

import common  # noqa: F401  (puts the package on sys.path)


def test_h5lite_rejects_what_it_does_not_read(tmp_path):
    """knpemidg.h5lite is a deliberately small reader: anything outside the classic layout is an
    error naming what was met, never a guess (the positive test reads the reference's own mesh file:
    tests/test_reference_scripts.py::test_xdmf_hdf5_mesh_of_the_emix_example)"""
    from knpemidg import h5lite
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file at all")
    with pytest.raises(h5lite.H5Error, match="not an HDF5 file"):
        h5lite.File(str(p))
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + bytes(100))
    with pytest.raises(h5lite.H5Error, match="superblock version 2"):
        h5lite.File(str(p))
