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


def test_local_part_numbers_interior_cells_first():
    """the library overlaps the halo of a vector with the rows of the cells that have no ghost
    neighbour: the partitioner numbers those first (knpemidg/partition.py:LocalPart)"""
    import numpy as np
    from knpemidg import mesh as kmesh, partition
    mesh, sub, surf = kmesh.bundle_3d_mesh(dims=(8, 9, 9))
    part = partition.partition_cells(mesh, 4)
    for rank in range(4):
        L = partition.LocalPart(mesh, sub.array(), surf.array(), part, rank)
        fc = L.mesh.facet_cells
        both = fc[:, 1] >= 0
        touches_ghost = np.zeros(L.l2g.size, dtype=bool)
        ghost = np.arange(L.l2g.size) >= L.nc_owned
        np.logical_or.at(touches_ghost, fc[both, 0], ghost[fc[both, 1]])
        np.logical_or.at(touches_ghost, fc[both, 1], ghost[fc[both, 0]])
        assert 0 < L.nc_interior < L.nc_owned
        assert not touches_ghost[:L.nc_interior].any() and touches_ghost[L.nc_interior:L.nc_owned].all()
        assert np.all(np.diff(L.l2g[:L.nc_interior]) > 0)          # both groups keep ascending global order
