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


def test_h5lite_writer_round_trip(tmp_path):
    """knpemidg.h5lite.Writer (the results.h5 of Solver.save_h5): nested groups, more links in a group
    than a default symbol node holds, int32 / int64 / float64 and empty datasets, read back by the reader"""
    import numpy as np
    from knpemidg import h5lite
    rng = np.random.default_rng(0)
    data = {"/mesh/coordinates": rng.random((11, 3)), "/mesh/topology": rng.integers(0, 11, (7, 4)).astype(np.int64),
            "/subdomains/values": np.arange(7, dtype=np.int32), "/empty": np.zeros((0, 3)),
            "/a/b/c/deep": np.array([1.5, -2.5])}
    for i in range(70):
        data[f"/potential/vector_{i}"] = rng.random(28)
    path = str(tmp_path / "t.h5")
    with h5lite.Writer(path) as w:
        for k, v in data.items():
            w.write(k, v)
        with pytest.raises(h5lite.H5Error, match="already written"):
            w.write("/mesh/topology", np.zeros(3))
    f = h5lite.File(path)
    assert f.keys() == ["a", "empty", "mesh", "potential", "subdomains"]
    assert len(f["/potential"].keys()) == 70
    for k, v in data.items():
        got = f[k].read()
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
    raw = open(path, "rb").read()
    assert raw[:8] == bytes([0x89, 0x48, 0x44, 0x46, 0x0D, 0x0A, 0x1A, 0x0A])       # signature
    assert int.from_bytes(raw[40:48], "little") == len(raw)                          # end-of-file address of the superblock
