"""`dolfin` stand-in for hosts WITHOUT legacy FEniCS.

The reference's run scripts and mesh scripts start with `from dolfin import *`
(examples/idealized-geometries/run_2D.py:3, make_mesh_2D.py:15) and use a small part of
dolfin for SETUP only: `Constant`, `Point`, `RectangleMesh`/`BoxMesh`, `MeshFunction`,
`SubDomain`/`CompiledSubDomain`, the `cells`/`facets`/`SubsetIterator` entity iterators,
`near`, and `File`/`Mesh` for dolfin-XML files; the manufactured-solution tests additionally
use UFL symbolics (`SpatialCoordinate`, `sin/cos`, `grad/div/dot/inner`, `Expression`,
`Measure`/`assemble` for the error norms), provided by knpemidg.symbolic on top of sympy.  This package provides exactly that on top of
`knpemidg.mesh`, so that those scripts execute UNCHANGED with the B200 `knpemidg` package:
put `knp-emi-dg_b200/shims` on `sys.path` (tests/test_reference_scripts.py does) - never when
a real dolfin is installed, which `knpemidg.dolfin_adapter` handles instead.

Everything here is host-side, once-per-run setup code; nothing is on the time-step path.
Mesh numbering follows knpemidg.mesh, not dolfin's internals: files
written and read back here are dolfin-XML in form and self-consistent, physics does not
depend on the numbering.
"""
from __future__ import annotations

import os
import re
import xml.etree.ElementTree as ET

import numpy as np

from knpemidg import mesh as _kmesh
from knpemidg.frontend import Constant  # noqa: F401  (re-exported: `from dolfin import *`)
# symbolic stand-ins for the UFL the reference's MMS tests use (tests/mms_space.py, run_MMS_*.py)
from knpemidg.symbolic import (Expression, Measure, SpatialCoordinate, assemble, cos, div, dot, exp,  # noqa: F401
                               grad, inner, ln, pi, sin, sqrt)

parameters = {}           # `parameters['ghost_mode'] = 'shared_vertex'` (solver.py:16): accepted, no effect

DOLFIN_EPS = 3.0e-16


def near(a, b, eps=DOLFIN_EPS):
    return abs(a - b) < eps


class Point:
    def __init__(self, *xs):
        if len(xs) == 1 and hasattr(xs[0], "__len__"):
            xs = tuple(xs[0])
        self._x = np.zeros(3)
        self._x[: len(xs)] = xs

    def x(self):
        return float(self._x[0])

    def y(self):
        return float(self._x[1])

    def z(self):
        return float(self._x[2])

    def array(self):
        return self._x.copy()

    def __getitem__(self, i):
        return float(self._x[i])

    def __len__(self):
        return 3


class Mesh(_kmesh.SimplexMesh):
    """SimplexMesh with dolfin's constructor `Mesh(path_to_xml)`."""

    def __init__(self, arg=None, cells=None):
        if arg is None and cells is None:                 # `mesh = Mesh()`, filled by XDMFFile.read
            super().__init__(np.zeros((0, 3)), np.zeros((0, 4), dtype=np.int32))
        elif isinstance(arg, (str, os.PathLike)):
            coords, cells = _read_mesh_xml(str(arg))
            super().__init__(coords, cells)
        elif isinstance(arg, _kmesh.SimplexMesh):
            super().__init__(arg.coords, arg.cells)
        else:
            super().__init__(arg, cells)

    class _Comm:
        rank, size = 0, 1

    def mpi_comm(self):
        return Mesh._Comm()

    def init(self, *dims):
        self.init_topology()

    def num_entities(self, dim):
        if dim == self.gdim:
            return self.num_cells()
        if dim == self.gdim - 1:
            return self.num_facets()
        if dim == 0:
            return self.num_vertices()
        raise NotImplementedError(dim)


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    m = _kmesh.rectangle_mesh((p0[0], p0[1]), (p1[0], p1[1]), nx, ny, diagonal)
    return Mesh(m)


def UnitSquareMesh(nx, ny, diagonal="right"):
    return Mesh(_kmesh.rectangle_mesh((0.0, 0.0), (1.0, 1.0), nx, ny, diagonal))


class FunctionSpace:
    """what user code builds itself: the facet space handed to a free-standing MembraneModel
    (run_calibration.py:13-14); the Solver's spaces are knpemidg.frontend.FunctionSpace"""

    def __init__(self, mesh, family, degree=0):
        self._mesh, self.family, self.degree = mesh, family, degree

    def mesh(self):
        return self._mesh


def info(msg):
    print(msg)


def BoxMesh(p0, p1, nx, ny, nz):
    m = _kmesh.box_mesh((p0[0], p0[1], p0[2]), (p1[0], p1[1], p1[2]), nx, ny, nz)
    return Mesh(m)


class _Entity:
    def __init__(self, mesh, dim, index):
        self._mesh, self._dim, self._index = mesh, dim, int(index)

    def index(self):
        return self._index

    def dim(self):
        return self._dim

    def entities(self, dim):
        """facet -> its cells (one on the boundary), cell -> its facets / vertices"""
        m = self._mesh
        m.init_topology()
        if self._dim == m.gdim - 1 and dim == m.gdim:
            c = m.facet_cells[self._index]
            return c[c >= 0]
        if self._dim == m.gdim and dim == m.gdim - 1:
            return m.cell_facets[self._index]
        if dim == 0:
            return m.cells[self._index] if self._dim == m.gdim else m.facet_verts[self._index]
        raise NotImplementedError((self._dim, dim))

    def midpoint(self):
        m = self._mesh
        if self._dim == m.gdim:
            return Point(m.coords[m.cells[self._index]].mean(axis=0))
        m.init_topology()
        return Point(m.coords[m.facet_verts[self._index]].mean(axis=0))


def cells(mesh):
    for i in range(mesh.num_cells()):
        yield _Entity(mesh, mesh.gdim, i)


def facets(mesh):
    for i in range(mesh.num_facets()):
        yield _Entity(mesh, mesh.gdim - 1, i)


class MeshFunction(_kmesh.MeshFunction):
    """dolfin's signatures: MeshFunction('size_t', mesh, dim[, value]) and
    MeshFunction('size_t', mesh, path_to_xml)."""

    def __init__(self, value_type, mesh, dim_or_path, value=0):
        assert value_type in ("size_t", "int", "uint"), value_type
        if isinstance(dim_or_path, (str, os.PathLike)):
            dim, vals = _read_meshfunction_xml(str(dim_or_path), mesh)
            super().__init__(mesh, dim, 0)
            self._a[:] = vals
        else:
            super().__init__(mesh, int(dim_or_path), value)

    @staticmethod
    def _idx(i):
        return i.index() if isinstance(i, _Entity) else i

    def __getitem__(self, i):
        return int(self._a[self._idx(i)])

    def __setitem__(self, i, v):
        self._a[self._idx(i)] = int(v)

    def set_all(self, v):
        self._a[:] = int(v)

    def rename(self, name, label=""):
        self._name = str(name)

    def name(self):
        return getattr(self, "_name", "f")


def SubsetIterator(mf, value):
    for i in np.flatnonzero(mf.array() == value):
        yield _Entity(mf.mesh(), mf.dim(), i)


class SubDomain:
    def inside(self, x, on_boundary):
        raise NotImplementedError

    def mark(self, mf, value):
        """mark the entities whose midpoint is inside (dolfin also tests the vertices; the
        reference's sub-domains are decided by the midpoint or by `on_boundary` alone)"""
        mesh = mf.mesh()
        mesh.init_topology()
        d = mesh.gdim
        if mf.dim() == d:
            mids = mesh.cell_midpoints()
            onb = np.zeros(len(mids), dtype=bool)
        else:
            mids = mesh.facet_midpoints()
            onb = mesh.facet_cells[:, 1] < 0
        for i, (x, b) in enumerate(zip(mids, onb)):
            if self.inside(x, bool(b)):
                mf[i] = value


class DomainBoundary(SubDomain):
    def inside(self, x, on_boundary):
        return on_boundary


class CompiledSubDomain(SubDomain):
    """C++ boolean expression in x[i], on_boundary, near(a, b) and DOLFIN_EPS, evaluated in Python."""

    def __init__(self, cpp, **params):
        expr = cpp.replace("&&", " and ").replace("||", " or ")
        expr = re.sub(r"!(?!=)", " not ", expr)
        expr = " ".join(expr.split())
        self._code = compile(expr, "<CompiledSubDomain>", "eval")
        self._params = dict(params)

    def inside(self, x, on_boundary):
        env = {"x": x, "on_boundary": on_boundary, "near": near, "DOLFIN_EPS": DOLFIN_EPS,
               "true": True, "false": False, "fabs": abs, "abs": abs}
        env.update(self._params)
        return bool(eval(self._code, {"__builtins__": {}}, env))


class File:
    """`File(path) << mesh|mesh_function` (dolfin XML; `.pvd` paths get a ParaView collection
    with one ASCII .vtu piece)."""

    def __init__(self, path):
        self.path = str(path)

    def __lshift__(self, obj):
        os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
        if self.path.endswith(".pvd"):
            _write_pvd(self.path, obj)
        elif isinstance(obj, _kmesh.MeshFunction):
            _write_meshfunction_xml(self.path, obj)
        elif isinstance(obj, _kmesh.SimplexMesh):
            _write_mesh_xml(self.path, obj)
        else:
            raise NotImplementedError(type(obj))
        return self

    def __rshift__(self, obj):
        raise NotImplementedError("use Mesh(path) / MeshFunction('size_t', mesh, path)")


class XDMFFile:
    """`XDMFFile([comm,] path)`: `read(mesh)`, `read(mesh_function, name)`, `write(obj)`; also a
    context manager (examples/emix-simulations/run_EMIx_simulation.py:160-193,
    examples/rat-neuron/run_rat_neuron.py:156-164, 204-205).  Heavy data in HDF5 is read with
    knpemidg.h5lite (gzip-chunked meshio files included), inline `Format="XML"` items as text;
    `write` produces XDMF with inline data (no HDF5 writer here)."""

    def __init__(self, *args):
        self.path = str(args[-1])
        self.parameters = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass

    def _item(self, node):
        """numpy array of a <DataItem>"""
        item = node if node.tag == "DataItem" else node.find("DataItem")
        dims = [int(v) for v in item.get("Dimensions").split()]
        text = (item.text or "").strip()
        if item.get("Format", "XML").upper() == "HDF":
            fname, dset = text.split(":", 1)
            from knpemidg import h5lite
            path = os.path.join(os.path.dirname(self.path), fname)
            cache = self.__dict__.setdefault("_h5", {})
            if path not in cache:
                cache[path] = h5lite.File(path)
            return cache[path][dset].read().reshape(dims)
        kind = item.get("NumberType", item.get("DataType", "Float"))
        return np.array(text.split(), dtype=float if kind == "Float" else np.int64).reshape(dims)

    def _grids(self):
        return list(ET.parse(self.path).getroot().iter("Grid"))

    def read(self, obj, name=None):
        grids = self._grids()
        if isinstance(obj, _kmesh.MeshFunction):
            return self._read_function(obj, name, grids)
        g = grids[0]
        topo = g.find("Topology")
        nd = {"triangle": 3, "tetrahedron": 4}.get(topo.get("TopologyType").lower())
        if nd is None:
            raise NotImplementedError("XDMF topology " + topo.get("TopologyType"))
        coords = self._item(g.find("Geometry"))[:, :nd - 1]
        _kmesh.SimplexMesh.__init__(obj, coords, self._item(topo).reshape(-1, nd))

    def _read_function(self, mf, name, grids):
        mesh = mf.mesh()
        for g in grids:
            for att in g.findall("Attribute"):
                if name is None or att.get("Name") == name:
                    vals = self._item(att).reshape(-1)
                    ents = np.sort(self._item(g.find("Topology")).reshape(len(vals), -1), axis=1)
                    break
            else:
                continue
            break
        else:
            raise RuntimeError(f"{self.path}: no attribute named {name!r}")
        mesh.init_topology()
        if ents.shape[1] != mf.dim() + 1:
            raise RuntimeError(f"{self.path}: '{name}' lives on {ents.shape[1] - 1}-dimensional entities, "
                               f"the MeshFunction on {mf.dim()}-dimensional ones")
        mine = np.sort(mesh.cells if mf.dim() == mesh.gdim else mesh.facet_verts, axis=1).astype(np.int64)
        if mine.shape == ents.shape and np.array_equal(mine, ents):
            mf.array()[:] = np.rint(vals).astype(np.int64)
            return
        # match entities by their vertex sets (the file's ordering need not be the mesh's)
        both = np.concatenate([mine, ents.astype(np.int64)])
        _, inv = np.unique(both, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        slot = np.full(inv.max() + 1, -1, dtype=np.int64)
        slot[inv[:len(mine)]] = np.arange(len(mine))
        where = slot[inv[len(mine):]]
        if (where < 0).any():
            raise RuntimeError(f"{self.path}: '{name}' holds entities that are not in the mesh")
        mf.array()[where] = np.rint(vals).astype(np.int64)

    def write(self, obj, *a):
        os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
        mesh = obj.mesh() if isinstance(obj, _kmesh.MeshFunction) else obj
        mesh.init_topology()
        d = mesh.gdim
        ents = mesh.cells
        if isinstance(obj, _kmesh.MeshFunction) and obj.dim() == d - 1:
            ents = mesh.facet_verts
        kind = {2: "PolyLine", 3: "Triangle", 4: "Tetrahedron"}[ents.shape[1]]

        def block(a, fmt):
            return "\n".join(" ".join(fmt % v for v in row) for row in np.atleast_2d(a))
        with open(self.path, "w") as f:
            f.write('<?xml version="1.0"?>\n<Xdmf Version="3.0"><Domain><Grid Name="mesh" GridType="Uniform">\n')
            f.write(f'<Topology NumberOfElements="{len(ents)}" TopologyType="{kind}" '
                    f'NodesPerElement="{ents.shape[1]}"><DataItem Dimensions="{len(ents)} {ents.shape[1]}" '
                    f'NumberType="UInt" Format="XML">\n{block(ents, "%d")}\n</DataItem></Topology>\n')
            f.write(f'<Geometry GeometryType="{"XY" if d == 2 else "XYZ"}"><DataItem Dimensions="{len(mesh.coords)} {d}" '
                    f'Format="XML">\n{block(mesh.coords, "%.17g")}\n</DataItem></Geometry>\n')
            if isinstance(obj, _kmesh.MeshFunction):
                v = obj.array()
                f.write(f'<Attribute Name="{getattr(obj, "_name", "f")}" AttributeType="Scalar" Center="Cell"><DataItem Dimensions="{len(v)} 1" '
                        f'NumberType="UInt" Format="XML">\n{block(v.reshape(-1, 1), "%d")}\n</DataItem></Attribute>\n')
            f.write("</Grid></Domain></Xdmf>\n")


def _write_pvd(path, obj):
    mesh = obj.mesh() if isinstance(obj, _kmesh.MeshFunction) else obj
    d = mesh.gdim
    piece = os.path.splitext(path)[0] + "000000.vtu"
    cells = mesh.cells
    with open(piece, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1">\n<UnstructuredGrid>\n')
        f.write(f'<Piece NumberOfPoints="{len(mesh.coords)}" NumberOfCells="{len(cells)}">\n<Points>\n'
                '<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        pts = np.zeros((len(mesh.coords), 3))
        pts[:, :d] = mesh.coords
        f.write("\n".join(" ".join("%.17g" % v for v in p) for p in pts))
        f.write('\n</DataArray>\n</Points>\n<Cells>\n<DataArray type="UInt32" Name="connectivity" format="ascii">\n')
        f.write("\n".join(" ".join(str(int(v)) for v in c) for c in cells))
        f.write('\n</DataArray>\n<DataArray type="UInt32" Name="offsets" format="ascii">\n')
        f.write(" ".join(str((i + 1) * (d + 1)) for i in range(len(cells))))
        f.write('\n</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
        f.write(" ".join([str(5 if d == 2 else 10)] * len(cells)))
        f.write("\n</DataArray>\n</Cells>\n")
        if isinstance(obj, _kmesh.MeshFunction) and obj.dim() == d:
            f.write('<CellData Scalars="f">\n<DataArray type="UInt32" Name="f" format="ascii">\n')
            f.write(" ".join(str(int(v)) for v in obj.array()))
            f.write("\n</DataArray>\n</CellData>\n")
        f.write("</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n'
                f'<DataSet timestep="0" part="0" file="{os.path.basename(piece)}" />\n</Collection>\n</VTKFile>\n')


# ---- dolfin XML -------------------------------------------------------------------------
_CELL = {2: "triangle", 3: "tetrahedron"}


def _write_mesh_xml(path, mesh):
    d = mesh.gdim
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n')
        f.write(f'  <mesh celltype="{_CELL[d]}" dim="{d}">\n    <vertices size="{mesh.num_vertices()}">\n')
        for i, x in enumerate(mesh.coords):
            xyz = " ".join(f'{ax}="{float(v)!r}"' for ax, v in zip("xyz", x))
            f.write(f'      <vertex index="{i}" {xyz} />\n')
        f.write(f'    </vertices>\n    <cells size="{mesh.num_cells()}">\n')
        for i, c in enumerate(mesh.cells):
            vs = " ".join(f'v{k}="{int(v)}"' for k, v in enumerate(c))
            f.write(f'      <{_CELL[d]} index="{i}" {vs} />\n')
        f.write("    </cells>\n  </mesh>\n</dolfin>\n")


def _read_mesh_xml(path):
    root = ET.parse(path).getroot()
    m = root.find("mesh")
    d = int(m.get("dim"))
    verts = m.find("vertices")
    coords = np.zeros((int(verts.get("size")), d))
    for v in verts:
        coords[int(v.get("index"))] = [float(v.get(ax)) for ax in "xyz"[:d]]
    cs = m.find("cells")
    cells = np.zeros((int(cs.get("size")), d + 1), dtype=np.int32)
    for c in cs:
        cells[int(c.get("index"))] = [int(c.get(f"v{k}")) for k in range(d + 1)]
    return coords, cells


def _write_meshfunction_xml(path, mf):
    """mesh_value_collection form (cell_index, local_entity, value), as dolfin writes it"""
    mesh = mf.mesh()
    mesh.init_topology()
    d, dim = mesh.gdim, mf.dim()
    a = mf.array()
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n  <mesh_function>\n')
        f.write(f'    <mesh_value_collection name="f" type="uint" dim="{dim}" size="{len(a)}">\n')
        if dim == d:
            for i, v in enumerate(a):
                f.write(f'      <value cell_index="{i}" local_entity="0" value="{int(v)}" />\n')
        else:
            for i, v in enumerate(a):
                f.write(f'      <value cell_index="{int(mesh.facet_cells[i, 0])}" '
                        f'local_entity="{int(mesh.facet_local[i, 0])}" value="{int(v)}" />\n')
        f.write("    </mesh_value_collection>\n  </mesh_function>\n</dolfin>\n")


def _read_meshfunction_xml(path, mesh):
    root = ET.parse(path).getroot()
    col = root.find("mesh_function").find("mesh_value_collection")
    dim = int(col.get("dim"))
    mesh.init_topology()
    n = mesh.num_cells() if dim == mesh.gdim else mesh.num_facets()
    vals = np.zeros(n, dtype=np.int64)
    for v in col:
        c, le, val = int(v.get("cell_index")), int(v.get("local_entity")), int(v.get("value"))
        if dim == mesh.gdim:
            vals[c] = val
        else:
            vals[mesh.cell_facets[c, le]] = val
    return dim, vals
