"""Host-side simplex mesh, facet topology and the synthetic benchmark meshes.

This is setup-only code (runs once, on the host, numpy): it produces the flat
arrays that `knp_mesh_set` uploads.  It replaces what the reference obtains
from dolfin (`Mesh`, `MeshFunction`, facet<->cell connectivity):

* geometry of the structured generators follows SURVEY.md Appendix C
  (`RectangleMesh` right/crossed, `BoxMesh` 6 tets per box);
* tagging rules restate the reference's mesh scripts
  (tests/make_mesh_MMS.py:68-102, examples/idealized-geometries/make_mesh_2D.py:23-92,
  make_mesh_3D.py:18-111, examples/emix-simulations/make_mesh.py:58-122):
  cell tag by cell midpoint inside a box, facet tag by facet midpoint on a box
  face, exterior facets overwritten last, coordinates scaled last.

Facet numbering: facets are numbered by ascending sorted-vertex tuple.  The
reference orders ODE points by ascending dolfin facet index
(src/knpemidg/dlt_dof_extraction.py:18-48); dolfin's own facet numbering is
internal, so only "ascending facet index" is reproduced, not dolfin's indices.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "SimplexMesh", "MeshFunction", "rectangle_mesh", "box_mesh", "mms_mesh",
    "neuron_2d_mesh", "bundle_3d_mesh", "emix_like_mesh", "astro_like_mesh",
]


class SimplexMesh:
    """Triangle (d=2) or tetrahedron (d=3) mesh with facet topology."""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.gdim = self.coords.shape[1]
        self.nd = self.gdim + 1
        assert self.cells.shape[1] == self.nd
        self._topo = False

    # -- dolfin-flavoured accessors used by run scripts --------------------
    def coordinates(self):
        return self.coords

    def num_cells(self):
        return self.cells.shape[0]

    def num_vertices(self):
        return self.coords.shape[0]

    def num_facets(self):
        self.init_topology()
        return self.facet_cells.shape[0]

    class _Geom:
        def __init__(self, d):
            self._d = d

        def dim(self):
            return self._d

    def geometry(self):
        return SimplexMesh._Geom(self.gdim)

    def topology(self):
        return SimplexMesh._Geom(self.gdim)

    def hmin(self):
        return float(self.cell_diameter().min())

    # -- topology ----------------------------------------------------------
    def init_topology(self):
        """Build facet tables.  Local facet i of a cell is the one opposite
        local vertex i (the usual simplex convention)."""
        if self._topo:
            return
        nc, nd, d = self.cells.shape[0], self.nd, self.gdim
        nv = self.coords.shape[0]
        assert nv < 2_000_000, "facet key would overflow int64"
        cells = self.cells.astype(np.int64)
        keys = np.empty((nc, nd), dtype=np.int64)
        fverts = np.empty((nc, nd, d), dtype=np.int64)
        for i in range(nd):
            sub = np.sort(np.delete(cells, i, axis=1), axis=1)
            fverts[:, i, :] = sub
            k = sub[:, 0]
            for a in range(1, d):
                k = k * nv + sub[:, a]
            keys[:, i] = k
        flat = keys.ravel()
        order = np.argsort(flat, kind="stable")
        sk = flat[order]
        first = np.ones(sk.size, dtype=bool)
        first[1:] = sk[1:] != sk[:-1]
        fid_sorted = np.cumsum(first) - 1
        nf = int(fid_sorted[-1]) + 1
        cell_facets = np.empty(nc * nd, dtype=np.int32)
        cell_facets[order] = fid_sorted
        self.cell_facets = cell_facets.reshape(nc, nd)
        facet_cells = np.full((nf, 2), -1, dtype=np.int32)
        facet_local = np.full((nf, 2), -1, dtype=np.int32)
        start = np.flatnonzero(first)
        c_of = (order // nd).astype(np.int32)
        l_of = (order % nd).astype(np.int32)
        facet_cells[:, 0] = c_of[start]
        facet_local[:, 0] = l_of[start]
        second = np.flatnonzero(~first)
        facet_cells[fid_sorted[second], 1] = c_of[second]
        facet_local[fid_sorted[second], 1] = l_of[second]
        cnt = np.bincount(fid_sorted, minlength=nf)
        assert cnt.max() <= 2, "non-manifold mesh"
        self.facet_cells = facet_cells
        self.facet_local = facet_local
        self.facet_verts = fverts.reshape(nc * nd, d)[order[start]].astype(np.int32)
        self._topo = True

    def facet_midpoints(self):
        self.init_topology()
        return self.coords[self.facet_verts].mean(axis=1)

    def cell_midpoints(self):
        return self.coords[self.cells].mean(axis=1)

    def cell_volume(self):
        X = self.coords[self.cells]
        T = X[:, 1:, :] - X[:, :1, :]
        fact = 2.0 if self.gdim == 2 else 6.0
        return np.abs(np.linalg.det(T)) / fact

    def cell_diameter(self):
        """max vertex distance (UFL CellDiameter, used at solver.py:102-103)."""
        X = self.coords[self.cells]
        h = np.zeros(X.shape[0])
        for a in range(self.nd):
            for b in range(a + 1, self.nd):
                h = np.maximum(h, np.linalg.norm(X[:, a] - X[:, b], axis=1))
        return h

    def exterior_facets(self):
        self.init_topology()
        return np.flatnonzero(self.facet_cells[:, 1] < 0)

    def scale(self, s):
        self.coords *= s


class MeshFunction:
    """Minimal stand-in for dolfin.MeshFunction('size_t', mesh, dim, value)."""

    def __init__(self, mesh, dim, value=0):
        self._mesh = mesh
        self._dim = dim
        n = mesh.num_cells() if dim == mesh.gdim else mesh.num_facets()
        self._a = np.full(n, value, dtype=np.int64)

    def mesh(self):
        return self._mesh

    def dim(self):
        return self._dim

    def array(self):
        return self._a

    def where_equal(self, v):
        return np.flatnonzero(self._a == v)

    def __getitem__(self, i):
        return self._a[i]

    def __setitem__(self, i, v):
        self._a[i] = v


# ---------------------------------------------------------------------------
# structured generators
# ---------------------------------------------------------------------------
def rectangle_mesh(p0, p1, nx, ny, diagonal="right"):
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(x, y, indexing="xy")
    coords = np.column_stack([X.ravel(), Y.ravel()])
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ix, iy = ix.ravel(), iy.ravel()
    v0 = iy * (nx + 1) + ix
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    if diagonal == "right":
        cells = np.stack([np.column_stack([v0, v1, v3]),
                          np.column_stack([v0, v2, v3])], axis=1).reshape(-1, 3)
    elif diagonal == "crossed":
        xm = 0.5 * (x[:-1] + x[1:])
        ym = 0.5 * (y[:-1] + y[1:])
        XM, YM = np.meshgrid(xm, ym, indexing="xy")
        mid = np.column_stack([XM.ravel(), YM.ravel()])
        vm = (nx + 1) * (ny + 1) + iy * nx + ix
        coords = np.vstack([coords, mid])
        cells = np.stack([np.column_stack([v0, v1, vm]),
                          np.column_stack([v0, v2, vm]),
                          np.column_stack([v1, v3, vm]),
                          np.column_stack([v2, v3, vm])], axis=1).reshape(-1, 3)
    else:
        raise ValueError(diagonal)
    return SimplexMesh(coords, cells)


def box_mesh(p0, p1, nx, ny, nz):
    x = np.linspace(p0[0], p1[0], nx + 1)
    y = np.linspace(p0[1], p1[1], ny + 1)
    z = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    coords = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    v0 = iz * sz + iy * sy + ix
    v1 = v0 + sx
    v2 = v0 + sy
    v3 = v1 + sy
    v4 = v0 + sz
    v5 = v1 + sz
    v6 = v2 + sz
    v7 = v3 + sz
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
            (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    cells = np.stack([np.column_stack(t) for t in tets], axis=1).reshape(-1, 4)
    return SimplexMesh(coords, cells)


# ---------------------------------------------------------------------------
# tagging helpers
# ---------------------------------------------------------------------------
def _inside(mid, a, b, tol):
    ok = np.ones(mid.shape[0], dtype=bool)
    for k in range(mid.shape[1]):
        ok &= (mid[:, k] >= a[k] - tol) & (mid[:, k] <= b[k] + tol)
    return ok


def _on_box_surface(mid, a, b, tol):
    """facet midpoint lies on one of the faces of the box [a,b]."""
    d = mid.shape[1]
    on = np.zeros(mid.shape[0], dtype=bool)
    for k in range(d):
        others = np.ones(mid.shape[0], dtype=bool)
        for m in range(d):
            if m != k:
                others &= (mid[:, m] >= a[m] - tol) & (mid[:, m] <= b[m] + tol)
        on |= (np.abs(mid[:, k] - a[k]) < tol) & others
        on |= (np.abs(mid[:, k] - b[k]) < tol) & others
    return on


def mms_mesh(resolution):
    """tests/make_mesh_MMS.py:64-102: unit square 2^r x 2^r (right diagonals),
    ICS = [0.25,0.75]^2 tag 1, interface tags 1..4, exterior 5..8."""
    n = 2 ** resolution
    mesh = rectangle_mesh((0.0, 0.0), (1.0, 1.0), n, n)
    mesh.init_topology()
    tol = 1e-10
    sub = MeshFunction(mesh, 2, 0)
    sub.array()[_inside(mesh.cell_midpoints(), (0.25, 0.25), (0.75, 0.75), 0.0)] = 1
    surf = MeshFunction(mesh, 1, 0)
    m = mesh.facet_midpoints()
    a, b = (0.25, 0.25), (0.75, 0.75)
    s1 = (np.abs(m[:, 0] - a[0]) < tol) & (m[:, 1] >= a[1]) & (m[:, 1] <= b[1])
    s2 = (np.abs(m[:, 1] - a[1]) < tol) & (m[:, 0] >= a[0]) & (m[:, 0] <= b[0])
    s3 = (np.abs(m[:, 0] - b[0]) < tol) & (m[:, 1] >= a[1]) & (m[:, 1] <= b[1])
    s4 = (np.abs(m[:, 1] - b[1]) < tol) & (m[:, 0] >= a[0]) & (m[:, 0] <= b[0])
    arr = surf.array()
    # `side_1*1 or side_2*2 or ...` -> first true side wins
    for s, t in ((s4, 4), (s3, 3), (s2, 2), (s1, 1)):
        arr[s] = t
    ext = mesh.exterior_facets()
    me = m[ext]
    arr[ext[np.abs(me[:, 0] - 0.0) < tol]] = 5
    arr[ext[np.abs(me[:, 1] - 0.0) < tol]] = 6
    arr[ext[np.abs(me[:, 0] - 1.0) < tol]] = 7
    arr[ext[np.abs(me[:, 1] - 1.0) < tol]] = 8
    return mesh, sub, surf


def neuron_2d_mesh(resolution):
    """examples/idealized-geometries/make_mesh_2D.py:75-92: 62x4 um crossed mesh,
    ICS = [1,61]x[1,3] (tag 1), membrane facets tag 1, exterior 5, scaled to m."""
    nx = 31 * 2 ** resolution
    ny = 2 * 2 ** resolution
    mesh = rectangle_mesh((0.0, 0.0), (62.0, 4.0), nx, ny, "crossed")
    mesh.init_topology()
    tol = 1e-9
    sub = MeshFunction(mesh, 2, 0)
    surf = MeshFunction(mesh, 1, 0)
    a, b = (1.0, 1.0), (61.0, 3.0)
    sub.array()[_inside(mesh.cell_midpoints(), a, b, 0.0)] = 1
    surf.array()[_on_box_surface(mesh.facet_midpoints(), a, b, tol)] += 1
    surf.array()[mesh.exterior_facets()] = 5
    mesh.scale(1e-6)
    return mesh, sub, surf


def bundle_3d_mesh(resolution=0, dims=None, nblocks=1):
    """examples/idealized-geometries/make_mesh_3D.py:81-111: 32x0.9x0.9 um box,
    four axons (all cell tag 1), membrane tag 1 for the first axon and 2 for the
    other three, exterior 5, scaled to m.  `dims=(nx,ny,nz)` overrides the
    2^r refinement (SURVEY.md 8d 'scale 3' = (96,27,27)).  `nblocks` > 1 (weak-scaling
    runs) puts that many copies of the four-axon block side by side in y: a
    32 x 0.9*nblocks x 0.9 um box with nx x ny*nblocks x nz cells and 4*nblocks axons."""
    if dims is None:
        nx, ny, nz = 32 * 2 ** resolution, 9 * 2 ** resolution, 9 * 2 ** resolution
    else:
        nx, ny, nz = dims
    W = 0.9 * nblocks
    mesh = box_mesh((0.0, 0.0, 0.0), (32.0, W, 0.9), nx, ny * nblocks, nz)
    mesh.init_topology()
    tol = 1e-9
    sub = MeshFunction(mesh, 3, 0)
    surf = MeshFunction(mesh, 2, 0)
    cm = mesh.cell_midpoints()
    fm = mesh.facet_midpoints()
    if nblocks > 1:
        # every block is a y-translate of the first: fold y into the first block and tag once
        # (the axon surfaces lie at y in [0.2, 0.7] of a block, never on a block boundary)
        cm, fm = cm.copy(), fm.copy()
        for m in (cm, fm):
            m[:, 1] -= 0.9 * np.clip(np.floor(m[:, 1] / 0.9), 0, nblocks - 1)
    axons = [((5, 0.2, 0.2), (27, 0.4, 0.4), 1),
             ((5, 0.5, 0.5), (27, 0.7, 0.7), 2),
             ((5, 0.5, 0.2), (27, 0.7, 0.4), 2),
             ((5, 0.2, 0.5), (27, 0.4, 0.7), 2)]
    for a, b, tag in axons:
        sub.array()[_inside(cm, a, b, tol)] = 1
        surf.array()[_on_box_surface(fm, a, b, tol)] = tag
    surf.array()[mesh.exterior_facets()] = 5
    mesh.scale(1e-6)
    return mesh, sub, surf


def emix_like_mesh(M, n_cells=100, seed=1234, length=1.0e-3):
    """Synthetic EMIx-like tissue block: BoxMesh M^3 x 6 tets,
    ~n_cells non-touching axis-aligned ICS cuboids placed on the grid by
    numpy.random.default_rng(seed); alternately glia (cell tag 1) and neuron
    (cell tag 2) as in examples/emix-simulations/run_EMIx_simulation.py:172-185;
    membrane facet tag = cell tag of the enclosed cuboid; exterior facets get
    tag 5 (no terms).  Coordinates in cm-like units scaled by `length`."""
    mesh = box_mesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), M, M, M)
    mesh.init_topology()
    rng = np.random.default_rng(seed)
    occ = np.zeros((M, M, M), dtype=np.int32)  # 0 free, -1 halo, >0 tag
    boxes = []
    tries = 0
    while len(boxes) < n_cells and tries < 200 * n_cells:
        tries += 1
        sz = rng.integers(max(2, M // 12), max(3, M // 5), size=3)
        lo = np.array([rng.integers(1, M - s - 1) if M - s - 1 > 1 else 1 for s in sz])
        hi = lo + sz
        if np.any(hi > M - 1):
            continue
        region = occ[max(lo[0] - 1, 0):hi[0] + 1, max(lo[1] - 1, 0):hi[1] + 1,
                     max(lo[2] - 1, 0):hi[2] + 1]
        if np.any(region != 0):
            continue
        tag = 1 + (len(boxes) % 2)
        region[...] = -1
        occ[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = tag
        boxes.append((lo, hi, tag))
    sub = MeshFunction(mesh, 3, 0)
    surf = MeshFunction(mesh, 2, 0)
    # box index of each cell (6 tets per box, x fastest)
    box_tag = np.where(occ > 0, occ, 0).transpose(2, 1, 0).ravel()  # [iz,iy,ix]
    sub.array()[:] = np.repeat(box_tag, 6)
    fc = mesh.facet_cells
    interior = fc[:, 1] >= 0
    t0 = sub.array()[fc[:, 0]]
    t1 = np.where(interior, sub.array()[np.maximum(fc[:, 1], 0)], t0)
    memb = interior & (t0 != t1)
    surf.array()[memb] = np.maximum(t0, t1)[memb]
    surf.array()[~interior] = 5
    mesh.scale(length)
    mesh.n_ics_cells = len(boxes)
    return mesh, sub, surf


def astro_like_mesh(M, length=5.0e-4):
    """Synthetic stand-in for the mesh of examples/local-astrocyte-depolarization/run_tortuosity.py
    (BASELINE configs[3]; the emimesh data set behind meshes/synapse.yml is not in the checkout):
    BoxMesh M^3 x 6 tets (M >= 8) with two neuronal cuboids (cell tag 1; membrane facet tags 1 and 3)
    and one glial cuboid (cell tag 2; membrane tag 2) - the tag layout the script's
    `ode_models = {1: mm_hh, 2: mm_glial, 3: mm_hh}` and `rho_sub = {0, 1, 2}` expect
    (run_tortuosity.py:107, 298).  Exterior facets 5.  `mesh.source_box` = (lo, hi), the ECS slab
    between the first neuron and the glial cell where the K+/Na+ source of run_tortuosity.py:180-200
    acts.  Coordinates in cm (the script's units), edge `length`."""
    assert M >= 8
    mesh = box_mesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), M, M, M)
    mesh.init_topology()
    occ = np.zeros((M, M, M), dtype=np.int32)      # [ix, iy, iz] -> membrane tag of the enclosing cuboid
    boxes = [((M // 8, M // 4, M // 4), (3 * M // 8, 3 * M // 4, M // 2), 1),
             ((M // 2, M // 4, M // 4), (7 * M // 8, M // 2, 3 * M // 4), 2),
             ((M // 8, M // 4, 5 * M // 8), (3 * M // 8, 3 * M // 4, 7 * M // 8), 3)]
    for lo, hi, tag in boxes:
        occ[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = tag
    mem_of_box = occ.transpose(2, 1, 0).ravel()      # box index = iz*M*M + iy*M + ix
    mem_of_cell = np.repeat(mem_of_box, 6)
    sub = MeshFunction(mesh, 3, 0)
    surf = MeshFunction(mesh, 2, 0)
    sub.array()[:] = np.where(mem_of_cell == 2, 2, np.where(mem_of_cell > 0, 1, 0))
    fc = mesh.facet_cells
    interior = fc[:, 1] >= 0
    m0 = mem_of_cell[fc[:, 0]]
    m1 = np.where(interior, mem_of_cell[np.maximum(fc[:, 1], 0)], m0)
    memb = interior & (m0 != m1)
    surf.array()[memb] = np.maximum(m0, m1)[memb]
    surf.array()[~interior] = 5
    mesh.scale(length)
    h = length / M
    mesh.source_box = (np.array([3 * M // 8, M // 4, M // 4]) * h, np.array([M // 2, 3 * M // 4, 3 * M // 4]) * h)
    return mesh, sub, surf
