"""Shared fixtures of the parity tests: a synthetic case is built once as an
oracle `Problem` (CPU restatement, oracle/) and once as a library context
(knpemidg._lib.Context) holding the same mesh, parameters and fields.

`lib_for(kind)`: "gpu" -> the product libknpemi.so (CUDA, sm_100a);
                 "emu" -> the host-emulation build of the same sources
                          (tests/emu/libknpemi_emu.so), CPU test-suite only.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "knp-emi-dg_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from knpemidg import _lib  # noqa: E402
from knpemidg import mesh as kmesh  # noqa: E402
from oracle import forms  # noqa: E402

_LIBS = {}


def lib_for(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    if kind == "gpu":
        lib = _lib.get()
    else:
        path = os.path.join(ROOT, "tests", "emu", "libknpemi_emu.so")
        subprocess.run([sys.executable, os.path.join(PKG, "build.py"), "--emu"], check=True,
                       stdout=subprocess.DEVNULL)
        lib = _lib.Lib(path)
        assert not lib.is_cuda()
    _LIBS[kind] = lib
    return lib


PHYS = dict(F=96485.0, R=8.314, T=300.0, C_M=0.02, dt=1.0e-4)
D_PHYS = [1.96e-9, 2.03e-9, 1.33e-9]      # K, Cl, Na (run_2D.py:117-139)
Z_PHYS = [1.0, -1.0, 1.0]
C_ICS = [125.0, 137.0, 12.0]
C_ECS = [4.0, 104.0, 100.0]


def make_mesh(name):
    if name == "2d":
        return kmesh.neuron_2d_mesh(0) + ((1,),)
    if name == "2d_r1":
        return kmesh.neuron_2d_mesh(1) + ((1,),)
    if name == "3d":
        return kmesh.bundle_3d_mesh(dims=(16, 9, 9)) + ((1, 2),)
    if name == "3d_small":
        return kmesh.bundle_3d_mesh(dims=(8, 9, 9)) + ((1, 2),)
    if name == "3d_r0":
        return kmesh.bundle_3d_mesh(0) + ((1, 2),)
    if name == "emix":
        return kmesh.emix_like_mesh(10, n_cells=6) + ((1, 2),)
    if name in ("unstr2d", "unstr3d"):
        return unstructured_mesh(2 if name.endswith("2d") else 3) + ((1,),)
    raise ValueError(name)


def unstructured_mesh(d, n=None, seed=3):
    """UNSTRUCTURED simplex mesh with an embedded cell (scaled to um): cells whose midpoint lies in
    [0.3, 0.7]^d get tag 1, the interface facets membrane tag 1, exterior facets 5.
    2D: scipy Delaunay triangulation of jittered points (irregular valences).  3D: Delaunay of
    near-grid points produces slivers, on which the SIP form with h = max edge (the reference's
    CellDiameter) loses coercivity - so the 6-tets-per-box mesh with jittered interior vertices.
    In both, the vertex numbering, the cell numbering and every cell's local vertex order are
    randomly permuted: neighbour tables, facet permutations and orientations are arbitrary, which
    the structured generators never produce."""
    rng = np.random.default_rng(seed)
    m = n or (14 if d == 2 else 7)
    if d == 2:
        from scipy.spatial import Delaunay
        g = np.linspace(0.0, 1.0, m + 1)
        pts = np.stack(np.meshgrid(*([g] * d), indexing="ij"), axis=-1).reshape(-1, d)
        interior = np.all((pts > 1e-12) & (pts < 1 - 1e-12), axis=1)
        pts[interior] += rng.uniform(-0.3, 0.3, (int(interior.sum()), d)) / m
        cells = Delaunay(pts).simplices.astype(np.int32)
        X = pts[cells]
        vol = np.abs(np.linalg.det(X[:, 1:, :] - X[:, :1, :]))
        cells = cells[vol > 1e-9 * vol.max()]
    else:
        base = kmesh.box_mesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), m, m, m)
        pts, cells = base.coords.copy(), base.cells.copy()
        interior = np.all((pts > 1e-12) & (pts < 1 - 1e-12), axis=1)
        pts[interior] += rng.uniform(-0.15, 0.15, (int(interior.sum()), d)) / m
    vperm = rng.permutation(len(pts))                            # new vertex numbering
    inv = np.empty_like(vperm)
    inv[vperm] = np.arange(len(pts))
    pts = pts[vperm]
    cells = inv[cells]
    cells = cells[rng.permutation(len(cells))]                   # new cell numbering
    for k in range(len(cells)):                                  # arbitrary local vertex order
        cells[k] = cells[k][rng.permutation(d + 1)]
    mesh = kmesh.SimplexMesh(pts * 1e-6, cells.astype(np.int32))
    mesh.init_topology()
    sub = kmesh.MeshFunction(mesh, d, 0)
    surf = kmesh.MeshFunction(mesh, d - 1, 0)
    mid = mesh.cell_midpoints() / 1e-6
    sub.array()[np.all((mid > 0.3) & (mid < 0.7), axis=1)] = 1
    fc = mesh.facet_cells
    inter = fc[:, 1] >= 0
    t0 = sub.array()[fc[:, 0]]
    t1 = np.where(inter, sub.array()[np.maximum(fc[:, 1], 0)], t0)
    surf.array()[inter & (t0 != t1)] = 1
    surf.array()[~inter] = 5
    return mesh, sub, surf


class Case:
    """physiological KNP-EMI case with seeded, perturbed fields"""

    def __init__(self, name, lib, seed=0, splitting=True, D_scale=(1.0, 1.0)):
        mesh, sub, surf, mtags = make_mesh(name)
        self.mesh, self.sub, self.surf, self.mtags = mesh, sub, surf, mtags
        tags = np.unique(sub.array())
        region = np.searchsorted(tags, sub.array()).astype(np.int32)
        self.tags = tags
        dt, C_M = PHYS["dt"], PHYS["C_M"]
        # per-region diffusion (second entry scales ICS regions to exercise jump(D u))
        D_sub = [{int(t): D_PHYS[k] * (D_scale[0] if t == 0 else D_scale[1]) for t in tags}
                 for k in range(3)]
        rho_sub = {int(t): 0.0 for t in tags}
        self.P = forms.Problem(mesh, sub.array(), surf.array(), F=PHYS["F"], R=PHYS["R"], T=PHYS["T"],
                               C_M=C_M, C_phi=C_M / dt, dt=dt, z=Z_PHYS, D_sub=D_sub, rho_sub=rho_sub,
                               membrane_tags=mtags)
        P = self.P
        rng = np.random.default_rng(seed)
        ics = (sub.array() != 0)[:, None]
        c_all = np.stack([np.where(ics, C_ICS[k], C_ECS[k]) * (1.0 + 0.01 * rng.uniform(-1, 1, (P.nc, P.nd)))
                          for k in range(3)])
        self.c_all = c_all
        self.c_n = c_all[:2] * (1.0 + 0.001 * rng.uniform(-1, 1, (2, P.nc, P.nd)))
        self.phi = np.where(ics, -0.07, 0.0) * (1.0 + 0.01 * rng.uniform(-1, 1, (P.nc, P.nd)))
        self.phi_M = -0.07 * (1.0 + 0.01 * rng.uniform(-1, 1, P.nm))
        self.I_ch = 1e-3 * rng.uniform(-1, 1, (3, P.nm))
        self.splitting = splitting
        # library side
        ctx = _lib.Context(0, lib)
        ctx.set_mesh(mesh.coords, mesh.cells, region, mesh.facet_cells, surf.array(), mtags)
        Dtab = np.array([[D_sub[k][int(t)] for t in tags] for k in range(3)])
        ctx.set_params(F=P.F, R=P.R, T=P.T, C_M=P.C_M, C_phi=P.C_phi, dt=P.dt, tau_emi=P.tau, tau_knp=P.tau,
                       Lp=P.Lp, z=Z_PHYS, D=Dtab, rho=[0.0] * len(tags), splitting=splitting)
        for k in range(3):
            ctx.set_field(_lib.F_C, k, c_all[k])
            ctx.set_field(_lib.F_ICH, k, self.I_ch[k])
        for k in range(2):
            ctx.set_field(_lib.F_CN, k, self.c_n[k])
        ctx.set_field(_lib.F_PHI, 0, self.phi)
        ctx.set_field(_lib.F_PHIM, 0, self.phi_M)
        self.ctx = ctx


def rel_err(a, b):
    """max |a-b| relative to max |b| (entrywise parity measure for matrices/vectors)"""
    import scipy.sparse as sp
    if sp.issparse(a):
        d = (a - b).tocoo()
        num = np.abs(d.data).max() if d.nnz else 0.0
        den = np.abs(b.tocoo().data).max()
        return num / den
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def entrywise_failures(a, b, rtol=1e-12, row_floor=1e-14):
    """Entry-by-entry comparison (north_star: "entries must agree within 1e-12 relative"):
    an entry fails when |a - b| > rtol |b| + row_floor max|row of b|.  The row-relative floor is
    what round-off in the sum of a row's contributions allows for entries that cancel to (almost)
    zero.  Returns (number of failing entries, worst |a - b| / bound)."""
    import scipy.sparse as sp
    if sp.issparse(a) or sp.issparse(b):
        a, b = sp.csr_matrix(a), sp.csr_matrix(b)
        d = (a - b).tocoo()
        if d.nnz == 0:
            return 0, 0.0
        rowmax = np.asarray(abs(b).max(axis=1).todense()).ravel()
        bv = np.abs(np.asarray(b[d.row, d.col]).ravel())
        bound = rtol * bv + row_floor * rowmax[d.row]
        ratio = np.abs(d.data) / np.maximum(bound, 1e-300)
        return int((ratio > 1.0).sum()), float(ratio.max())
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    bound = rtol * np.abs(b) + row_floor * np.abs(b).max()
    ratio = np.abs(a - b) / np.maximum(bound, 1e-300)
    return int((ratio > 1.0).sum()), float(ratio.max()) if ratio.size else 0.0


def load_h5_series(path):
    """results.h5 written by Solver.save_h5 (reference layout, solver.py:1214-1242) -> dict with
    'potential' [nsteps+1, nc, nd], 'concentrations' [nsteps+1, N_ions, nc, nd], 'elim_concentration',
    'subdomains' [nc], 'surfaces' [nf], 'coordinates', 'topology' (vector_0 = state before the first step)"""
    from knpemidg import h5lite
    f = h5lite.File(str(path))
    cells = f["/mesh/topology"].read()
    nc, nd = cells.shape
    out = {"coordinates": f["/mesh/coordinates"].read(), "topology": cells,
           "subdomains": f["/subdomains/values"].read(), "surfaces": f["/surfaces/values"].read()}
    n = len([k for k in f["/potential"].keys() if k.startswith("vector_")])
    out["potential"] = np.stack([f[f"/potential/vector_{i}"].read().reshape(nc, nd) for i in range(n)])
    out["elim_concentration"] = np.stack([f[f"/elim_concentration/vector_{i}"].read().reshape(nc, nd) for i in range(n)])
    out["concentrations"] = np.stack([f[f"/concentrations/vector_{i}"].read().reshape(-1, nc, nd) for i in range(n)])
    return out
