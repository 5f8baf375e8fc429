"""Adapter from live dolfin objects to the flat arrays the library takes.  Used once at
setup and only where legacy FEniCS exists (it does not in this image, so this module is
NOT exercised by the test-suite).  north_star: "UFL/dolfin is used
only once at setup to extract mesh connectivity, facet markers".
"""
import numpy as np

from .mesh import SimplexMesh


def from_dolfin(mesh, subdomains, surfaces):
    coords = np.array(mesh.coordinates(), dtype=float)
    cells = np.array(mesh.cells(), dtype=np.int32)
    sm = SimplexMesh(coords, cells)
    sm.init_topology()
    d = sm.gdim
    mesh.init(d - 1, 0)
    f2v = np.array(mesh.topology()(d - 1, 0)(), dtype=np.int64).reshape(-1, d)
    f2v.sort(axis=1)
    nv = coords.shape[0]
    key = f2v[:, 0]
    for a in range(1, d):
        key = key * nv + f2v[:, a]
    mine = sm.facet_verts.astype(np.int64)
    mkey = mine[:, 0]
    for a in range(1, d):
        mkey = mkey * nv + mine[:, a]
    order = np.argsort(key)
    pos = order[np.searchsorted(key[order], mkey)]       # dolfin facet index of each of our facets
    facet_tags = np.asarray(surfaces.array())[pos]
    return sm, np.asarray(subdomains.array()), facet_tags
