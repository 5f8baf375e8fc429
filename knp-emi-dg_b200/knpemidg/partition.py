"""Host-side cell partition of the mesh for the multi-GPU path (setup only, numpy).

The reference distributes the problem with dolfin's mesh partitioner when run_*.py is
started under mpirun (`parameters['ghost_mode'] = 'shared_vertex'`,
src/knpemidg/solver.py:16); PETSc then owns the row blocks of the matrices and the
VecScatter of every MatMult.  Here:

* `partition_cells`  assigns every cell to one of `nparts` parts: recursive coordinate
  bisection (balanced by cell count) where each cut is made along the axis that severs the
  smallest total COUPLING WEIGHT of the cell-facet dual graph (|F| / midpoint distance),
  followed by boundary refinement sweeps on the cell-facet dual graph that move a cell to
  the part most of its face neighbours belong to when that lowers the edge cut and keeps
  the balance;
* `local_part`       builds what one rank uploads with `knp_mesh_set` + `knp_dist_set`:
  its owned cells (ascending global index) followed by the ghost cells (face neighbours
  owned by other parts, grouped by owner, ascending global index), the facets with at
  least one owned cell (ascending global facet index, so membrane rows keep the
  reference's ODE-point order, dlt_dof_extraction.py:18-48), and the send/recv lists of
  the DG halo.

A membrane facet whose two cells are owned by different ranks is a membrane row on BOTH
ranks (each integrates its ODE point redundantly from identical inputs, so no membrane
data is ever exchanged); `owned_rows` marks the copy that counts in global outputs (the
rank that owns the ICS-side cell).
"""
from __future__ import annotations

import numpy as np

from .mesh import SimplexMesh


def _facet_weights(mesh):
    """coupling strength of every interior facet in the DG operators: |F| / distance of the two
    cell midpoints (two-point-flux transmissibility).  Cutting strong couplings costs Krylov
    iterations - the AMG aggregates stop at partition boundaries - so the bisection below cuts
    where the sum of these weights is smallest, not simply across the longest extent."""
    fc = mesh.facet_cells
    interior = np.flatnonzero(fc[:, 1] >= 0)
    X = mesh.coords[mesh.facet_verts[interior]]
    if mesh.gdim == 2:
        area = np.linalg.norm(X[:, 1] - X[:, 0], axis=1)
    else:
        area = 0.5 * np.linalg.norm(np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), axis=1)
    mid = mesh.cell_midpoints()
    a, b = fc[interior, 0], fc[interior, 1]
    dist = np.linalg.norm(mid[a] - mid[b], axis=1)
    return a, b, area / np.maximum(dist, 1e-300)


def _bisect(mid, idx, nparts, first, out, graph=None):
    if nparts == 1:
        out[idx] = first
        return
    left = nparts // 2
    cut = (len(idx) * left) // nparts
    best = None
    inset = None
    if graph is not None:
        inset = np.zeros(mid.shape[0], dtype=np.int8)
        inset[idx] = 1
    for axis in range(mid.shape[1]):
        order = idx[np.argsort(mid[idx, axis], kind="stable")]
        if graph is None:
            ext = mid[idx, axis].max() - mid[idx, axis].min()
            cost = -ext                                   # no graph: cut across the longest extent
        else:
            a, b, w = graph
            side = inset.copy()
            side[order[cut:]] = 2
            sel = (side[a] != 0) & (side[b] != 0) & (side[a] != side[b])
            cost = float(w[sel].sum())
        if best is None or cost < best[0]:
            best = (cost, order)
    order = best[1]
    _bisect(mid, order[:cut], left, first, out, graph)
    _bisect(mid, order[cut:], nparts - left, first + left, out, graph)


def _refine(part, fc, nparts, sweeps=4, imbalance=1.03):
    """greedy boundary refinement: move a cell to the neighbouring part that holds more of
    its face neighbours than its own part does (strict gain), while no part grows beyond
    `imbalance` x the average"""
    interior = fc[:, 1] >= 0
    a, b = fc[interior, 0], fc[interior, 1]
    nc = part.size
    cap = int(np.ceil(imbalance * nc / nparts))
    for _ in range(sweeps):
        pa, pb = part[a], part[b]
        cut = pa != pb
        if not cut.any():
            break
        # number of neighbours of each cell in each part (only for boundary cells)
        bcells = np.unique(np.concatenate([a[cut], b[cut]]))
        pos = -np.ones(nc, dtype=np.int64)
        pos[bcells] = np.arange(bcells.size)
        cnt = np.zeros((bcells.size, nparts), dtype=np.int32)
        sel = pos[a] >= 0
        np.add.at(cnt, (pos[a[sel]], pb[sel]), 1)
        sel = pos[b] >= 0
        np.add.at(cnt, (pos[b[sel]], pa[sel]), 1)
        own = cnt[np.arange(bcells.size), part[bcells]]
        best = cnt.argmax(axis=1)
        gain = cnt[np.arange(bcells.size), best] - own
        size = np.bincount(part, minlength=nparts)
        moved = 0
        for k in np.argsort(-gain, kind="stable"):
            if gain[k] <= 0:
                break
            c, q = bcells[k], best[k]
            if size[q] + 1 > cap:
                continue
            # neighbours may have moved in this sweep: recompute the gain exactly
            size[part[c]] -= 1
            part[c] = q
            size[q] += 1
            moved += 1
        if moved == 0:
            break
    return part


def partition_cells(mesh, nparts, refine=True):
    """part[cell] in 0..nparts-1"""
    mesh.init_topology()
    nc = mesh.cells.shape[0]
    part = np.zeros(nc, dtype=np.int32)
    if nparts <= 1:
        return part
    mid = mesh.cell_midpoints()
    _bisect(mid, np.arange(nc), int(nparts), 0, part, _facet_weights(mesh))
    if refine:
        cut0 = edge_cut(mesh, part)
        trial = _refine(part.copy(), mesh.facet_cells, int(nparts))
        if edge_cut(mesh, trial) < cut0:
            part = trial
    return part


def edge_cut(mesh, part):
    fc = mesh.facet_cells
    interior = fc[:, 1] >= 0
    return int(np.count_nonzero(part[fc[interior, 0]] != part[fc[interior, 1]]))


class LocalPart:
    """One rank's share of a partitioned mesh (see module docstring)."""

    def __init__(self, mesh, cell_tags, facet_tags, part, rank):
        mesh.init_topology()
        part = np.asarray(part)
        rank = int(rank)
        fc = mesh.facet_cells
        nc = mesh.cells.shape[0]
        self.rank, self.world = rank, int(part.max()) + 1
        owned = np.flatnonzero(part == rank)
        if owned.size == 0:
            raise ValueError(f"part {rank} owns no cells")
        interior = fc[:, 1] >= 0
        a, b = fc[interior, 0].astype(np.int64), fc[interior, 1].astype(np.int64)
        pa, pb = part[a], part[b]
        ghosts = np.unique(np.concatenate([b[(pa == rank) & (pb != rank)], a[(pb == rank) & (pa != rank)]]))
        ghosts = ghosts[np.lexsort((ghosts, part[ghosts]))]
        gowner = part[ghosts]
        self.neigh = np.unique(gowner).astype(np.int32)
        self.recv_ptr = np.searchsorted(gowner, np.append(self.neigh, self.world + 1)).astype(np.int64)
        self.recv_ptr[-1] = ghosts.size
        # owned cells: interior ones first, the ones that touch another part last - the library runs the
        # rows of the interior cells while the halo of a vector is still in flight (csrc/knp_solve.cu)
        bnd = np.unique(np.concatenate([a[(pa == rank) & (pb != rank)], b[(pb == rank) & (pa != rank)]]))
        owned = np.concatenate([np.setdiff1d(owned, bnd, assume_unique=True), bnd])
        self.nc_interior = int(owned.size - bnd.size)
        self.nc_owned = int(owned.size)
        self.l2g = np.concatenate([owned, ghosts]).astype(np.int64)
        g2l = -np.ones(nc + 1, dtype=np.int64)            # index -1 (no cell) maps to -1
        g2l[self.l2g] = np.arange(self.l2g.size)
        self.g2l = g2l[:nc]
        # send lists: my owned cells that touch neighbour q, ascending (= q's ghost order)
        send_cells, send_ptr = [], [0]
        for q in self.neigh:
            mine = np.unique(np.concatenate([a[(pa == rank) & (pb == q)], b[(pb == rank) & (pa == q)]]))
            send_cells.append(g2l[mine])
            send_ptr.append(send_ptr[-1] + mine.size)
        self.send_cells = (np.concatenate(send_cells) if send_cells else np.zeros(0)).astype(np.int32)
        self.send_ptr = np.asarray(send_ptr, dtype=np.int64)
        # facets with at least one owned cell, ascending global index
        own0 = part[fc[:, 0]] == rank
        own1 = interior & (part[np.maximum(fc[:, 1], 0)] == rank)
        self.facets = np.flatnonzero(own0 | own1)
        lfc = g2l[fc[self.facets].astype(np.int64)].astype(np.int32)
        assert (lfc[:, 0] >= 0).all() and (lfc[fc[self.facets, 1] >= 0, 1] >= 0).all()
        lm = SimplexMesh(mesh.coords, mesh.cells[self.l2g])
        lm.facet_cells = lfc
        lm.facet_verts = mesh.facet_verts[self.facets]
        lm.facet_local = mesh.facet_local[self.facets]
        lm._topo = True
        self.mesh = lm
        self.cell_tags = np.asarray(cell_tags)[self.l2g]
        self.facet_tags = np.asarray(facet_tags)[self.facets]
        self.part = part

    def dist_args(self):
        return dict(rank=self.rank, world=self.world, nc_owned=self.nc_owned, neigh=self.neigh,
                    send_ptr=self.send_ptr, send_cells=self.send_cells, recv_ptr=self.recv_ptr)

    # -- global <-> local field helpers -------------------------------------
    def to_local_cells(self, values):
        """per-cell (or per-dof [nc, nd]) global array -> local (owned + ghost)"""
        return np.asarray(values)[self.l2g]

    def owned_global_cells(self):
        return self.l2g[: self.nc_owned]


class TorchTransport:
    """Plumbing between the ranks of a torch.distributed job and the C library.

    * NCCL (product): rank 0 creates the id, it is broadcast as a byte tensor, every rank
      calls knp_dist_init_nccl.  All data-path traffic then happens inside libknpemi.so.
    * gloo (CPU tests of the emulation build): the library's exchange/allreduce callbacks
      are served with torch.distributed point-to-point ops on host arrays.
    """

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._keep = []

    def attach(self, ctx):
        import ctypes as C
        import torch
        from . import _lib
        dist = self.dist
        if ctx.lib.is_cuda():
            buf = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = torch.frombuffer(bytearray(ctx.lib.nccl_unique_id()), dtype=torch.uint8).clone()
            dev = torch.device("cuda", ctx.device) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            t = buf.to(dev)
            dist.broadcast(t, src=0, group=self.group)
            ctx.init_nccl(bytes(t.cpu().numpy().tobytes()))
            return

        def exchange(user, nn, ranks, send, send_off, recv, recv_off):
            try:
                reqs, bufs = [], []
                for i in range(nn):
                    q = int(ranks[i])
                    ns = int(send_off[i + 1] - send_off[i])
                    nr = int(recv_off[i + 1] - recv_off[i])
                    if ns > 0:
                        src = np.ctypeslib.as_array(send, shape=(int(send_off[nn]),))[int(send_off[i]): int(send_off[i + 1])]
                        reqs.append(dist.isend(torch.from_numpy(src.copy()), dst=q, group=self.group))
                    if nr > 0:
                        t = torch.empty(nr, dtype=torch.float64)
                        bufs.append((t, int(recv_off[i]), nr))
                        reqs.append(dist.irecv(t, src=q, group=self.group))
                for r in reqs:
                    r.wait()
                if bufs:
                    out = np.ctypeslib.as_array(recv, shape=(int(recv_off[nn]),))
                    for t, off, nr in bufs:
                        out[off: off + nr] = t.numpy()
                return 0
            except Exception as e:  # pragma: no cover - surfaced through knp_last_error
                print("knpemidg transport: exchange failed:", e, flush=True)
                return 1

        def allreduce(user, buf, n):
            try:
                a = np.ctypeslib.as_array(buf, shape=(int(n),))
                t = torch.from_numpy(a.copy())
                dist.all_reduce(t, group=self.group)
                a[:] = t.numpy()
                return 0
            except Exception as e:  # pragma: no cover
                print("knpemidg transport: allreduce failed:", e, flush=True)
                return 1

        x, r = _lib.XFN(exchange), _lib.RFN(allreduce)
        self._keep += [x, r]
        ctx.set_callbacks(x, r)

    def gather_owned(self, local_owned, owned_ids, n_global):
        """assemble a global per-entity array from every rank's owned part (tests, output)"""
        import torch
        dist = self.dist
        items = [None] * self.world
        dist.all_gather_object(items, (np.asarray(owned_ids), np.asarray(local_owned)), group=self.group)
        first = items[0][1]
        out = np.zeros((n_global,) + first.shape[1:], dtype=first.dtype)
        for ids, vals in items:
            out[ids] = vals
        return out
