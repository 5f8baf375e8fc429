"""One rank of a partitioned KNP-EMI run (world_size >= 2), used by

* tests/test_dist_gloo.py  - CPU: gloo + the host-emulation library (callback transport);
* tests/test_gpu_dist.py   - GPU: NCCL + libknpemi.so, one process per GPU (torchrun).

Every rank builds the same global mesh, takes its part, steps the engine and gathers the
global fields; rank 0 compares them with a single-part run of the same library and writes
the verdict as JSON to `out`.

    python tests/dist_worker.py <kind: emu|gpu> <case> <nsteps> <out.json>
(rank / world / rendezvous from the torchrun-style environment variables)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "knp-emi-dg_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def build(case, lib, device, transport, tight=True):
    import bench
    from knpemidg import mesh as kmesh
    from knpemidg.engine import Engine
    from knpemidg.models import mm_hh, mm_hh_no_stim
    if case == "bundle":
        mesh, sub, surf = kmesh.bundle_3d_mesh(dims=(16, 9, 9))
        mtags, models = (1, 2), {1: mm_hh, 2: mm_hh_no_stim}
        phys, cinit, names = bench.PHYS, bench.C_INIT, bench.ION_NAMES
    elif case in ("bundle_r0", "bundle_r0_quad"):
        mesh, sub, surf = kmesh.bundle_3d_mesh(0)
        mtags, models = (1, 2), {1: mm_hh, 2: mm_hh_no_stim}
        phys, cinit, names = bench.PHYS, bench.C_INIT, bench.ION_NAMES
    elif case == "neuron2d":
        mesh, sub, surf = kmesh.neuron_2d_mesh(1)
        mtags, models = (1,), {1: mm_hh}
        phys, cinit, names = bench.PHYS, bench.C_INIT, bench.ION_NAMES
    elif case == "unstr3d":
        # jittered tetrahedra, random vertex / cell / local-vertex numbering (tests/common.py): the
        # general graph partitioner, irregular halos and membrane facets cut by part boundaries
        from common import unstructured_mesh
        mesh, sub, surf = unstructured_mesh(3, n=8)
        mtags, models = (1,), {1: mm_hh}
        phys, cinit, names = bench.PHYS, bench.C_INIT, bench.ION_NAMES
    else:
        raise ValueError(case)
    part = None
    if case.endswith("_quad") and transport is not None:
        # 2 x 2 split in (x, y): every part has three neighbours, two of them through an edge only
        # as far as faces go - exercises multi-neighbour halos and cut membranes
        mid = mesh.cell_midpoints()
        half = 0.5 * (mesh.coords.max(axis=0) + mesh.coords.min(axis=0))
        part = (2 * (mid[:, 0] > half[0]) + (mid[:, 1] > half[1])).astype(np.int32) % transport.world
    eng = Engine(mesh, sub.array(), surf.array(), membrane_tags=mtags, device=device, lib=lib,
                 transport=transport, part=part, **phys)
    eng.set_concentrations_by_tag(cinit)
    for tag, mod in models.items():
        eng.add_membrane_model(tag, mod, names, stimulus=bench.STIMULUS, stimulus_locator=bench.stim_locator)
    if tight:
        eng.rtol_emi, eng.rtol_knp = 1e-11, 1e-12
    eng.initialize(pc=1)
    return eng


def main():
    kind, case, nsteps, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    import torch
    import torch.distributed as dist
    from common import lib_for, rel_err
    from knpemidg.partition import TorchTransport
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    lib = lib_for(kind)
    if kind == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        obj_group = dist.new_group(backend="gloo")
    else:
        dist.init_process_group("gloo")
        obj_group = None
    tr = TorchTransport()
    device = local_rank if kind == "gpu" else 0
    eng = build(case, lib, device, tr)
    if obj_group is not None:                      # gather python objects over gloo, not NCCL
        tr.group = obj_group
    for _ in range(nsteps):
        eng.step()
    got = {"phi": eng.phi(gather=True), "phi_M": eng.phi_M(gather=True)}
    for k in range(3):
        got[f"c{k}"] = eng.concentration(k, gather=True)
    info = eng.ctx.dist_info()
    its = dict(eng.stats)
    levels = eng.ctx.amg_info()
    if rank == 0:
        ref = build(case, lib, device, None)
        for _ in range(nsteps):
            ref.step()
        want = {"phi": ref.phi(), "phi_M": ref.phi_M()}
        for k in range(3):
            want[f"c{k}"] = ref.concentration(k)
        errs = {}
        for key in got:
            a, b = got[key], want[key]
            if key == "phi":                       # pure Neumann: defined up to a constant
                a, b = a - a.mean(), b - b.mean()
            errs[key] = float(rel_err(a, b))
        res = {"errs": errs, "iterations": its, "ref_iterations": dict(ref.stats), "dist": info,
               "levels": levels, "ref_levels": ref.ctx.amg_info(),
               "phi_M_range": [float(want["phi_M"].min()), float(want["phi_M"].max())]}
        with open(out, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
