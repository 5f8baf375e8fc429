"""ctypes binding of libknpemi.so (C ABI: include/knpemi.h).

The product library is `libknpemi.so` next to this file, built for sm_100a by
knp-emi-dg_b200/build.py.  There is no CPU fallback: `get()` raises if the
library is missing, and creating a context raises if no CUDA device is
present.  (tests/ load the host-emulation build explicitly through
`Lib(path)`; nothing in this package does.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libknpemi.so")

# field ids (include/knpemi.h)
F_C, F_CN, F_PHI, F_PHIM, F_ICH, F_NERNST, F_RHS_EMI, F_RHS_KNP, F_LOAD_EMI, F_LOAD_KNP = range(10)

POST_ELIMINATED, POST_PHIM, POST_NERNST, POST_ALL = 1, 2, 4, 7

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)
_ctx = C.c_void_p
# transport callbacks of the host-emulation build (include/knpemi.h knp_exchange_fn / knp_allreduce_fn)
XFN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _ip, _dp, _lp, _dp, _lp)
RFN = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int64)

_PROTOS = {
    "knp_last_error": (C.c_char_p, []),
    "knp_version": (C.c_int, []),
    "knp_is_cuda_build": (C.c_int, []),
    "knp_ctx_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "knp_ctx_destroy": (C.c_int, [_ctx]),
    "knp_sync": (C.c_int, [_ctx]),
    "knp_mesh_set": (C.c_int, [_ctx, C.c_int, C.c_int64, C.c_int64, _dp, _ip, _ip, C.c_int64, _ip, _ip,
                               C.c_int, _ip]),
    "knp_mesh_info": (C.c_int, [_ctx, _lp]),
    "knp_membrane_table": (C.c_int, [_ctx, _ip, _ip, _ip, _ip]),
    "knp_params_set": (C.c_int, [_ctx] + [C.c_double] * 9 + [C.c_int, _dp, C.c_int, _dp, _dp, _dp,
                                                           C.c_int, C.c_int]),
    "knp_field_set": (C.c_int, [_ctx, C.c_int, C.c_int, _dp, C.c_int64]),
    "knp_field_get": (C.c_int, [_ctx, C.c_int, C.c_int, _dp, C.c_int64]),
    "knp_assemble_emi": (C.c_int, [_ctx]),
    "knp_assemble_knp": (C.c_int, [_ctx]),
    "knp_matrix_export": (C.c_int, [_ctx, C.c_int, _lp, _ip, _dp]),
    "knp_spmv": (C.c_int, [_ctx, C.c_int, _dp, _dp]),
    "knp_amg_setup": (C.c_int, [_ctx, C.c_double, C.c_int, C.c_int]),
    "knp_amg_info": (C.c_int, [_ctx, _lp, _lp, _lp, C.c_int]),
    "knp_solver_options": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]),
    "knp_solve_emi": (C.c_int, [_ctx, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), _dp]),
    "knp_solve_knp": (C.c_int, [_ctx, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int), _dp]),
    "knp_post_step": (C.c_int, [_ctx, C.c_int]),
    "knp_facet_trace": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _dp]),
    "knp_model_count": (C.c_int, []),
    "knp_model_name": (C.c_char_p, [C.c_int]),
    "knp_model_dims": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "knp_membrane_register": (C.c_int, [_ctx, C.c_int, C.c_int64, _ip, _dp, _dp, C.POINTER(C.c_int)]),
    "knp_membrane_states_get": (C.c_int, [_ctx, C.c_int, _dp]),
    "knp_membrane_states_set": (C.c_int, [_ctx, C.c_int, _dp]),
    "knp_membrane_params_get": (C.c_int, [_ctx, C.c_int, _dp]),
    "knp_membrane_params_set": (C.c_int, [_ctx, C.c_int, _dp]),
    "knp_membrane_link": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "knp_membrane_outputs": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, _ip]),
    "knp_membrane_stimulus": (C.c_int, [_ctx, C.c_int, _bp, C.c_int, _ip, _dp]),
    "knp_ode_step": (C.c_int, [_ctx, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, _lp]),
    "knp_dist_set": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int64, C.c_int, _ip, _lp, _ip, _lp]),
    "knp_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "knp_dist_init_nccl": (C.c_int, [_ctx, C.c_char_p]),
    "knp_dist_set_callbacks": (C.c_int, [_ctx, XFN, RFN, C.c_void_p]),
    "knp_dist_info": (C.c_int, [_ctx, _lp]),
    "knp_field_halo": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "knp_timers_get": (C.c_int, [_ctx, _dp, C.c_int]),
    "knp_launch_count": (C.c_longlong, []),
    "knp_timer_start": (C.c_int, [_ctx]),
    "knp_timer_stop": (C.c_int, [_ctx, _dp]),
    "knp_bench_kernel": (C.c_int, [_ctx, C.c_int, C.c_int, _dp, _dp]),
}

SYMBOLS = tuple(_PROTOS)


class KnpError(RuntimeError):
    pass


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, typ):
    return a.ctypes.data_as(typ)


class Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise KnpError(f"{path} not found: build it with `python knp-emi-dg_b200/build.py` "
                           "(there is no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args

    def check(self, rc):
        if rc != 0:
            raise KnpError(self.dll.knp_last_error().decode())

    def is_cuda(self):
        return bool(self.dll.knp_is_cuda_build())

    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        self.check(self.dll.knp_nccl_unique_id(buf))
        return buf.raw

    def models(self):
        out = {}
        for i in range(self.dll.knp_model_count()):
            ns, npar = C.c_int(), C.c_int()
            self.check(self.dll.knp_model_dims(i, C.byref(ns), C.byref(npar)))
            out[self.dll.knp_model_name(i).decode()] = (i, ns.value, npar.value)
        return out


_instance = None


def get():
    """The product library (CUDA build).  Raises if it is not built."""
    global _instance
    if _instance is None:
        lib = Lib(LIB_PATH)
        if not lib.is_cuda():
            raise KnpError("libknpemi.so is not a CUDA build")
        _instance = lib
    return _instance


def variant_with(lib, user_modules):
    """A library that also carries the given user ODE modules (the reference accepts ANY module
    with the mm_*.py protocol, membrane.py:88; here its right-hand side has to become device
    code): the same sources are rebuilt with the modules' translated right-hand sides added
    (knp-emi-dg_b200/build.py:build_variant; nvcc for the CUDA library, g++ for the emulation
    build of the CPU tests; cached in knpemidg/_variants, git-ignored).  Returns (Lib, {module: name})."""
    import importlib.util
    path = os.path.join(os.path.dirname(_HERE), "build.py")
    spec = importlib.util.spec_from_file_location("knp_build", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    so, names = mod.build_variant(list(user_modules), emu=not lib.is_cuda())
    return Lib(so), names


class Context:
    """One device context; thin numpy-facing wrapper over the C ABI."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or get()
        self.h = _ctx()
        self.lib.check(self.lib.dll.knp_ctx_create(int(device), C.byref(self.h)))
        self.device = int(device)
        self.d = self.nc = self.n = self.nm = 0
        self.nc_owned = self.n_owned = 0
        self.rank, self.world = 0, 1
        self.N = 0

    def close(self):
        if self.h:
            self.lib.dll.knp_ctx_destroy(self.h)
            self.h = _ctx()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        self.lib.check(getattr(self.lib.dll, name)(self.h, *args))

    # -- mesh -------------------------------------------------------------
    def set_mesh(self, coords, cells, cell_region, facet_cells, facet_tag, mem_tags):
        coords = _f64(coords)
        cells = _i32(cells)
        d = coords.shape[1]
        assert cells.shape[1] == d + 1
        region = _i32(cell_region)
        fc = _i32(facet_cells)
        ft = _i32(facet_tag)
        mt = _i32(list(mem_tags))
        self._call("knp_mesh_set", d, cells.shape[0], coords.shape[0], _p(coords, _dp), _p(cells, _ip),
                   _p(region, _ip), fc.shape[0], _p(fc, _ip), _p(ft, _ip), len(mt), _p(mt, _ip))
        info = np.zeros(8, dtype=np.int64)
        self._call("knp_mesh_info", _p(info, _lp))
        self.d, self.nc, self.n, self.nm, self.nnz = (int(v) for v in info[:5])
        self.nd = self.d + 1
        self.nsip = int(info[5])
        self.nc_owned, self.n_owned = self.nc, self.n
        self.rank, self.world = 0, 1

    # -- multi-GPU --------------------------------------------------------
    def set_dist(self, rank, world, nc_owned, neigh, send_ptr, send_cells, recv_ptr):
        """declare this context one part of a cell-partitioned mesh (knp_dist_set)"""
        neigh = _i32(neigh)
        sp = np.ascontiguousarray(send_ptr, dtype=np.int64)
        sc = _i32(send_cells)
        rp = np.ascontiguousarray(recv_ptr, dtype=np.int64)
        self._call("knp_dist_set", int(rank), int(world), int(nc_owned), neigh.size, _p(neigh, _ip), _p(sp, _lp),
                   _p(sc, _ip), _p(rp, _lp))
        self.rank, self.world = int(rank), int(world)
        self.nc_owned, self.n_owned = int(nc_owned), int(nc_owned) * self.nd
        info = np.zeros(8, dtype=np.int64)
        self._call("knp_mesh_info", _p(info, _lp))
        self.nnz = int(info[4])

    def init_nccl(self, uid):
        assert len(uid) == 128
        self._call("knp_dist_init_nccl", C.create_string_buffer(bytes(uid), 128))

    def set_callbacks(self, exchange, allreduce):
        self._call("knp_dist_set_callbacks", exchange, allreduce, None)

    def dist_info(self):
        info = np.zeros(8, dtype=np.int64)
        self._call("knp_dist_info", _p(info, _lp))
        return dict(zip(("rank", "world", "owned_cells", "ghost_cells", "neighbours", "halos", "allreduces",
                         "peer_memory_ops"), (int(v) for v in info[:8])))

    def field_halo(self, which, idx=0):
        self._call("knp_field_halo", int(which), int(idx))

    def membrane_table(self):
        out = [np.zeros(self.nm, dtype=np.int32) for _ in range(4)]
        self._call("knp_membrane_table", *[_p(a, _ip) for a in out])
        return dict(facet=out[0], cell_i=out[1], cell_e=out[2], tag=out[3])

    def set_params(self, *, F, R, T, C_M, C_phi, dt, tau_emi, tau_knp, Lp, z, D, rho=None, C_sub=None,
                   splitting=True, mms=False):
        z = _f64(z)
        D = _f64(D)
        N, ntags = D.shape
        assert len(z) == N
        rho = _f64(np.zeros(ntags) if rho is None else rho)
        cs = None if C_sub is None else _f64(C_sub)
        self._call("knp_params_set", F, R, T, C_M, C_phi, dt, tau_emi, tau_knp, Lp, N, _p(z, _dp), ntags,
                   _p(D, _dp), _p(rho, _dp), None if cs is None else _p(cs, _dp), int(bool(splitting)),
                   int(bool(mms)))
        self.N = N

    # -- fields -----------------------------------------------------------
    def _count(self, which):
        return self.nm if which in (F_PHIM, F_ICH, F_NERNST) else self.n

    def set_field(self, which, idx, values):
        v = _f64(values).ravel()
        self._call("knp_field_set", which, idx, _p(v, _dp), v.size)

    def get_field(self, which, idx=0, out=None):
        """copy a field to the host; `out` (contiguous float64, e.g. a pinned buffer) receives it
        directly when given"""
        n = self._count(which)
        if out is None:
            out = np.empty(n)
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size != n:
            raise ValueError("get_field: out must be a contiguous float64 array of the field's size")
        self._call("knp_field_get", which, idx, _p(out, _dp), n)
        return out

    # -- assembly ---------------------------------------------------------
    def assemble_emi(self):
        self._call("knp_assemble_emi")

    def assemble_knp(self):
        self._call("knp_assemble_knp")

    def matrix(self, which):
        """scipy CSR copy of matrix `which` (0 A_emi, 1 B_emi, 2+k A_knp[k])."""
        import scipy.sparse as sp
        ptr = np.zeros(self.n_owned + 1, dtype=np.int64)
        col = np.zeros(self.nnz, dtype=np.int32)
        val = np.zeros(self.nnz)
        self._call("knp_matrix_export", which, _p(ptr, _lp), _p(col, _ip), _p(val, _dp))
        # rows: owned dofs; columns: local dofs (owned + ghost)
        return sp.csr_matrix((val, col, ptr), shape=(self.n_owned, self.n))

    def spmv(self, which, x):
        x = _f64(x).ravel()
        y = np.empty_like(x)
        self._call("knp_spmv", which, _p(x, _dp), _p(y, _dp))
        return y

    # -- solvers ----------------------------------------------------------
    def amg_setup(self, theta=0.08, max_levels=12, coarse_size=64):
        self._call("knp_amg_setup", float(theta), int(max_levels), int(coarse_size))

    def amg_info(self):
        nl = C.c_int64()
        rows = np.zeros(32, dtype=np.int64)
        nnz = np.zeros(32, dtype=np.int64)
        self._call("knp_amg_info", C.byref(nl), _p(rows, _lp), _p(nnz, _lp), 32)
        k = int(nl.value)
        return list(map(int, rows[:k])), list(map(int, nnz[:k]))

    def solver_options(self, pc=1, nu_pre=1, nu_post=1, gamma=1, omega=0.0, restart=30, knp_min_it=5):
        self._call("knp_solver_options", pc, nu_pre, nu_post, gamma, float(omega), restart, knp_min_it)

    def solve_emi(self, rtol=1e-5, atol=1e-40, maxit=1000):
        it, res = C.c_int(), C.c_double()
        self._call("knp_solve_emi", rtol, atol, maxit, C.byref(it), C.byref(res))
        return it.value, res.value

    def solve_knp(self, rtol=1e-7, atol=1e-40, maxit=1000):
        it, res = C.c_int(), C.c_double()
        self._call("knp_solve_knp", rtol, atol, maxit, C.byref(it), C.byref(res))
        return it.value, res.value

    def post_step(self, what=POST_ALL):
        self._call("knp_post_step", int(what))

    def facet_trace(self, which, idx, side):
        out = np.empty(self.nm)
        self._call("knp_facet_trace", which, idx, int(side), _p(out, _dp))
        return out

    # -- membranes --------------------------------------------------------
    def membrane_register(self, model_id, rows, states, params):
        rows = _i32(rows)
        states = _f64(states)
        params = _f64(params)
        ns, npar = self._model_dims(model_id)
        if states.size != rows.size * ns or params.size != rows.size * npar:
            raise KnpError(f"membrane_register: model {model_id} needs states [{rows.size}, {ns}] and parameters "
                           f"[{rows.size}, {npar}], got {states.shape} and {params.shape}")
        h = C.c_int()
        self._call("knp_membrane_register", model_id, rows.size, _p(rows, _ip), _p(states, _dp),
                   _p(params, _dp), C.byref(h))
        self._mem_shapes = getattr(self, "_mem_shapes", {})
        self._mem_shapes[h.value] = {"states": rows.size * ns, "params": rows.size * npar}
        return h.value

    def _model_dims(self, model_id):
        for _, (mid, ns, npar) in self.lib.models().items():
            if mid == model_id:
                return ns, npar
        raise KnpError(f"unknown membrane model id {model_id}")

    def membrane_get(self, handle, what, shape):
        out = np.empty(shape)
        want = getattr(self, "_mem_shapes", {}).get(handle, {}).get(what)
        if want is not None and out.size != want:
            raise KnpError(f"membrane_get({what}): shape {shape} for a table of {want} values")
        self._call("knp_membrane_%s_get" % what, handle, _p(out, _dp))
        return out

    def membrane_set(self, handle, what, values):
        v = _f64(values)
        want = getattr(self, "_mem_shapes", {}).get(handle, {}).get(what)
        if want is not None and v.size != want:
            raise KnpError(f"membrane_set({what}): {v.size} values for a table of {want}")
        self._call("knp_membrane_%s_set" % what, handle, _p(v, _dp))

    def membrane_link(self, handle, col, kind, which, idx=0, side=0):
        self._call("knp_membrane_link", handle, col, kind, which, idx, side)

    def membrane_outputs(self, handle, v_col, ich_cols):
        cols = _i32(ich_cols)
        self._call("knp_membrane_outputs", handle, v_col, cols.size, _p(cols, _ip))

    def membrane_stimulus(self, handle, mask, cols, values):
        cols = _i32(cols)
        values = _f64(values)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._call("knp_membrane_stimulus", handle, None if m is None else _p(m, _bp), cols.size,
                   _p(cols, _ip), _p(values, _dp))

    def ode_step(self, handle, t0, dt, rtol=1e-8, atol=0.0, set_v=True):
        stats = np.zeros(3, dtype=np.int64)
        self._call("knp_ode_step", handle, t0, dt, rtol, atol, int(bool(set_v)), _p(stats, _lp))
        self.ode_stiff_facets = int(stats[2])          # facets that finished the interval on the implicit pair
        return int(stats[0]), int(stats[1])

    def timers(self, reset=False):
        out = np.zeros(6)
        self._call("knp_timers_get", _p(out, _dp), int(reset))
        return dict(zip(("emi_assemble", "emi_solve", "knp_assemble", "knp_solve", "ode", "post"), out))

    def sync(self):
        self._call("knp_sync")

    # -- measurement hooks ---------------------------------------------------
    def launch_count(self):
        return int(self.lib.dll.knp_launch_count())

    def timer_start(self):
        self._call("knp_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._call("knp_timer_stop", C.byref(ms))
        return ms.value

    def bench_kernel(self, kernel, reps=20):
        ms, nbytes = C.c_double(), C.c_double()
        self._call("knp_bench_kernel", int(kernel), int(reps), C.byref(ms), C.byref(nbytes))
        return ms.value, nbytes.value
