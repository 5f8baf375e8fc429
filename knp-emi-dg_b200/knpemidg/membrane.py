"""MembraneModel with the reference's interface (src/knpemidg/membrane.py:7-184) over the
device-resident ODE tables of libknpemi.so.

`states` / `parameters` are numpy *copies* fetched from the device on access (the
reference exposes the live tables); assignments go back through `set_*`.  The integration
itself (step_lsoda) is one kernel launch: knp_ode_step (one thread per membrane facet).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .frontend import FacetField, FacetMean, HostFacetField


def _table_engine(ode, n, lib=None):
    """Device context for FREE-STANDING ODE tables (MembraneModel without a Solver, as in
    examples/emix-simulations/run_calibration.py:13-22): the C ABI registers tables against the
    membrane rows of a mesh, so a strip of 2 x n squares whose mid line has exactly n interface
    facets carries them.  ODE modules without a compiled counterpart get a library variant first."""
    from . import mesh as kmesh
    from .engine import Engine, find_compiled_model
    lib, names = lib or _lib.get(), {}
    if find_compiled_model(lib, ode) is None:
        print(f"knpemidg: compiling membrane model '{ode.__name__}' into a library variant ...")
        lib, names = _lib.variant_with(lib, [ode])
    m = kmesh.rectangle_mesh((0.0, 0.0), (float(n), 2.0), n, 2)
    m.init_topology()
    ctags = (m.cell_midpoints()[:, 1] > 1.0).astype(np.int64)
    fc = m.facet_cells
    ftags = ((fc[:, 1] >= 0) & (ctags[fc[:, 0]] != ctags[np.maximum(fc[:, 1], 0)])).astype(np.int64)
    one = {0: 1.0, 1: 1.0}
    eng = Engine(m, ctags, ftags, F=1.0, R=1.0, T=1.0, C_M=1.0, C_phi=1.0, dt=1.0, z=[1.0, -1.0],
                 D_sub=[one, one], membrane_tags=(1,), lib=lib)
    eng.user_models.update(names)
    assert eng.nm == n
    return eng


class _LazyTable:
    """what step_lsoda returns: the state table, downloaded from the device only if it is looked at"""

    def __init__(self, model):
        self._model = model

    def __array__(self, dtype=None, copy=None):
        a = self._model.states
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        return self._model.states[key]

    def __len__(self):
        return self._model.nodes

    @property
    def shape(self):
        return (self._model.nodes, len(self._model.ode.init_state_values()))


class MembraneModel:
    def __init__(self, ode, facet_f, tag, V):
        """facets with facet_f == tag are governed by `ode`; V is the solver's Q space
        (it carries the engine that owns the device context) or, for free-standing tables, any
        facet space of facet_f's mesh."""
        assert isinstance(tag, int)
        engine = getattr(V, "engine", None)
        standalone = engine is None
        if standalone:
            mesh = facet_f.mesh()
            mesh.init_topology()
            facets = np.flatnonzero(np.asarray(facet_f.array()) == tag)
            engine = _table_engine(ode, len(facets), getattr(V, "lib", None))
        self.engine, self.V = engine, V
        self.ode, self.tag = ode, tag
        self.prefix = ode.__name__
        lib = engine.ctx.lib
        self._model_name = engine.resolve_model(ode)
        self.model_id, self.ns, self.np_ = lib.models()[self._model_name]
        if standalone:
            rows = np.arange(len(facets), dtype=np.int32)
            self.facets = facets
            self.dof_locations = mesh.facet_midpoints()[facets]
        else:
            rows = np.flatnonzero(engine.mem["tag"] == tag).astype(np.int32)   # ascending facet index
            self.facets = engine.mem["facet"][rows]
            self.dof_locations = engine.membrane_midpoints()[rows]
        self.indices = rows                                   # rows of Q used by this model
        self.nodes = len(rows)
        states = np.array([ode.init_state_values() for _ in range(self.nodes)], dtype=float).reshape(self.nodes, self.ns)
        params = np.array([ode.init_parameter_values() for _ in range(self.nodes)], dtype=float).reshape(self.nodes, self.np_)
        self.handle = engine.ctx.membrane_register(self.model_id, rows, states, params)
        try:
            v_col = ode.state_indices("V")
        except ValueError:                                    # mm_calibration: V_n, V_g - no PDE coupling
            if not standalone:
                raise
            v_col = 0
        engine.ctx.membrane_outputs(self.handle, v_col, [])
        self.time = 0
        self._set_v = False
        self._stim_key = None

    # -- tables ----------------------------------------------------------------
    @property
    def states(self):
        return self.engine.ctx.membrane_get(self.handle, "states", (self.nodes, self.ns))

    @states.setter
    def states(self, values):
        self.engine.ctx.membrane_set(self.handle, "states", values)

    @property
    def parameters(self):
        return self.engine.ctx.membrane_get(self.handle, "params", (self.nodes, self.np_))

    @parameters.setter
    def parameters(self, values):
        self.engine.ctx.membrane_set(self.handle, "params", values)

    @property
    def V_index(self):
        return self.ode.state_indices("V")

    def _mask(self, locator):
        if locator is None:
            return np.ones(self.nodes, dtype=bool)
        return np.fromiter(map(locator, self.dof_locations), dtype=bool, count=self.nodes)

    # -- PDE -> ODE ------------------------------------------------------------
    def _set_ODE(self, what, which, u, locator=None):
        col = (self.ode.state_indices if what == "state" else self.ode.parameter_indices)(which)
        ctx = self.engine.ctx
        if locator is None and what == "parameter" and isinstance(u, FacetMean):
            f = u.trace.field                          # evaluated inside the ODE kernel each step
            ctx.membrane_link(self.handle, col, 1, f.which, f.idx, u.trace.side)
            return None
        if locator is None and what == "parameter" and isinstance(u, FacetField):
            ctx.membrane_link(self.handle, col, 0, u.which, u.idx, 0)
            return None
        if locator is None and what == "state" and which == "V" and isinstance(u, FacetField) \
                and u.which == _lib.F_PHIM:
            self._set_v = True                         # gather phi_M at the start of the next step
            return None
        src = u.vector().get_local()
        m = self._mask(locator)
        table = self.states if what == "state" else self.parameters
        table[m, col] = src[self.indices[m]]
        ctx.membrane_set(self.handle, "states" if what == "state" else "params", table)
        return table

    def set_state(self, which, u, locator=None):
        return self._set_ODE("state", which, u, locator)

    def set_parameter(self, which, u, locator=None):
        return self._set_ODE("parameter", which, u, locator)

    def set_membrane_potential(self, u, locator=None):
        return self.set_state("V", u, locator)

    # -- ODE -> PDE ------------------------------------------------------------
    def _get_PDE(self, what, which, u, locator=None):
        col = (self.ode.state_indices if what == "state" else self.ode.parameter_indices)(which)
        table = self.states if what == "state" else self.parameters
        m = self._mask(locator)
        dest = u.vector().get_local()
        dest[self.indices[m]] = table[m, col]
        u.vector().set_local(dest)
        return u

    def get_state(self, which, u, locator=None):
        return self._get_PDE("state", which, u, locator)

    def get_parameter(self, which, u, locator=None):
        return self._get_PDE("parameter", which, u, locator)

    def get_membrane_potential(self, u, locator=None):
        return self.get_state("V", u, locator)

    # -- constants -------------------------------------------------------------
    def _set_ODE_values(self, what, value_dict, locator=None):
        table = self.states if what == "state" else self.parameters
        get_col = self.ode.state_indices if what == "state" else self.ode.parameter_indices
        m = np.flatnonzero(self._mask(locator))
        for key, fn in value_dict.items():
            col = get_col(key)
            for row in m:
                table[row, col] = fn(self.dof_locations[row])
        self.engine.ctx.membrane_set(self.handle, "states" if what == "state" else "params", table)
        return table

    def set_state_values(self, value_dict, locator=None):
        return self._set_ODE_values("state", value_dict, locator)

    def set_parameter_values(self, value_dict, locator=None):
        return self._set_ODE_values("parameter", value_dict, locator)

    # -- integration -----------------------------------------------------------
    def step_lsoda(self, dt, stimulus, stimulus_locator=None):
        """advance all ODE points by dt (membrane.py:84-119) with the PDE->ODE links
        registered since the last step; returns the new state table."""
        ctx = self.engine.ctx
        key = (tuple(sorted((stimulus or {}).items())), id(stimulus_locator))
        if key != self._stim_key:
            if stimulus:
                mask = self._mask(stimulus_locator).astype(np.uint8)
                cols = [self.ode.parameter_indices(k) for k in stimulus]
                ctx.membrane_stimulus(self.handle, mask, cols, [float(v) for v in stimulus.values()])
            else:
                ctx.membrane_stimulus(self.handle, None, [], [])
            self._stim_key = key
        ctx.ode_step(self.handle, float(self.time), float(dt), self.engine.ode_rtol, self.engine.ode_atol,
                     self._set_v)
        self._set_v = False
        self.time = self.time + dt
        return _LazyTable(self)         # (the reference returns self.states; nobody in its loop reads it)
