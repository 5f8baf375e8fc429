"""Import shim for the `numbalsoda` package (absent from this image).

The reference's ODE modules do `from numbalsoda import lsoda_sig` at import
time (e.g. examples/idealized-geometries/mm_hh.py:112) and decorate their
right-hand side with `@cfunc(lsoda_sig)`.  This shim supplies that signature
so such modules import unchanged.  The GPU path never calls the compiled
cfunc: it translates the function's Python source to CUDA
(knpemidg/odegen.py).  `lsoda` is deliberately not a CPU integrator here:
the product path has no CPU fallback (the CPU LSODA lives in oracle/ode.py).
"""
try:
    from numba import types as _t

    lsoda_sig = _t.void(_t.double, _t.CPointer(_t.double), _t.CPointer(_t.double),
                        _t.CPointer(_t.double))
except Exception:  # numba missing: the decorator is replaced in knpemidg.odegen
    lsoda_sig = None


def lsoda(*args, **kwargs):
    raise NotImplementedError(
        "numbalsoda shim: the B200 build integrates membrane ODEs on the GPU "
        "(MembraneModel.step_lsoda); there is no CPU LSODA in the product path")
