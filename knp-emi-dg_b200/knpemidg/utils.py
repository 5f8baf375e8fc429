"""DG/trace helpers with the reference's names (src/knpemidg/utils.py:44-124).  On the
B200 path they do not assemble anything: they build handles that the library evaluates on
the device (knp_facet_trace, or a link executed inside the membrane ODE kernel)."""
from .frontend import CellField, FacetMean, InterfaceNormal, Trace


def interface_normal(subdomains=None, mesh=None):
    """n_g: on a membrane facet it points from the lower to the higher cell tag
    (utils.py:80; README.md:67-72).  The library encodes this orientation in its membrane
    table (ICS side = higher tag); the returned object only carries the convention."""
    return InterfaceNormal()


def plus(phi, normal=None):
    """restriction of phi to the cell the normal originates from = ECS / lower tag (utils.py:87-92)"""
    if not isinstance(phi, CellField):
        raise TypeError("plus() expects a cell field handle (e.g. solver.c_prev_k.split()[0])")
    return Trace(phi, 0)


def minus(phi, normal=None):
    """restriction of phi to the cell the normal ends in = ICS / higher tag (utils.py:94-98)"""
    if not isinstance(phi, CellField):
        raise TypeError("minus() expects a cell field handle (e.g. solver.ion_list[-1]['c'])")
    return Trace(phi, 1)


def pcws_constant_project(f, V=None, fV=None):
    """facet mean of a one-sided trace (utils.py:100-124)"""
    if not isinstance(f, Trace):
        raise TypeError("pcws_constant_project() supports plus(field, n_g) / minus(field, n_g) arguments")
    return FacetMean(f)


def subdomain_marking_foo(subdomains, V=None):
    """cell tags as an array (the reference interpolates them into DG0, utils.py:44-59)"""
    return subdomains.array().copy()
