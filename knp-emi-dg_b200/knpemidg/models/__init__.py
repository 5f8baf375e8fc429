"""Bundled membrane (ODE) models; each is compiled into the CUDA library by
`__graft_entry__.build()` via knpemidg.odegen."""
BUNDLED = ("mm_hh", "mm_hh_no_stim", "mm_leak", "mm_hh_emix", "mm_glial_emix",
           "mm_hh_astro", "mm_glial_astro", "mm_calibration")
