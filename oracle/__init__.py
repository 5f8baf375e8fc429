"""CPU oracle for the KNP-EMI-DG hot path: TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package; the product (knp-emi-dg_b200/) never
does.  Pinned by executing the reference itself (oracle/refexec, tests/golden/ref_*.npz); see the
headers of the individual modules.
"""
