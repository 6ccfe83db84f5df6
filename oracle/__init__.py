"""CPU oracle for the BPMF Gibbs sweep — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. See bpmf_oracle.hpp for what it restates and for the parity status
(pinned against the reference's own hot-path sources compiled with stand-in Eigen / Random123 headers, see
bpmf_oracle.hpp; unpinned only in the operation order inside Eigen's kernels).
"""
from .oracle import Oracle, lib, build, philox4x32_10, words, randn, gamma_then_randn, hyper  # noqa: F401
