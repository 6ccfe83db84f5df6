"""CPU oracle for the BPMF Gibbs sweep — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. See bpmf_oracle.hpp for what it restates and for the parity status
("parity unpinned against reference outputs").
"""
from .oracle import Oracle, lib, build, philox4x32_10, words, randn, gamma_then_randn, hyper  # noqa: F401
