"""CPU oracle for the PhyloDist hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See felsenstein_oracle.c for what it restates and how it is pinned.
"""
from .oracle import (build, codes_to_dense, felsenstein, num_threads, transition,  # noqa: F401
                     compound_dirichlet_logpdf, compound_dirichlet_gradlogpdf, exponential_bl_gradlogpdf)
from .extended import expm_extended, felsenstein_extended  # noqa: F401
