"""CPU oracle for the go-vectorsearch hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package (see oracle/oracle.c).  PARITY UNPINNED: the reference has no
tests or golden vectors and cannot be built here (no Go toolchain).
"""
from .binding import *  # noqa: F401,F403
