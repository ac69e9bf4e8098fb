"""CPU oracle for the Swiftest hot path -- TEST INFRASTRUCTURE ONLY (see swiftest_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product package `swiftest_b200` never does.
"""
from .oracle import Oracle, load  # noqa: F401
