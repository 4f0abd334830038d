"""CPU oracle of the D-equation hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (fluidstructureinteraction_b200) never does.  Bit-level parity with foam-extend is unpinned (no
tree, no logs); the restated algorithm is pinned by three known answers (Kirsch on config 1, Lame, thermal expansion): see
oracle/ldu_oracle.c and oracle/ASSUMPTIONS.md.
"""
from .binding import *  # noqa: F401,F403
