"""TEST INFRASTRUCTURE -- CPU oracle for the batched-dynamics hot path (see rbd_oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from .oracle import Oracle, build_oracle  # noqa: F401
