"""CPU oracle for the equity hot path -- TEST INFRASTRUCTURE, never imported by neuron_poker_b200/.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this package.
See oracle/npk_oracle.c for what is restated and how it is pinned.
"""
from .oracle import *  # noqa: F401,F403
