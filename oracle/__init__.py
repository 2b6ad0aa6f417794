"""CPU oracle for the Molly.jl pairwise non-bonded + VelocityVerlet path.

TEST INFRASTRUCTURE ONLY. Nothing in the product package imports this; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm do.
"""
from .oracle import *  # noqa: F401,F403
