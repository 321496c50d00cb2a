"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/openmg_oracle.py header).

CPU restatement of the reference's V-cycle path used as the checker by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Nothing in
openmg_b200/ imports it.
"""
from .openmg_oracle import *  # noqa: F401,F403
from . import openmg_oracle  # noqa: F401
