"""dcora_b200: B200-native DCORA Riemannian-staircase hot path.

The compute lives in libdcora_b200.so (hand-written sm_100a CUDA behind the
C-ABI of include/dcora_b200.h).  This package is the thin host-side mirror of
the reference's QuadraticProblem / QuadraticOptimizer interface plus the data
plumbing the tests and the bench need; it never computes on the CPU.
"""
from .api import (AgentData, QuadraticOptimizer, QuadraticProblem, ROptParameters,
                  ROPTResult, ORDER_RA, ORDER_SE, RETRACT_POLAR, RETRACT_QF)

__all__ = ["AgentData", "QuadraticOptimizer", "QuadraticProblem", "ROptParameters",
           "ROPTResult", "ORDER_RA", "ORDER_SE", "RETRACT_POLAR", "RETRACT_QF"]
