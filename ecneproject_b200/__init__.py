"""ecneproject_b200 — B200-native engine for Ecne's R1CS soundness-propagation hot path.

Only what the path needs: csrc/ (sm_100a CUDA kernels + the C ABI of include/ecne_abi.h and the
host-side parser/abstraction of include/ecne_host.h) and the Python mirror of the reference's
operator interface (api.py).
"""
from .api import (R1CS, Specials, SolveResult, ProblemHandle, readR1CS, abstraction,
                  SolveConstraintsSymbolic, solveWithTrustedFunctions, prepare)

__all__ = ["R1CS", "Specials", "SolveResult", "ProblemHandle", "readR1CS", "abstraction",
           "SolveConstraintsSymbolic", "solveWithTrustedFunctions", "prepare"]
