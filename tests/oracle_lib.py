"""Test-side loader of the CPU oracle (oracle/_build/liboracle.so).  tests/ only."""
import ctypes as C
import os

import numpy as np

from ecneproject_b200 import _abi
from ecneproject_b200.api import ProblemHandle, SolveResult

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(path):
            from ecneproject_b200 import build
            build.build_oracle()
        l = C.CDLL(path)
        l.ecne_oracle_solve.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Result)]
        l.ecne_oracle_solve.restype = C.c_int
        l.ecne_oracle_last_error.restype = C.c_char_p
        l.ecne_oracle_counters.argtypes = [_abi.u64p, C.c_int]
        l.ecne_oracle_set_max_pops.argtypes = [C.c_uint64]
        l.ecne_oracle_set_max_outer.argtypes = [C.c_uint64]
        l.ecne_oracle_fr.argtypes = [C.c_int, C.c_uint64, _abi.u64p, _abi.u64p, _abi.u64p]
        l.ecne_oracle_fr.restype = C.c_int
        _lib = l
    return _lib


class OracleError(Exception):
    def __init__(self, status, msg):
        super().__init__(f"[oracle status {status}] {msg}")
        self.status = status


def solve(constraints, specials, known, targets, n_vars, secp_solve=False, full_state=True):
    """Run the oracle on the same flattened problem the engine takes.  Returns SolveResult."""
    ph = ProblemHandle(constraints, specials, known, targets, n_vars, secp_solve)
    res = SolveResult(int(n_vars), full_state=full_state)
    st = lib().ecne_oracle_solve(C.byref(ph.c), C.byref(res.c))
    if st != 0:
        raise OracleError(st, lib().ecne_oracle_last_error().decode())
    cnt = np.zeros(32, dtype=np.uint64)
    lib().ecne_oracle_counters(cnt.ctypes.data_as(_abi.u64p), 32)
    res.oracle_counters = cnt
    return res
