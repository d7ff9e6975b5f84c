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
        l.ecne_oracle_odd_permutation_sum.argtypes = [C.c_uint32, _abi.u64p, _abi.u64p, _abi.u64p]
        l.ecne_oracle_odd_permutation_sum.restype = C.c_int
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


def odd_permutation_sum(matrix, enumerate_too=True):
    """(by enumeration of the k! permutations or None, by the subset recurrence) of a k x k list of ints mod p."""
    from helpers import to_limbs, from_limbs
    k = len(matrix)
    m = np.ascontiguousarray(to_limbs([v for row in matrix for v in row]))
    oe, od = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    st = lib().ecne_oracle_odd_permutation_sum(k, m.ctypes.data_as(_abi.u64p),
                                               oe.ctypes.data_as(_abi.u64p) if enumerate_too else None,
                                               od.ctypes.data_as(_abi.u64p))
    assert st == 0, st
    return (from_limbs(oe)[0] if enumerate_too else None), from_limbs(od)[0]
