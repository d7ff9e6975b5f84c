"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the Julia it stands in for (Julia is not in
this image, so the host side above the C ABI is Python; INTEGRATION.md shows the ccall shim a Julia
maintainer would use instead):

  readR1CS(path)                      ParseR1CS.jl:50-124                  -> R1CS
  abstraction(name, cons, sub)        R1CSConstraintSolver.jl:237-395      -> (Specials, R1CS)
  SolveConstraintsSymbolic(...)       R1CSConstraintSolver.jl:583-1646     -> bool   (THE hot path: one
                                                                           ecne_solve call)
  solveWithTrustedFunctions(...)      R1CSConstraintSolver.jl:502-581      -> bool
  printState / printEquation / the "Bad Constraints" listing   :396-456, :1599-1644 (ecne_report_resident)

Everything that computes runs in native code: libecne_host.so (parser, abstraction) and
libecne_b200.so (sm_100a CUDA engine).  There is no Python or CPU fallback for the solver.
"""
import ctypes as C
import time

import numpy as np

from . import _abi
from ._abi import Problem, Result

# Julia exception class -> Python exception raised by the mirror
_STATUS_EXC = {
    _abi.ECNE_E_BADARG: AssertionError,
    _abi.ECNE_E_DIVZERO: ZeroDivisionError,   # DivideError
    _abi.ECNE_E_BOUNDS: IndexError,           # BoundsError
    _abi.ECNE_E_NODSU: NameError,             # UndefVarError(:dsu)
    _abi.ECNE_E_CUDA: RuntimeError,
    _abi.ECNE_E_NCCL: RuntimeError,
    _abi.ECNE_E_UNSUPPORTED: NotImplementedError,
    _abi.ECNE_E_NOCONVERGE: RuntimeError,
    _abi.ECNE_E_INTERNAL: RuntimeError,
    _abi.ECNE_E_KEYERROR: KeyError,
    _abi.ECNE_E_IO: OSError,
    _abi.ECNE_E_ASSERT: AssertionError,
}

_KIND_OF_NAME = {"BigMultModP": _abi.SPECIAL_BIGMULTMODP, "BigLessThan": _abi.SPECIAL_BIGLESSTHAN}


def _raise(status, msg):
    raise _STATUS_EXC.get(status, RuntimeError)(f"[ecne status {status}] {msg}")


def _as_np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


class R1CS:
    """A flattened constraint system (include/ecne_abi.h layout).  Owns its native struct."""

    def __init__(self, native):
        self._native = native  # POINTER(R1CSStruct)
        s = native.contents
        self.n_rows = int(s.n_rows)
        self.n_vars = int(s.n_vars)
        self.nnz = int(s.nnz)
        self.seg_ptr = _as_np(s.seg_ptr, 3 * self.n_rows + 1, np.uint64)
        self.col = _as_np(s.col, self.nnz, np.uint32)
        # (None after readR1CS(..., compact_only=True): the rows exist in the compact form only)
        self.coef = _as_np(s.coef, 4 * self.nnz, np.uint64).reshape(-1, 4) if s.coef else None
        self.known = _as_np(s.known, int(s.n_known), np.uint32)
        self.targets = _as_np(s.targets, int(s.n_targets), np.uint32)
        # the compact form of the same rows, when the reader made it (include/ecne_host.h)
        self.compact = None
        if s.coef_class and s.seg_ptr32:
            n_o = int(s.n_coef_other)
            self.compact = (_as_np(s.coef_class, self.nnz, np.uint8), _as_np(s.coef_other, 4 * max(n_o, 1), np.uint64)[:4 * n_o],
                            _as_np(s.coef_other_term, max(n_o, 1), np.uint32)[:n_o], _as_np(s.seg_ptr32, 3 * self.n_rows + 1, np.uint32))
        self.n_pub_out = int(s.n_pub_out)
        self.n_pub_in = int(s.n_pub_in)
        self.n_prv_in = int(s.n_prv_in)

    def __len__(self):
        return self.n_rows

    def __del__(self):
        try:
            if self._native:
                _abi.host_lib().ecne_r1cs_free(self._native)
                self._native = None
        except Exception:
            pass


class Specials:
    """special_constraints (:585): a list of (name, inputs, outputs)."""

    def __init__(self):
        self._native = _abi.host_lib().ecne_specials_new()
        self.names = []

    def __len__(self):
        return int(self._native.contents.n)

    def as_list(self):
        s = self._native.contents
        n = int(s.n)
        ip = _as_np(s.in_ptr, n + 1, np.uint64)
        op = _as_np(s.out_ptr, n + 1, np.uint64)
        iv = _as_np(s.in_, int(ip[-1]) if n else 0, np.uint32)
        ov = _as_np(s.out, int(op[-1]) if n else 0, np.uint32)
        return [(self.names[i], iv[int(ip[i]):int(ip[i + 1])].tolist(),
                 ov[int(op[i]):int(op[i + 1])].tolist()) for i in range(n)]

    def __del__(self):
        try:
            if self._native:
                _abi.host_lib().ecne_specials_free(self._native)
                self._native = None
        except Exception:
            pass


class SolveResult:
    """Final VariableState arrays + counters of one SolveConstraintsSymbolic call."""

    def __init__(self, n_vars, full_state=False):
        words = (n_vars + 63) // 64
        self.n_vars = n_vars
        self.unique_bits = np.zeros(words, dtype=np.uint64)
        self.known_bits = np.zeros(words, dtype=np.uint64)
        self.lb = self.ub = self.nvalues = self.values = self.abz = None
        if full_state:
            self.lb = np.zeros((n_vars, 4), dtype=np.uint64)
            self.ub = np.zeros((n_vars, 4), dtype=np.uint64)
            self.nvalues = np.zeros(n_vars, dtype=np.uint8)
            self.values = np.zeros((n_vars, 2, 4), dtype=np.uint64)
            self.abz = np.zeros(n_vars, dtype=np.int32)
        self.c = Result()
        self.c.unique_bits = self.unique_bits.ctypes.data_as(_abi.u64p)
        self.c.known_bits = self.known_bits.ctypes.data_as(_abi.u64p)
        if full_state:
            self.c.lb = self.lb.ctypes.data_as(_abi.u64p)
            self.c.ub = self.ub.ctypes.data_as(_abi.u64p)
            self.c.nvalues = self.nvalues.ctypes.data_as(_abi.u8p)
            self.c.values = self.values.ctypes.data_as(_abi.u64p)
            self.c.abz = self.abz.ctypes.data_as(_abi.i32p)

    @property
    def verdict(self):
        return bool(self.c.verdict)

    def unique_bytes(self):
        """Packed bitmap as SURVEY.md Appendix B.1 hashes it: bit (w-1)&7 of byte (w-1)>>3."""
        return self.unique_bits.view(np.uint8)[: (self.n_vars + 7) // 8].tobytes()

    def known_bytes(self):
        return self.known_bits.view(np.uint8)[: (self.n_vars + 7) // 8].tobytes()

    def counters(self):
        c = self.c
        return {k: getattr(c, k) for k in (
            "n_unique_nontrivial", "n_nontrivial", "n_targets_unique", "n_unique", "outer_rounds",
            "inner_rounds", "constraint_evals", "sweep_launches", "ms_h2d", "ms_classify",
            "ms_solve", "ms_d2h", "ms_exchange", "ms_total", "ms_sweep", "rule_evals")}


def compact_coef(coef):
    """(class bytes, other values [n, 4], their term indices) of a [nnz, 4] limb array: ecne_compact_coef of the host
    library (include/ecne_host.h), the compact coefficient form of ecne_problem_t."""
    lib = _abi.host_lib()
    coef = np.ascontiguousarray(np.asarray(coef, dtype=np.uint64).reshape(-1, 4))
    nnz = coef.shape[0]
    cls = np.zeros(max(nnz, 1), dtype=np.uint8)
    n = C.c_uint64(0)
    cp = coef.ctypes.data_as(_abi.u64p) if nnz else None
    st = lib.ecne_compact_coef(cp, nnz, cls.ctypes.data_as(_abi.u8p), None, None, C.byref(n))
    if st != 0:
        _raise(st, lib.ecne_host_last_error().decode())
    other = np.zeros((max(int(n.value), 1), 4), dtype=np.uint64)
    term = np.zeros(max(int(n.value), 1), dtype=np.uint32)
    st = lib.ecne_compact_coef(cp, nnz, cls.ctypes.data_as(_abi.u8p), other.ctypes.data_as(_abi.u64p),
                               term.ctypes.data_as(_abi.u32p), C.byref(n))
    if st != 0:
        _raise(st, lib.ecne_host_last_error().decode())
    return cls[:nnz], other[:int(n.value)], term[:int(n.value)]


class ProblemHandle:
    """Keeps the numpy buffers an ecne_problem_t points into alive."""

    def __init__(self, constraints, specials, known_variables, target_variables, num_variables,
                 secp_solve=False, debug=False, compact=False):
        """compact: hand the coefficients over in the compact form of include/ecne_abi.h (class bytes + the values
        that are not 0, 1, p - 1) with 32-bit offsets — a third of the bytes of the full form cross PCIe."""
        self.keep = []
        p = Problem()
        p.n_rows = constraints.n_rows
        p.n_vars = num_variables
        p.col = self._buf(constraints.col, np.uint32, _abi.u32p)
        ready = getattr(constraints, "compact", None)
        if compact and ready is not None:  # the reader made the compact form already: no pass over the coefficients
            cls, other, term, seg32 = ready
            p.seg_ptr = None
            p.seg_ptr32 = self._buf(seg32, np.uint32, _abi.u32p)
            p.coef = None
            p.coef_class = self._buf(cls, np.uint8, _abi.u8p)
            p.coef_other = self._buf(other, np.uint64, _abi.u64p)
            p.coef_other_term = self._buf(term, np.uint32, _abi.u32p)
            p.n_coef_other = len(term)
        elif compact:
            cls, other, term = compact_coef(constraints.coef)
            p.seg_ptr = None
            p.seg_ptr32 = self._buf(np.asarray(constraints.seg_ptr).astype(np.uint32), np.uint32, _abi.u32p)
            p.coef = None
            p.coef_class = self._buf(cls, np.uint8, _abi.u8p)
            p.coef_other = self._buf(other.reshape(-1), np.uint64, _abi.u64p)
            p.coef_other_term = self._buf(term, np.uint32, _abi.u32p)
            p.n_coef_other = len(term)
        else:
            p.seg_ptr = self._buf(constraints.seg_ptr, np.uint64, _abi.u64p)
            p.coef = self._buf(constraints.coef.reshape(-1), np.uint64, _abi.u64p)
        self.compact = bool(compact)
        kn = np.asarray(known_variables, dtype=np.uint32)
        tg = np.asarray(target_variables, dtype=np.uint32)
        p.known = self._buf(kn, np.uint32, _abi.u32p)
        p.n_known = len(kn)
        p.targets = self._buf(tg, np.uint32, _abi.u32p)
        p.n_targets = len(tg)
        sl = specials.as_list() if isinstance(specials, Specials) else list(specials or [])
        kinds, ip, iv, op, ov = [], [0], [], [0], []
        for name, ins, outs in sl:
            kinds.append(_KIND_OF_NAME.get(name, _abi.SPECIAL_GENERIC))
            iv += list(ins)
            ov += list(outs)
            ip.append(len(iv))
            op.append(len(ov))
        p.n_specials = len(sl)
        p.sp_kind = self._buf(np.asarray(kinds, dtype=np.int32), np.int32, _abi.i32p)
        p.sp_in_ptr = self._buf(np.asarray(ip, dtype=np.uint64), np.uint64, _abi.u64p)
        p.sp_in = self._buf(np.asarray(iv, dtype=np.uint32), np.uint32, _abi.u32p)
        p.sp_out_ptr = self._buf(np.asarray(op, dtype=np.uint64), np.uint64, _abi.u64p)
        p.sp_out = self._buf(np.asarray(ov, dtype=np.uint32), np.uint32, _abi.u32p)
        p.secp_solve = int(bool(secp_solve))
        p.debug = int(bool(debug))
        self.c = p
        self.n_vars = int(num_variables)
        self.n_rows = int(constraints.n_rows)
        self.nnz = int(constraints.nnz)

    def _buf(self, arr, dtype, ptype):
        a = np.ascontiguousarray(arr, dtype=dtype)
        if a.size == 0:
            a = np.zeros(1, dtype=dtype)
        self.keep.append(a)
        return a.ctypes.data_as(ptype)


# ---------------------------------------------------------------------------------------------
# the mirrored reference functions
# ---------------------------------------------------------------------------------------------
class BadConstraints:
    """The rows and wires of the reference's "Bad Constraints" listing (:1599-1635), found and compacted
    on the device by ecne_report_resident: `rows` are the 0-based ids of the constraints that mention a
    wire that is not unique, `wire` the 1-based ids (ascending, wire 1 left out) of the wires of those
    rows, with their final VariableState."""

    def __init__(self, handle, n_rows):
        lib = _engine()
        self.row_bits = np.zeros((n_rows + 63) // 64, dtype=np.uint64)
        rep = _abi.Report()
        rep.bad_row_bits = self.row_bits.ctypes.data_as(_abi.u64p)
        st = lib.ecne_report_resident(handle, C.byref(rep))   # pass 1: bitmap + counts
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())
        n = int(rep.n_wires)
        self.n_bad_rows = int(rep.n_bad_rows)
        self.wire = np.zeros(n, dtype=np.uint32)
        self.flags = np.zeros(n, dtype=np.uint8)
        self.lb = np.zeros((n, 4), dtype=np.uint64)
        self.ub = np.zeros((n, 4), dtype=np.uint64)
        self.nvalues = np.zeros(n, dtype=np.uint8)
        self.values = np.zeros((n, 2, 4), dtype=np.uint64)
        if n:
            rep.cap_wires = n
            rep.wire = self.wire.ctypes.data_as(_abi.u32p)
            rep.flags = self.flags.ctypes.data_as(_abi.u8p)
            rep.lb = self.lb.ctypes.data_as(_abi.u64p)
            rep.ub = self.ub.ctypes.data_as(_abi.u64p)
            rep.nvalues = self.nvalues.ctypes.data_as(_abi.u8p)
            rep.values = self.values.ctypes.data_as(_abi.u64p)
            st = lib.ecne_report_resident(handle, C.byref(rep))   # pass 2: the compacted state
            if st != 0:
                _raise(st, lib.ecne_last_error().decode())
        bits = np.unpackbits(self.row_bits.view(np.uint8), bitorder="little")[:n_rows]
        self.rows = np.flatnonzero(bits)

    def state(self, w):
        """(unique, lb, ub, values) of a listed wire."""
        i = int(np.searchsorted(self.wire, w))
        if i >= len(self.wire) or int(self.wire[i]) != w:
            raise KeyError(w)
        return (bool(self.flags[i] & 1), _int(self.lb[i]), _int(self.ub[i]),
                [_int(self.values[i, k]) for k in range(int(self.nvalues[i]))])


_P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
# fix_number's threshold (:421-429) is NOT p - something round: the constant is restated digit for digit
_FIX_NUMBER_ABOVE = 21888242871839275222246405745257275088548363400416034343698204186575808495517


def _int(limbs):
    return int.from_bytes(np.ascontiguousarray(limbs, dtype=np.uint64).tobytes(), "little")


def fix_number(x):
    """:421-429."""
    return x - _P if x > _FIX_NUMBER_ABOVE else x


def format_state(unique, lb, ub, values):
    """printState (:396-419) as a string (Julia prints Bool as true/false, a BigInt vector as BigInt[..])."""
    out = ["Uniquely Determined: " + ("true" if unique else "false")]
    if lb == 0 and ub == _P - 1:
        out.append("Bounds: None")
    else:
        out.append(f"Bounds: [{lb}, {ub}]")
    if values:
        out.append("All possible values: BigInt[" + ", ".join(str(v) for v in sorted(values)) + "]")
    out.append("")
    return "\n".join(out) + "\n"


def read_sym(input_sym):
    """The CSV read of :1603-1607: fourth column of every line = the signal name of wire (line + 1)."""
    with open(input_sym) as f:
        return [ln.rstrip("\n").split(",", 3)[3] for ln in f if ln.strip()]


def format_equation(constraints, row, index_to_signal):
    """printEquation (:431-456).  Terms are printed in stored order (the reference iterates a Julia Set:
    its order is a hash order no caller can rely on)."""
    def lin(k):
        b, e = int(constraints.seg_ptr[3 * row + k]), int(constraints.seg_ptr[3 * row + k + 1])
        terms = []
        for j in range(b, e):
            c = _int(constraints.coef[j])
            if c == 0:
                continue
            key = int(constraints.col[j]) - 1
            terms.append(f"{fix_number(c)} * {index_to_signal[key - 1] if key > 0 else 1}")
        return "(" + " + ".join(terms) + ")" if terms else "0"
    return lin(0) + " * " + lin(1) + " = " + lin(2)


def row_variables(constraints, row):
    """getVariables (:36-56) of one row, ascending (the reference's Set has no defined order)."""
    b, e = int(constraints.seg_ptr[3 * row]), int(constraints.seg_ptr[3 * row + 3])
    nz = constraints.coef[b:e].any(axis=1)
    return np.unique(constraints.col[b:e][nz])


def nontrivial_variables(constraints):
    """all_nontrivial_vars (:600-618): every wire with a non-zero coefficient somewhere, ascending."""
    return np.unique(constraints.col[constraints.coef.any(axis=1)])


def format_all_states(constraints, result):
    """The debug dump of :1573-1577: printState of every non-trivial variable (wire 1 included, as in the reference)."""
    ubits = np.unpackbits(result.unique_bits.view(np.uint8), bitorder="little")
    out = []
    for i in nontrivial_variables(constraints):
        i = int(i)
        vals = [_int(result.values[i - 1, k]) for k in range(int(result.nvalues[i - 1]))]
        out.append(format_state(bool(ubits[i - 1]), _int(result.lb[i - 1]), _int(result.ub[i - 1]), vals))
    return "".join(out)


def format_listing(constraints, bad, result, index_to_signal):
    """Everything :1609-1643 prints after the "------ Bad Constraints ------" header."""
    out = []
    for row in bad.rows:
        row = int(row)
        out.append(f"constraint #{row + 1}\n")
        out.append(format_equation(constraints, row, index_to_signal) + "\n")
        for j in row_variables(constraints, row):
            j = int(j)
            if j == 1:
                continue
            out.append(index_to_signal[j - 2] + "\n")
            out.append(format_state(*bad.state(j)))
    out.append("------ All Variables ------\n\n")
    ubits = np.unpackbits(result.unique_bits.view(np.uint8), bitorder="little")
    for i in nontrivial_variables(constraints):
        i = int(i)
        if i == 1:
            continue
        out.append(index_to_signal[i - 2] + "\n")
        vals = [_int(result.values[i - 1, k]) for k in range(int(result.nvalues[i - 1]))]
        out.append(format_state(bool(ubits[i - 1]), _int(result.lb[i - 1]), _int(result.ub[i - 1]), vals))
    return "".join(out)


def readR1CS(filename, compact_only=False):
    """ParseR1CS.readR1CS (ParseR1CS.jl:50).  The reference returns (equations, knowns, outs,
    num_wires+1); here those four live on the returned R1CS (.known, .targets, .n_vars).
    compact_only: leave out the full 32-byte coefficient array (.coef is None) — for rows that go to the device in the
    compact form (solve_with_device_abstraction); it is 77 % of what the reader writes."""
    lib = _abi.host_lib()
    out = C.POINTER(_abi.R1CSStruct)()
    st = lib.ecne_read_r1cs_opts(str(filename).encode(), 1 if compact_only else 0, C.byref(out))
    if st != 0:
        _raise(st, lib.ecne_host_last_error().decode())
    return R1CS(out)


def abstraction(function_name, constraints, sub_equation, specials=None):
    """abstraction() (R1CSConstraintSolver.jl:237).  known_inputs / known_outputs of the trusted
    circuit travel inside `sub_equation`.  Returns (specials, reduced)."""
    lib = _abi.host_lib()
    sp = specials if specials is not None else Specials()
    red = C.POINTER(_abi.R1CSStruct)()
    n = C.c_uint64(0)
    kind = _KIND_OF_NAME.get(function_name, _abi.SPECIAL_GENERIC)
    st = lib.ecne_abstraction(kind, constraints._native, sub_equation._native, C.byref(red),
                              sp._native, C.byref(n))
    if st != 0:
        _raise(st, lib.ecne_host_last_error().decode())
    sp.names += [function_name] * int(n.value)
    return sp, R1CS(red)


class PreparedTrusted:
    """A trusted circuit with the part of abstraction() that depends on it alone done ahead (ecne_abstract_prepare:
    host work that touches neither the GPU nor any library state — run it on another thread while the main circuit is
    still being read)."""

    def __init__(self, function_name, sub):
        self.name = function_name
        self.sub = sub
        self.ph = ProblemHandle(sub, None, sub.known, sub.targets, sub.n_vars)
        self.handle = C.c_void_p()
        st = _engine().ecne_abstract_prepare(C.byref(self.ph.c), C.byref(self.handle))
        if st != 0:
            _raise(st, "ecne_abstract_prepare failed")

    def __len__(self):
        return len(self.sub)

    def free(self):
        if self.handle:
            _engine().ecne_abstract_prepared_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceAbstraction:
    """abstraction() (R1CSConstraintSolver.jl:237-395) on the GPU (include/ecne_abi.h "abstraction() on the device"):
    the unreduced system is uploaded once, every `apply` replaces the windows isomorphic to one trusted circuit by
    special constraints with kernels, `upload` classifies the reduced system where it lies and returns the resident
    handle ecne_solve_resident takes.  `export` brings the reduced system and the specials back (tests)."""

    def __init__(self, main):
        lib = _engine()
        self.main = main
        # (the compact form the reader made beside the full one: a quarter of the bytes cross PCIe)
        self.ph = ProblemHandle(main, None, main.known, main.targets, main.n_vars,
                                compact=getattr(main, "compact", None) is not None)
        self.handle = C.c_void_p()
        self.names = []
        st = lib.ecne_abstract_begin(C.byref(self.ph.c), C.byref(self.handle))
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())

    def apply(self, function_name, sub):
        lib = _engine()
        n = C.c_uint64(0)
        if isinstance(sub, PreparedTrusted):
            st = lib.ecne_abstract_apply_prepared(self.handle, _KIND_OF_NAME.get(function_name, _abi.SPECIAL_GENERIC),
                                                  C.byref(sub.ph.c), sub.handle, C.byref(n))
        else:
            sp = ProblemHandle(sub, None, sub.known, sub.targets, sub.n_vars)
            st = lib.ecne_abstract_apply(self.handle, _KIND_OF_NAME.get(function_name, _abi.SPECIAL_GENERIC), C.byref(sp.c), C.byref(n))
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())
        self.names += [function_name] * int(n.value)
        return int(n.value)

    def sizes(self):
        z = (C.c_uint64 * 5)()
        st = _engine().ecne_abstract_sizes(self.handle, z)
        if st != 0:
            _raise(st, _engine().ecne_last_error().decode())
        return [int(x) for x in z]

    def export(self):
        """(seg_ptr, col, coef[nnz, 4]) of the reduced system and the specials as [(name, inputs, outputs)]."""
        rows, nnz, ns, n_in, n_out = self.sizes()
        seg = np.zeros(3 * rows + 1, dtype=np.uint64)
        col = np.zeros(max(nnz, 1), dtype=np.uint32)
        coef = np.zeros((max(nnz, 1), 4), dtype=np.uint64)
        kind = np.zeros(max(ns, 1), dtype=np.int32)
        ip, op = np.zeros(ns + 1, dtype=np.uint64), np.zeros(ns + 1, dtype=np.uint64)
        iv, ov = np.zeros(max(n_in, 1), dtype=np.uint32), np.zeros(max(n_out, 1), dtype=np.uint32)
        st = _engine().ecne_abstract_export(self.handle, seg.ctypes.data_as(_abi.u64p), col.ctypes.data_as(_abi.u32p),
                                            coef.ctypes.data_as(_abi.u64p), kind.ctypes.data_as(_abi.i32p),
                                            ip.ctypes.data_as(_abi.u64p), iv.ctypes.data_as(_abi.u32p),
                                            op.ctypes.data_as(_abi.u64p), ov.ctypes.data_as(_abi.u32p))
        if st != 0:
            _raise(st, _engine().ecne_last_error().decode())
        sp = [(self.names[i], iv[int(ip[i]):int(ip[i + 1])].tolist(), ov[int(op[i]):int(op[i + 1])].tolist()) for i in range(ns)]
        return (seg, col[:nnz], coef[:nnz]), sp

    def upload(self, secp_solve=False):
        h = C.c_void_p()
        st = _engine().ecne_abstract_upload(self.handle, int(bool(secp_solve)), C.byref(h))
        if st != 0:
            _raise(st, _engine().ecne_last_error().decode())
        return h

    def free(self):
        if self.handle:
            _engine().ecne_abstract_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def read_and_prepare(input_r1cs, trusted_r1cs=(), trusted_r1cs_names=()):
    """The host half of solve_with_device_abstraction: the main circuit is parsed on a worker thread (the native parser
    is multi-threaded itself and releases the GIL) while this thread reads the — small — trusted circuits and prepares
    them (ecne_abstract_prepare).  Returns (main, [(name, PreparedTrusted)] longest first, :527)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(1) as ex:
        fut = ex.submit(readR1CS, input_r1cs, True)  # (the device takes the compact form: no full coefficient array)
        function_list = [(trusted_r1cs_names[i], PreparedTrusted(trusted_r1cs_names[i], readR1CS(trusted_r1cs[i])))
                         for i in range(len(trusted_r1cs))]
        function_list.sort(key=lambda x: -len(x[1]))  # stable, longest first (:527)
        main = fut.result()
    return main, function_list


def solve_with_device_abstraction(input_r1cs, trusted_r1cs=(), trusted_r1cs_names=(), secp_solve=False, full_state=False):
    """solveWithTrustedFunctions (:502-581) with abstraction() on the GPU: parse (host), upload the unreduced system,
    abstract the trusted circuits longest first (:527-544) on the device, classify in place, solve.
    Returns (verdict, SolveResult, DeviceAbstraction sizes)."""
    global last_result
    main, function_list = read_and_prepare(input_r1cs, trusted_r1cs, trusted_r1cs_names)
    da = DeviceAbstraction(main)
    try:
        for name, sub in function_list:
            da.apply(name, sub)
        sizes = da.sizes()
        h = da.upload(secp_solve)
    finally:
        da.free()
    lib = _engine()
    res = SolveResult(main.n_vars, full_state=full_state)
    try:
        st = lib.ecne_solve_resident(h, C.byref(res.c))
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())
    finally:
        lib.ecne_free_resident(h)
    last_result = res
    return bool(res.c.verdict), res, sizes


_initialised = False


def init_multi(n_gpus):
    """One process, the GPUs 0 .. n_gpus-1 of the box (ecne_init_multi): every later solve shards its rows over them."""
    global _initialised
    lib = _abi.engine_lib()
    st = lib.ecne_init_multi(int(n_gpus))
    if st != 0:
        _raise(st, lib.ecne_last_error().decode())
    _initialised = True


def _engine():
    global _initialised
    lib = _abi.engine_lib()
    if not _initialised:
        import os
        if int(os.environ.get("ECNE_GPUS", "1")) > 1:   # one process, several GPUs
            init_multi(int(os.environ["ECNE_GPUS"]))
            return lib
        dev = int(os.environ.get("LOCAL_RANK", os.environ.get("ECNE_DEVICE", "0")))
        st = lib.ecne_init(dev)
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())
        _initialised = True
    return lib


last_result = None
last_bad_constraints = None


def SolveConstraintsSymbolic(constraints, special_constraints, known_variables, debug=False,
                             target_variables=(), num_variables=-1, input_sym="default.sym",
                             secp_solve=False, full_state=False):
    """R1CSConstraintSolver.jl:583-1646, executed by the CUDA engine through ecne_solve.

    Returns function_good (:1645).  With a non-empty `input_sym` the listing of :1599-1644 is printed as
    the reference prints it (a missing file raises, as CSV.File does); the rows and wires of its first part
    come compacted from the device (ecne_report_resident) and are kept in `api.last_bad_constraints`.  The
    per-wire state of the call is kept in `ecneproject_b200.api.last_result` (a SolveResult)."""
    global last_result, last_bad_constraints
    listing = input_sym != ""
    index_to_signal = read_sym(input_sym) if listing else None   # a bad path fails before any GPU work
    lib = _engine()
    ph = ProblemHandle(constraints, special_constraints, known_variables, target_variables,
                       num_variables, secp_solve, debug)
    res = SolveResult(int(num_variables), full_state=full_state or listing or debug)
    bad = None
    if not listing:
        st = lib.ecne_solve(C.byref(ph.c), C.byref(res.c))
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())
    else:
        h = C.c_void_p()
        st = lib.ecne_upload(C.byref(ph.c), C.byref(h))
        if st != 0:
            _raise(st, lib.ecne_last_error().decode())
        try:
            st = lib.ecne_solve_resident(h, C.byref(res.c))
            if st != 0:
                _raise(st, lib.ecne_last_error().decode())
            bad = BadConstraints(h, len(constraints))
        finally:
            lib.ecne_free_resident(h)
    last_result = res
    last_bad_constraints = bad
    print(f"Solved for {res.c.n_unique_nontrivial} variables out of {res.c.n_nontrivial} total variables")
    if debug:  # :1573-1577: the state of every variable that occurs in a constraint (ascending; the Julia walks a Set)
        print(format_all_states(constraints, res), end="")
    print(f"Solved for {res.c.n_targets_unique} target variables out of {len(target_variables)} total target variables")
    print("------ Bad Constraints ------")
    print()
    if listing:
        print(format_listing(constraints, bad, res, index_to_signal), end="")
    return res.verdict


def prepare(input_r1cs, trusted_r1cs=(), trusted_r1cs_names=(), printRes=False):
    """Lines :513-544 of solveWithTrustedFunctions: parse, sort trusted longest-first, abstract.
    Returns (reduced R1CS, Specials, main R1CS)."""
    assert len(trusted_r1cs) == len(trusted_r1cs_names)
    main = readR1CS(input_r1cs)
    function_list = [(trusted_r1cs_names[i], readR1CS(trusted_r1cs[i])) for i in range(len(trusted_r1cs))]
    function_list.sort(key=lambda x: -len(x[1]))  # stable, longest first (:527)
    specials = Specials()
    reduced = main
    for name, sub in function_list:
        if printRes:
            print("called abstraction")
        specials, reduced = abstraction(name, reduced, sub, specials)
    return reduced, specials, main


def solveWithTrustedFunctions(input_r1cs, input_r1cs_name, trusted_r1cs=(), trusted_r1cs_names=(),
                              debug=False, printRes=True, abstractionOnly=False, input_sym="",
                              secp_solve=False):
    """solveWithTrustedFunctions (R1CSConstraintSolver.jl:502-581)."""
    a = time.time()
    reduced, specials, main = prepare(input_r1cs, list(trusted_r1cs), list(trusted_r1cs_names), printRes)
    if abstractionOnly:
        print(specials.as_list())
        return True
    print("time to prep inputs", round(time.time() - a, 3), "seconds")
    result = SolveConstraintsSymbolic(reduced, specials, main.known, debug, main.targets,
                                      main.n_vars, input_sym, secp_solve)
    if result:
        if len(trusted_r1cs) != 0:
            if printRes:
                print("R1CS function " + input_r1cs_name +
                      " has sound constraints assuming trusted functions " + ", ".join(trusted_r1cs_names))
        elif printRes:
            print("R1CS function " + input_r1cs_name + " has sound constraints (No trusted functions needed!)")
        return True
    if printRes:
        print("R1CS function " + input_r1cs_name + " has potentially unsound constraints")
    return False
