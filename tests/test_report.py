"""The report path (R1CSConstraintSolver.jl:1599-1644).

CPU part: the formatters (printState :396-419, fix_number :421-429, printEquation :431-456, the .sym read
:1603-1607) against strings written out by hand from the Julia source.
GPU part: ecne_report_resident against what the ORACLE's final state implies for the same problem — the rows
that mention a non-unique wire, their wires, and those wires' state — and the whole printed listing."""
import contextlib
import ctypes as C
import io
import types

import numpy as np
import pytest

from configs import CONFIGS
from ecneproject_b200 import _abi, api, fixtures
from helpers import MiniR1CS, P, to_limbs
import oracle_lib


# ---------------------------------------------------------------- CPU: formatters
def test_format_state_matches_printstate():
    assert api.format_state(False, 0, P - 1, []) == "Uniquely Determined: false\nBounds: None\n\n"
    assert api.format_state(True, 0, 1, [1, 0]) == \
        "Uniquely Determined: true\nBounds: [0, 1]\nAll possible values: BigInt[0, 1]\n\n"
    # bounds are printed as soon as ONE end moved (:398)
    assert api.format_state(False, 0, 255, []) == "Uniquely Determined: false\nBounds: [0, 255]\n\n"
    assert api.format_state(True, 7, 7, [7]).splitlines()[2] == "All possible values: BigInt[7]"


def test_fix_number_uses_the_reference_threshold():
    # :422 compares with p - 10^33 - 100 (sic), not with p / 2
    thr = 21888242871839275222246405745257275088548363400416034343698204186575808495517
    assert api.fix_number(P - 1) == -1
    assert api.fix_number(thr) == thr
    assert api.fix_number(thr + 1) == thr + 1 - P
    assert api.fix_number(5) == 5


def test_format_equation_and_sym(tmp_path):
    sym = tmp_path / "t.sym"
    sym.write_text("1,1,0,main.out\n2,2,0,main.in[0]\n3,-1,0,main.a,b\n")
    names = api.read_sym(str(sym))
    assert names == ["main.out", "main.in[0]", "main.a,b"]  # only the first three commas split
    # (2*w2 - 1) * (w3) = (), with an explicit stored zero on wire 4 of A that getVariables ignores
    m = MiniR1CS([({2: 2, 1: -1, 4: 0}, {3: 1}, {})], n_vars=4, known=[1], targets=[2])
    assert api.format_equation(m, 0, names) == "(2 * main.out + -1 * 1) * (1 * main.in[0]) = 0"
    assert list(api.row_variables(m, 0)) == [1, 2, 3]


def test_missing_sym_raises_before_any_gpu_work():
    m = MiniR1CS([({2: 1}, {1: 1}, {2: 1})], n_vars=2, known=[1], targets=[2])
    with pytest.raises(OSError):  # CSV.File on a missing path throws in the reference (:1603)
        api.SolveConstraintsSymbolic(m, [], m.known, False, m.targets, m.n_vars, "/nonexistent/default.sym")


def test_report_struct_matches_header():
    import os, re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ecne_abi.h")).read()
    body = re.search(r"typedef struct ecne_report \{(.*?)\} ecne_report_t;", hdr, re.S).group(1)
    names = re.findall(r"\b(\w+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == [f[0] for f in _abi.Report._fields_]


# ---------------------------------------------------------------- GPU: parity with the oracle
def _prepare(name):
    cfg = CONFIGS[name]
    return api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])],
                       cfg.get("trusted_names", [])), cfg.get("secp_solve", False)


def _expected(reduced, unique_bits, n_vars):
    """Rows / wires of the listing from a final `unique` bitmap: numpy restatement of :1609-1633."""
    uniq = np.unpackbits(unique_bits.view(np.uint8), bitorder="little")[:n_vars].astype(bool)
    nz = reduced.coef.any(axis=1)
    seg = reduced.seg_ptr.astype(np.int64)
    row_of_term = np.repeat(np.arange(reduced.n_rows), np.diff(seg[::3]))
    bad_term = nz & ~uniq[reduced.col.astype(np.int64) - 1]
    bad_row = np.zeros(reduced.n_rows, dtype=bool)
    bad_row[row_of_term[bad_term]] = True
    in_bad = nz & bad_row[row_of_term]
    wires = np.unique(reduced.col[in_bad])
    return np.flatnonzero(bad_row), wires[wires != 1], uniq


def _solve_with_report(reduced, specials, main, secp):
    lib = api._engine()
    ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, secp)
    h = C.c_void_p()
    assert lib.ecne_upload(C.byref(ph.c), C.byref(h)) == 0, lib.ecne_last_error()
    try:
        res = api.SolveResult(main.n_vars, full_state=True)
        assert lib.ecne_solve_resident(h, C.byref(res.c)) == 0, lib.ecne_last_error()
        bad = api.BadConstraints(h, reduced.n_rows)
    finally:
        lib.ecne_free_resident(h)
    return res, bad


REPORT_CONFIGS = ["target/division", "root/bad_bd_check", "circomlib/Decoder@multiplexer", "circomlib/IsZero@comparators",
                  "circomlib/Point2Bits@pointbits", "circomlib/BabyPbk@babyjub", "circomlib/EdDSAPoseidonVerifier@eddsaposeidon",
                  "root/biglessthan", "root/secp256k1", "tornado/withdraw", "secp256k1+bmmp+blt", "root/poseidon",
                  "tornado/withdraw+pedersen", "benchmarks/bigmod_86_3",
                  # the two circuits whose listing shows schedule-dependent wires: compared against the
                  # oracle everywhere except on the pinned wires, which must hold exactly their pinned state
                  "circomlib/Bits2Point_Strict@pointbits", "circomlib/EdDSAVerifier@eddsa"]
import json as _json
import os as _os
SCHEDULE_DIFFS = _json.load(open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden",
                                               "schedule_dependent_diffs.json")))


def _toint(a):
    return sum(int(a[i]) << (64 * i) for i in range(4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", REPORT_CONFIGS)
def test_bad_constraints_match_oracle(name):
    (reduced, specials, main), secp = _prepare(name)
    res, bad = _solve_with_report(reduced, specials, main, secp)
    o = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, secp)
    rows, wires, uniq = _expected(reduced, o.unique_bits, main.n_vars)
    assert np.array_equal(bad.rows, rows)
    assert bad.n_bad_rows == len(rows)
    assert np.array_equal(bad.wire, wires)
    w0 = wires.astype(np.int64) - 1
    assert np.array_equal((bad.flags & 1).astype(bool), uniq[w0])
    # the compacted state is the engine's own full export restricted to the listed wires ...
    kbits = np.unpackbits(res.known_bits.view(np.uint8), bitorder="little")[:main.n_vars]
    assert np.array_equal((bad.flags >> 1) & 1, kbits[w0])
    assert np.array_equal(bad.lb, res.lb[w0]) and np.array_equal(bad.ub, res.ub[w0])
    assert np.array_equal(bad.nvalues, res.nvalues[w0]) and np.array_equal(bad.values, res.values[w0])
    # ... and on the wires that are NOT unique (what a reader of the listing looks at) it is the oracle's — except on
    # the pinned schedule-dependent wires, which hold exactly their pinned state and print as pinned
    pinned = {e["wire"]: e for e in SCHEDULE_DIFFS.get(name, []) if not e["unique"]}
    nu = ~uniq[w0] & ~np.isin(wires, list(pinned))
    assert np.array_equal(bad.lb[nu], o.lb[w0][nu]) and np.array_equal(bad.ub[nu], o.ub[w0][nu])
    assert np.array_equal(bad.nvalues[nu], o.nvalues[w0][nu])
    assert np.array_equal(bad.values[nu], o.values[w0][nu])
    for w, e in pinned.items():
        k = int(np.flatnonzero(wires == w)[0])  # a pinned non-unique wire is always listed
        assert hex(_toint(bad.lb[k])) == e["engine"]["lb"] and hex(_toint(bad.ub[k])) == e["engine"]["ub"]
        assert bool((bad.flags[k] >> 1) & 1) == e["engine"]["is_known"] and int(bad.nvalues[k]) == e["engine"]["nvalues"]
        # what the listing prints for the wire (:396-419) on either side
        got = api.format_state(False, _toint(bad.lb[k]), _toint(bad.ub[k]), [])
        want_oracle = api.format_state(False, _toint(o.lb[w - 1]), _toint(o.ub[w - 1]), [])
        assert got == f"Uniquely Determined: false\nBounds: [0, {int(e['engine']['ub'], 16)}]\n\n"
        assert want_oracle == "Uniquely Determined: false\nBounds: None\n\n"
    if name in ("circomlib/Bits2Point_Strict@pointbits", "circomlib/EdDSAVerifier@eddsa"):
        assert len(pinned) == (2 if "Bits2Point" in name else 8)
    if len(rows) == 0:  # a sound system with every wire determined lists nothing
        assert len(wires) == 0 and bad.n_bad_rows == 0


@pytest.mark.gpu
def test_report_protocol_errors_and_capacity():
    (reduced, specials, main), secp = _prepare("target/division")
    lib = api._engine()
    ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, secp)
    h = C.c_void_p()
    assert lib.ecne_upload(C.byref(ph.c), C.byref(h)) == 0
    try:
        bits = np.zeros((reduced.n_rows + 63) // 64, dtype=np.uint64)
        rep = _abi.Report()
        rep.bad_row_bits = bits.ctypes.data_as(_abi.u64p)
        assert lib.ecne_report_resident(h, C.byref(rep)) == _abi.ECNE_E_BADARG  # nothing solved yet
        res = api.SolveResult(main.n_vars)
        assert lib.ecne_solve_resident(h, C.byref(res.c)) == 0
        assert lib.ecne_report_resident(h, C.byref(rep)) == 0  # bitmap + counts only
        assert rep.n_wires > 0 and rep.n_bad_rows > 0
        wire = np.zeros(1, dtype=np.uint32)
        rep.wire = wire.ctypes.data_as(_abi.u32p)
        rep.cap_wires = 1
        n = int(rep.n_wires)
        rep.n_wires = 0
        assert lib.ecne_report_resident(h, C.byref(rep)) == _abi.ECNE_E_BADARG  # too small: sizes reported
        assert int(rep.n_wires) == n
        assert lib.ecne_report_resident(None, C.byref(rep)) == _abi.ECNE_E_BADARG
    finally:
        lib.ecne_free_resident(h)


@pytest.mark.gpu
@pytest.mark.parametrize("r1cs,sym", [("target/division.r1cs", "target/division.sym"),
                                      ("bad_bd_check.r1cs", "bad_bd_check.sym"),
                                      ("good_bd_check.r1cs", "good_bd_check.sym"),
                                      ("tornadocash_circuits/merkleTree.r1cs", "tornadocash_circuits/merkleTree.sym")])
def test_printed_listing_matches_oracle_state(r1cs, sym):
    """The text SolveConstraintsSymbolic prints with a .sym (:1558-1644) = the same formatters fed with
    the ORACLE's final state.  merkleTree ends with every wire unique: bounds of unique wires are
    schedule-dependent there, so its "All Variables" part is compared on the first two
    lines of every entry only."""
    reduced, specials, main = api.prepare(fixtures.path(r1cs))
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ok = api.SolveConstraintsSymbolic(reduced, specials, main.known, False, main.targets, main.n_vars,
                                          fixtures.path(sym))
    o = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, False)
    assert ok == o.verdict
    rows, wires, uniq = _expected(reduced, o.unique_bits, main.n_vars)
    w0 = wires.astype(np.int64) - 1
    kb = np.unpackbits(o.known_bits.view(np.uint8), bitorder="little")[:main.n_vars]
    stub = types.SimpleNamespace(rows=rows, wire=wires, flags=(uniq[w0].astype(np.uint8) | (kb[w0] << 1)),
                                 lb=o.lb[w0], ub=o.ub[w0], nvalues=o.nvalues[w0], values=o.values[w0])
    stub.state = types.MethodType(api.BadConstraints.state, stub)
    names = api.read_sym(fixtures.path(sym))
    want = (f"Solved for {o.c.n_unique_nontrivial} variables out of {o.c.n_nontrivial} total variables\n"
            f"Solved for {o.c.n_targets_unique} target variables out of {len(main.targets)} total target variables\n"
            "------ Bad Constraints ------\n\n" + api.format_listing(reduced, stub, o, names))
    got = out.getvalue()
    if "merkleTree" in r1cs:
        strip = lambda t: "\n".join(ln for ln in t.splitlines() if not ln.startswith(("Bounds", "All possible")))
        assert strip(got) == strip(want)
    else:
        assert got == want
    assert "------ All Variables ------" in got
    assert api.last_bad_constraints is not None and len(api.last_bad_constraints.rows) == len(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ecdsa+secp256k1", "ecdsa"])
def test_bad_constraints_full_size(name):
    """BASELINE.json's full sizes (694 k and 1.09 M rows; 78 and 391 133 listed wires) against the pins minted from the
    oracle's final state (tests/golden/make_report_goldens.py); the compacted state against the engine's own full
    export.  The D2H of the compact form is what the report path is for: 138 B per listed wire."""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "report_goldens.json")))[name]
    (reduced, specials, main), secp = _prepare(name)
    res, bad = _solve_with_report(reduced, specials, main, secp)
    assert (bad.n_bad_rows, len(bad.wire)) == (gold["n_bad_rows"], gold["n_wires"])
    row_bytes = bad.row_bits.view(np.uint8)[: (reduced.n_rows + 7) // 8].tobytes()
    assert hashlib.sha256(row_bytes).hexdigest() == gold["sha_rows"]
    assert hashlib.sha256(np.ascontiguousarray(bad.wire).tobytes()).hexdigest() == gold["sha_wires"]
    w0 = bad.wire.astype(np.int64) - 1
    ub = np.unpackbits(res.unique_bits.view(np.uint8), bitorder="little")[:main.n_vars]
    kb = np.unpackbits(res.known_bits.view(np.uint8), bitorder="little")[:main.n_vars]
    assert np.array_equal(bad.flags & 1, ub[w0]) and np.array_equal((bad.flags >> 1) & 1, kb[w0])
    assert np.array_equal(bad.lb, res.lb[w0]) and np.array_equal(bad.ub, res.ub[w0])
    assert np.array_equal(bad.nvalues, res.nvalues[w0]) and np.array_equal(bad.values, res.values[w0])


def test_debug_dump_formats_every_nontrivial_variable():
    """debug=true prints printState of every variable in all_nontrivial_vars (:1573-1577); fed here with the
    oracle's final state of target/division (no GPU): one entry per non-trivial variable, unique flags as solved."""
    reduced, specials, main = api.prepare(fixtures.path("target/division.r1cs"))
    o = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, False)
    txt = api.format_all_states(reduced, o)
    entries = txt.split("\n\n")[:-1]
    nt = api.nontrivial_variables(reduced)
    assert len(entries) == len(nt) == o.c.n_nontrivial
    uniq = np.unpackbits(o.unique_bits.view(np.uint8), bitorder="little")
    for e, w in zip(entries, nt):
        assert e.splitlines()[0] == "Uniquely Determined: " + ("true" if uniq[int(w) - 1] else "false")
    assert sum(e.startswith("Uniquely Determined: true") for e in entries) == o.c.n_unique_nontrivial
