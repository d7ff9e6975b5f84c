"""Host-side callers of the hot path: the .r1cs loader and abstraction() (libecne_host.so), and that
both native libraries load and export every symbol the headers declare."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ecneproject_b200 import _abi, api, fixtures
from helpers import P, from_limbs, py_read_r1cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name):
    txt = open(os.path.join(ROOT, "include", name)).read()
    return sorted(set(re.findall(r"\b(ecne_[a-z0-9_]+)\s*\(", txt)))


def test_engine_exports_every_declared_symbol():
    lib = _abi.engine_lib()
    syms = header_symbols("ecne_abi.h")
    assert set(syms) == set(_abi.ENGINE_SYMBOLS)
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.ecne_version() == _abi.ABI_VERSION == 3
    # the layout table the library reports is what the ctypes mirror has (checked at load time as well)
    n = lib.ecne_abi_layout(None, 0)
    buf = (C.c_uint32 * n)()
    assert lib.ecne_abi_layout(buf, n) == n and list(buf) == _abi.layout_table()
    assert n == 3 * 2 + 22 + 31 + 10


def test_host_exports_every_declared_symbol():
    lib = _abi.host_lib()
    syms = header_symbols("ecne_host.h")
    assert set(syms) == set(_abi.HOST_SYMBOLS)
    for s in syms:
        assert getattr(lib, s) is not None


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.engine_lib()
    assert lib.ecne_init(0) == _abi.ECNE_E_CUDA
    assert b"no CPU fallback" in lib.ecne_last_error()
    r = api.readR1CS(fixtures.path("trivial_mult.r1cs"))
    ph = api.ProblemHandle(r, [], r.known, r.targets, r.n_vars)
    res = api.SolveResult(r.n_vars)
    assert lib.ecne_solve(C.byref(ph.c), C.byref(res.c)) == _abi.ECNE_E_CUDA


@pytest.mark.parametrize("rel", ["trivial_mult.r1cs", "target/division.r1cs", "multiplexer_33.r1cs",
                                 "poseidon.r1cs", "ecne_circomlib_tests/Num2Bits_strict@bitify.r1cs",
                                 "ecne_circomlib_tests/Poseidon@poseidon.r1cs", "biglessthan.r1cs"])
def test_loader_matches_python_restatement(rel):
    path = fixtures.path(rel)
    r = api.readR1CS(path)
    rows, known, targets, n_vars = py_read_r1cs(path)
    assert r.n_rows == len(rows) and r.n_vars == n_vars
    assert r.known.tolist() == known and r.targets.tolist() == targets
    coef = from_limbs(r.coef)
    for i, forms in enumerate(rows):
        for f, d in enumerate(forms):
            s, e = int(r.seg_ptr[3 * i + f]), int(r.seg_ptr[3 * i + f + 1])
            got = {int(r.col[t]): coef[t] for t in range(s, e)}
            assert got == d, (i, f)


def test_loader_matches_python_restatement_on_the_whole_corpus():
    """Every .r1cs the reference ships but ecdsa (91 files, 231 757 rows): rows, stored keys (explicit zeros included),
    reduced coefficients, known / target wires against the independent Python parser."""
    import os
    root = os.path.dirname(fixtures.path("trivial_mult.r1cs"))
    files = sorted(os.path.relpath(os.path.join(dp, f), root) for dp, _, fn in os.walk(root) for f in fn
                   if f.endswith(".r1cs") and f != "ecdsa.r1cs")
    assert len(files) >= 90
    total = 0
    for rel in files:
        path = fixtures.path(rel)
        r = api.readR1CS(path)
        rows, known, targets, n_vars = py_read_r1cs(path)
        assert r.n_rows == len(rows) and r.n_vars == n_vars, rel
        assert r.known.tolist() == known and r.targets.tolist() == targets, rel
        col, seg = r.col.tolist(), r.seg_ptr.tolist()
        raw = np.ascontiguousarray(r.coef).view(np.uint8).reshape(-1, 32)
        for i, forms in enumerate(rows):
            for f, d in enumerate(forms):
                got = {col[k]: int.from_bytes(raw[k].tobytes(), "little") for k in range(seg[3 * i + f], seg[3 * i + f + 1])}
                assert got == d, (rel, i, f)
        total += r.n_rows
    assert total > 200000


def test_loader_errors():
    lib = _abi.host_lib()
    out = C.POINTER(_abi.R1CSStruct)()
    assert lib.ecne_read_r1cs(b"/nonexistent/x.r1cs", C.byref(out)) == _abi.ECNE_E_IO
    raw = bytearray(open(fixtures.path("trivial_mult.r1cs"), "rb").read())
    bad = bytes(raw[:4]) + (2).to_bytes(4, "little") + bytes(raw[8:])
    buf = (C.c_uint8 * len(bad)).from_buffer_copy(bad)
    assert lib.ecne_read_r1cs_mem(buf, len(bad), C.byref(out)) == _abi.ECNE_E_ASSERT  # version (:58)
    bad = bytes(raw[:8]) + (4).to_bytes(4, "little") + bytes(raw[12:])
    buf = (C.c_uint8 * len(bad)).from_buffer_copy(bad)
    assert lib.ecne_read_r1cs_mem(buf, len(bad), C.byref(out)) == _abi.ECNE_E_ASSERT  # sections (:62)
    with pytest.raises(OSError):
        api.readR1CS("/nonexistent/x.r1cs")


def test_abstraction_match_counts():
    # bench/bench_abstraction.jl:16 asserts exactly one match of bigmultshortlong in bigmultmodp86_3
    main = api.readR1CS(fixtures.path("bigmultmodp86_3.r1cs"))
    sub = api.readR1CS(fixtures.path("bigmultshortlong86_3.r1cs"))
    sp, red = api.abstraction("bigmultmodp", main, sub)
    assert len(sp) == 1
    assert red.n_rows == main.n_rows - sub.n_rows
    name, ins, outs = sp.as_list()[0]
    assert len(ins) == len(sub.known) - 1 and len(outs) == len(sub.targets)


def test_abstraction_trusted_configs():
    # reduced sizes / special counts of SURVEY.md §8a
    red, sp, main = api.prepare(fixtures.path("secp256k1.r1cs"),
                                [fixtures.path("bigmultmodp.r1cs"), fixtures.path("biglessthan.r1cs")],
                                ["BigMultModP", "BigLessThan"])
    assert (main.n_rows, red.n_rows, len(sp)) == (15935, 3985, 4)
    assert [n for n, _, _ in sp.as_list()] == ["BigMultModP"] * 3 + ["BigLessThan"]
    ped = ["tornadocash_circuits/Pedersen248@pedersen.r1cs", "tornadocash_circuits/Pedersen496@pedersen.r1cs"]
    red, sp, main = api.prepare(fixtures.path("tornadocash_circuits/withdraw.r1cs"),
                                [fixtures.path(p) for p in ped], ["Pedersen248", "Pedersen496"])
    assert (main.n_rows, red.n_rows, len(sp)) == (24129, 1996, 2)
    # longest trusted circuit is abstracted first (:527)
    assert sp.as_list()[0][0] == "Pedersen496"


def test_abstraction_no_match_is_identity():
    main = api.readR1CS(fixtures.path("poseidon.r1cs"))
    sub = api.readR1CS(fixtures.path("multiplexer_33.r1cs"))
    sp, red = api.abstraction("x", main, sub)
    assert len(sp) == 0 and red.n_rows == main.n_rows
    assert np.array_equal(red.col, main.col) and np.array_equal(red.coef, main.coef)


def _mk_r1cs(rows, n_wires, pub_out=1, pub_in=1, prv_in=1):
    """Serialise rows = [(A, B, C)] with each form a list of (wire0, int) into the iden3 .r1cs layout
    (ParseR1CS.jl:50-124); wire ids are 0-based on disk."""
    chunks = []
    for forms in rows:
        for form in forms:
            chunks.append(len(form).to_bytes(4, "little"))
            for w, c in form:
                chunks.append(w.to_bytes(4, "little") + (c % (1 << 256)).to_bytes(32, "little"))
    cons = b"".join(chunks)
    prime = (21888242871839275222246405745257275088548364400416034343698204186575808495617).to_bytes(32, "little")
    hdr = (32).to_bytes(4, "little") + prime + n_wires.to_bytes(4, "little") + pub_out.to_bytes(4, "little") + \
        pub_in.to_bytes(4, "little") + prv_in.to_bytes(4, "little") + (n_wires).to_bytes(8, "little") + \
        len(rows).to_bytes(4, "little")
    wmap = b"".join(i.to_bytes(8, "little") for i in range(n_wires))
    out = b"r1cs" + (1).to_bytes(4, "little") + (3).to_bytes(4, "little")
    for ty, body in ((1, hdr), (2, cons), (3, wmap)):
        out += ty.to_bytes(4, "little") + len(body).to_bytes(8, "little") + body
    return out


def _read_mem(blob):
    lib = _abi.host_lib()
    out = C.POINTER(_abi.R1CSStruct)()
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    st = lib.ecne_read_r1cs_mem(buf, len(blob), C.byref(out))
    return st, (api.R1CS(out) if st == 0 else None)


def test_loader_edge_forms():
    """Empty forms store an explicit zero on key 1 (ParseR1CS.jl:113-115), coefficients >= p are reduced
    (:111), a wire repeated inside one form keeps its LAST value (Dict assignment, :110) — the last case takes
    the serial path of the loader, the others its parallel fast path."""
    from helpers import P
    rows = [
        ([(1, 3)], [(2, 5)], []),                                  # empty C
        ([], [], [(3, P + 7), (1, 2 * P + 1)]),                    # values above p
        ([(2, 1), (2, 9)], [(1, 1)], [(3, 1), (1, 4), (3, 6)]),    # repeated wires
    ] + [([(1, 1)], [(2, 1)], [(k % 4, k + 1) for k in range(12)])]  # a longer form with repeats (sort path)
    st, r = _read_mem(_mk_r1cs(rows, n_wires=4))
    assert st == 0
    coef = from_limbs(r.coef)

    def form(i, f):
        s, e = int(r.seg_ptr[3 * i + f]), int(r.seg_ptr[3 * i + f + 1])
        return {int(r.col[t]): coef[t] for t in range(s, e)}, e - s
    assert form(0, 2) == ({1: 0}, 1)
    assert form(1, 0) == ({1: 0}, 1) and form(1, 1) == ({1: 0}, 1)
    assert form(1, 2)[0] == {4: 7, 2: 1}
    assert form(2, 0) == ({3: 9}, 1)                 # the later value wins, one key
    assert form(2, 2)[0] == {4: 6, 2: 4}
    assert form(3, 2)[0] == {1: 9, 2: 10, 3: 11, 4: 12}
    assert r.known.tolist() == [1, 3, 4] and r.targets.tolist() == [2]
    # the same content without repeated wires goes through the fast path and must agree form by form
    st2, r2 = _read_mem(_mk_r1cs(rows[:2], n_wires=4))
    assert st2 == 0 and r2.n_rows == 2
    assert r2.seg_ptr.tolist() == r.seg_ptr[:7].tolist()
    assert r2.col.tolist() == r.col[:int(r.seg_ptr[6])].tolist()


def test_loader_truncated_file_is_a_bounds_error():
    blob = _mk_r1cs([([(1, 3)], [(2, 5)], [(3, 1)])] * 4, n_wires=4)
    for cut in (len(blob) - 8 * 4 - 20, len(blob) - 8 * 4 - 12 - 36 * 5, 40, 11):
        st, _ = _read_mem(blob[:cut])
        assert st in (_abi.ECNE_E_BOUNDS, _abi.ECNE_E_ASSERT), (cut, st)


def test_loader_hostile_sizes_are_bounds_errors_not_crashes():
    """Sizes read from the file never index or allocate unchecked: a section size that wraps the 64-bit cursor
    of the section-table walk, and a header announcing 2^32-1 / 2^28 constraints over an empty constraint
    section, are BoundsErrors (the reference raises a catchable BoundsError) — not a segfault or std::terminate."""
    good = _mk_r1cs([([(1, 3)], [(2, 5)], [(3, 1)])], n_wires=4)
    # (1) first section's size brings the cursor to 2^64 - 8: `cur + 12` wraps to 4
    blob = bytearray(good)
    blob[16:24] = ((1 << 64) - 8 - 12 - 12).to_bytes(8, "little")
    st, _ = _read_mem(bytes(blob))
    assert st == _abi.ECNE_E_BOUNDS, st
    # (2) nConstraints from the header with nothing behind it
    for n_cons in (0xFFFFFFFF, 0x10000000, 1 << 20):
        prime = (21888242871839275222246405745257275088548364400416034343698204186575808495617).to_bytes(32, "little")
        hdr = (32).to_bytes(4, "little") + prime + (4).to_bytes(4, "little") + (1).to_bytes(4, "little") * 3 + \
            (4).to_bytes(8, "little") + n_cons.to_bytes(4, "little")
        out = b"r1cs" + (1).to_bytes(4, "little") + (3).to_bytes(4, "little")
        for ty, body in ((1, hdr), (2, b""), (3, b"")):
            out += ty.to_bytes(4, "little") + len(body).to_bytes(8, "little") + body
        st, _ = _read_mem(out)
        assert st == _abi.ECNE_E_BOUNDS, (n_cons, st)


def test_abstraction_goldens():
    """Reduced system + special constraints of every trusted-function configuration (ecdsa included) against
    the committed pins (tests/golden/make_abstraction_goldens.py says what they are and are not)."""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_abstraction_goldens",
                                                  os.path.join(here, "golden", "make_abstraction_goldens.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "golden", "abstraction_goldens.json")))
    got = mod.mint()
    assert got == want
    assert want["ecdsa+secp256k1"]["specials"] == 25 and want["ecdsa+secp256k1"]["rows"] == 694264


def test_abstraction_only_prints_the_specials(capsys):
    """:545-548: abstractionOnly prints the special constraints and returns true before the solver (no GPU needed)."""
    ped = [fixtures.path("tornadocash_circuits/Pedersen248@pedersen.r1cs"), fixtures.path("tornadocash_circuits/Pedersen496@pedersen.r1cs")]
    assert api.solveWithTrustedFunctions(fixtures.path("tornadocash_circuits/withdraw.r1cs"), "Withdraw", trusted_r1cs=ped,
                                         trusted_r1cs_names=["Pedersen248", "Pedersen496"], abstractionOnly=True) is True
    out = capsys.readouterr().out
    assert "Pedersen496" in out and "Pedersen248" in out


def test_c_example_compiles_as_strict_c99_and_fails_loudly_without_a_gpu(tmp_path):
    """include/*.h are C headers (extern "C", plain pointers and sizes): examples/solve_r1cs.c builds with
    `gcc -std=c99 -pedantic -Werror` against the two libraries, runs the host side (reader + abstraction) and — in
    this container, without a CUDA device — stops at ecne_solve with ECNE_E_CUDA instead of computing on the CPU."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ecneproject_b200")
    exe = str(tmp_path / "solve_r1cs")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "solve_r1cs.c"), "-L", pkg, "-lecne_host", "-lecne_b200",
                    "-Wl,-rpath," + pkg, "-o", exe], check=True, capture_output=True, text=True)
    p = subprocess.run([exe, "--secp-solve", fixtures.path("secp256k1.r1cs"), fixtures.path("bigmultmodp.r1cs"), "BigMultModP",
                        fixtures.path("biglessthan.r1cs"), "BigLessThan"], capture_output=True, text=True, timeout=120)
    assert "BigMultModP: 3 window(s) abstracted, 4298 rows left" in p.stdout
    assert "BigLessThan: 1 window(s) abstracted, 3985 rows left" in p.stdout
    import torch
    if torch.cuda.is_available():
        assert p.returncode == 0 and "sound constraints" in p.stdout   # test/runtests.jl:35
    else:
        assert p.returncode == 2 and "no CPU fallback" in p.stderr


def test_loader_parallel_offset_walk_on_ecdsa():
    """Files with >= 2^18 forms take the speculative parallel offset walk (csrc/host_r1cs.cpp); its result on
    ecdsa.r1cs (3.3 M forms) is pinned to the digest the serial walk produced."""
    import hashlib
    r = api.readR1CS(fixtures.path("ecdsa.r1cs"))
    h = hashlib.sha256()
    for a in (r.seg_ptr, r.col, r.coef, r.known, r.targets):
        h.update(np.ascontiguousarray(a).tobytes())
    assert (r.n_rows, r.n_vars, r.nnz) == (1092639, 1089136, 4797431)
    assert h.hexdigest() == "b44a391a1612b970f447c78ca6c7e2beacd25651959d600500bce130c7629d07"


def test_loader_parallel_offset_walk_on_a_hostile_file(tmp_path):
    """A synthetic file big enough for the parallel walk and built to mislead its guesses: coefficients that are
    mostly zero bytes (they read as runs of empty forms), many truly empty forms, wire ids and term counts that look
    alike, and a few forms longer than everything around them.  Against the independent Python parser; then the same
    file truncated must still be a bounds error."""
    rng = np.random.default_rng(7)
    n_wires = 50
    coefs = [1, 2, 3, 0, P - 1, 1 << 64, (1 << 200) + 5]
    rows = []
    for i in range(90000):
        forms = []
        for _ in range(3):
            k = int(rng.integers(0, 4))
            ws = rng.choice(n_wires, size=k, replace=False)
            forms.append([(int(w), coefs[int(rng.integers(0, len(coefs)))]) for w in ws])
        if i % 20011 == 7:   # a long form (no repeated wire: those take the serial path by design)
            forms[2] = [(int(w), 1) for w in range(n_wires)]
        rows.append(tuple(forms))
    blob = _mk_r1cs(rows, n_wires=n_wires)
    path = tmp_path / "hostile.r1cs"
    path.write_bytes(blob)
    r = api.readR1CS(str(path))
    want, known, targets, n_vars = py_read_r1cs(str(path))
    assert r.n_rows == len(want) == 90000 and r.n_vars == n_vars
    col, seg = r.col.tolist(), r.seg_ptr.tolist()
    raw = np.ascontiguousarray(r.coef).view(np.uint8).reshape(-1, 32)
    for i, forms in enumerate(want):
        for f, d in enumerate(forms):
            got = {col[k]: int.from_bytes(raw[k].tobytes(), "little") for k in range(seg[3 * i + f], seg[3 * i + f + 1])}
            assert got == d, (i, f)
    # any number of byte ranges (= guessed starts) gives the same arrays
    import os
    try:
        for ranges in ("2", "3", "7", "61", "256"):
            os.environ["ECNE_HOST_WALK_RANGES"] = ranges
            r2 = api.readR1CS(str(path))
            assert np.array_equal(r2.seg_ptr, r.seg_ptr) and np.array_equal(r2.col, r.col) and np.array_equal(r2.coef, r.coef)
    finally:
        os.environ.pop("ECNE_HOST_WALK_RANGES", None)
    st, _ = _read_mem(blob[: len(blob) // 2])
    assert st in (_abi.ECNE_E_BOUNDS, _abi.ECNE_E_ASSERT)

def test_compact_coef_matches_a_numpy_classification():
    """ecne_compact_coef (include/ecne_host.h): class bytes, other values and their term indices."""
    rng = np.random.default_rng(7)
    n = 200_003
    coef = np.zeros((n, 4), dtype=np.uint64)
    kind = rng.integers(0, 5, n)
    pm1 = np.array([((P - 1) >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)
    coef[kind == 1, 0] = 1
    coef[kind == 2] = pm1
    coef[kind == 3] = rng.integers(0, 2**63, (int((kind == 3).sum()), 4), dtype=np.uint64)
    coef[kind == 4, 0] = 2          # small but not 0/1
    coef[0] = pm1; coef[0, 0] -= 1  # p - 2: shares three limbs with p - 1
    cls, other, term = api.compact_coef(coef)
    want = np.where(kind == 0, 0, np.where(kind == 1, 1, np.where(kind == 2, 2, 3))).astype(np.uint8)
    want[0] = 3
    assert np.array_equal(cls, want)
    assert np.array_equal(term, np.flatnonzero(want == 3).astype(np.uint32))
    assert np.array_equal(other, coef[want == 3])
    cls0, other0, term0 = api.compact_coef(np.zeros((0, 4), dtype=np.uint64))
    assert len(cls0) == 0 and len(term0) == 0


def test_compact_only_read_matches_the_full_read():
    """ecne_read_r1cs_opts(ECNE_READ_COMPACT_ONLY) (include/ecne_host.h): same rows, same compact form, no 32-byte
    coefficient array; the compact form the reader emits equals ecne_compact_coef of the full coefficients."""
    for name in ("ecne_circomlib_tests/EdDSAVerifier@eddsa.r1cs", "bigmultmodp86_3.r1cs", "bad_bd_check.r1cs", "secp256k1.r1cs"):
        path = fixtures.path(name)
        full = api.readR1CS(path)
        lean = api.readR1CS(path, compact_only=True)
        assert lean.coef is None and full.coef is not None
        assert np.array_equal(full.seg_ptr, lean.seg_ptr) and np.array_equal(full.col, lean.col)
        assert full.compact is not None and lean.compact is not None
        for a, b in zip(full.compact, lean.compact):
            assert np.array_equal(a, b)
        cls, other, term = api.compact_coef(full.coef)
        assert np.array_equal(cls, lean.compact[0])
        assert np.array_equal(other.reshape(-1), lean.compact[1]) and np.array_equal(term, lean.compact[2])
        assert np.array_equal(lean.compact[3].astype(np.uint64), full.seg_ptr)
        assert (lean.known == full.known).all() and (lean.targets == full.targets).all() and lean.n_vars == full.n_vars


def test_freed_arrays_are_recycled_and_trim_releases_them():
    """include/ecne_host.h: the large arrays of a freed system are handed to the next read (same results, and the second
    read of a file gets the first one's blocks back); ecne_host_trim() empties the cache."""
    path = fixtures.path("secp256k1.r1cs")
    lib = _abi.host_lib()
    lib.ecne_host_trim.restype = None
    a = api.readR1CS(path)
    col0, seg0, coef0 = a.col.copy(), a.seg_ptr.copy(), a.coef.copy()
    addr = a.col.ctypes.data
    del a
    import gc
    gc.collect()
    b = api.readR1CS(path)
    assert np.array_equal(b.col, col0) and np.array_equal(b.seg_ptr, seg0) and np.array_equal(b.coef, coef0)
    if b.nnz * 4 >= (4 << 20):          # (arrays below 4 MB are plain malloc)
        assert b.col.ctypes.data == addr
    del b
    gc.collect()
    lib.ecne_host_trim()
    c = api.readR1CS(path)
    assert np.array_equal(c.col, col0) and np.array_equal(c.coef, coef0)


def test_loader_range_guesses_do_not_fall_for_periodic_or_giant_trails(tmp_path):
    """The parallel offset walk guesses the first header of every byte range and proves the guess afterwards; a wrong
    guess is repaired by walking serially (slow, never wrong).  Two families of false trails that pass 64 plausible
    headers: (1) in a run of one-term forms with coefficient 1 (a * b = c rows) the chain that starts 8 bytes late reads
    "one term, wire 0" for ever; (2) a coefficient whose low word reads as a term count of tens of thousands and lands
    on a true header far away.  Both must be rejected up front: with any number of ranges the arrays are the serial
    ones AND (almost) nothing is walked serially to repair guesses — ecdsa.r1cs lost a sixteenth of its forms to (2)."""
    import subprocess, sys, textwrap
    n_wires = 200000
    rows = []
    for i in range(100000):
        a, b, c = 1 + (3 * i) % (n_wires - 1), 1 + (3 * i + 1) % (n_wires - 1), 1 + (3 * i + 2) % (n_wires - 1)
        if i % 500 == 499:    # family (2): low word 100 000 = "100 000 terms" = 3.6 MB ahead
            rows.append(([(a, 100000 + (1 << 64))], [(b, 1)], [(c, 1)]))
        else:                 # family (1)
            rows.append(([(a, 1)], [(b, 1)], [(c, 1)]))
    path = tmp_path / "trails.r1cs"
    path.write_bytes(_mk_r1cs(rows, n_wires=n_wires))
    ref = None
    for ranges in ("2", "5", "16", "37", "128"):
        code = textwrap.dedent(f"""
            import sys, hashlib
            sys.path.insert(0, {str(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))!r})
            from ecneproject_b200 import api
            r = api.readR1CS({str(path)!r})
            h = hashlib.sha256(r.seg_ptr.tobytes() + r.col.tobytes() + r.coef.tobytes()).hexdigest()
            print("SHA", h, r.n_rows, r.nnz)
        """)
        env = dict(os.environ, ECNE_HOST_WALK_RANGES=ranges, ECNE_HOST_PROF="1")
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        sha = [l for l in out.stdout.splitlines() if l.startswith("SHA")][0]
        ref = ref or sha
        assert sha == ref, ranges
        line = [l for l in out.stderr.splitlines() if "forms walked serially" in l]
        assert line and "proof ok" in line[0], out.stderr[-2000:]
        repaired = int(re.search(r"(\d+) forms walked serially", line[0]).group(1))
        assert repaired < 2000, (ranges, line[0])      # (a whole range would be 300 000 / ranges)
    r = api.readR1CS(str(path))
    assert r.n_rows == 100000 and r.nnz == 300000
    want, _, _, _ = py_read_r1cs(str(path))
    col, seg = r.col.tolist(), r.seg_ptr.tolist()
    raw = np.ascontiguousarray(r.coef).view(np.uint8).reshape(-1, 32)
    for i in list(range(0, 100000, 997)) + [499, 999, 99999]:
        for f, d in enumerate(want[i]):
            got = {col[k]: int.from_bytes(raw[k].tobytes(), "little") for k in range(seg[3 * i + f], seg[3 * i + f + 1])}
            assert got == d, (i, f)
