"""ecne_abstraction (libecne_host.so, CSR + flat signatures, multi-threaded) against oracle/abstraction_ref.py, a
plain-Python restatement of R1CSConstraintSolver.jl:237-395 on per-row dictionaries: the special constraints
(name, mapped inputs, mapped outputs) and the reduced system row by row, on every trusted-function configuration
the reference runs (ecdsa included) and on hand-built systems for the quirks of the pass — the match that starts
inside a consumed window and stalls the walk (:370), the last row that is not hashed (:261) but is verified (:301),
wires with identical signatures, the KeyError of :381-382."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import abstraction_ref as ref  # noqa: E402
from configs import CONFIGS  # noqa: E402
from ecneproject_b200 import _abi, api, fixtures  # noqa: E402
from helpers import P  # noqa: E402
from test_host_side import _mk_r1cs, _read_mem  # noqa: E402

KIND_NAME = {"BigMultModP": "BigMultModP", "BigLessThan": "BigLessThan"}


def rows_of(r):
    """An api.R1CS as the reference holds it: per row three dicts wire -> coefficient (stored keys, zeros included)."""
    seg = r.seg_ptr.astype(np.int64)
    col = r.col.tolist()
    raw = np.ascontiguousarray(r.coef).view(np.uint8).reshape(-1, 32)
    coef = [int.from_bytes(raw[t].tobytes(), "little") for t in range(len(col))]
    out = []
    for i in range(r.n_rows):
        out.append(tuple({col[t]: coef[t] for t in range(seg[3 * i + f], seg[3 * i + f + 1])} for f in range(3)))
    return out


def check_chain(main, subs):
    """The loop of solveWithTrustedFunctions (:527-544): abstract `subs` (name, R1CS) one after the other."""
    want_rows, want_specials = rows_of(main), []
    reduced, specials = main, api.Specials()
    for name, sub in subs:
        sp, want_rows = ref.abstraction(name, want_rows, sub.known.tolist(), rows_of(sub), sub.targets.tolist())
        want_specials += sp
        specials, reduced = api.abstraction(name, reduced, sub, specials)
    got = [(n, [int(x) for x in i], [int(x) for x in o]) for n, i, o in specials.as_list()]
    assert got == [(n, list(i), list(o)) for n, i, o in want_specials]
    assert reduced.n_rows == len(want_rows)
    assert rows_of(reduced) == want_rows
    assert reduced.known.tolist() == main.known.tolist() and reduced.targets.tolist() == main.targets.tolist()
    return got, reduced


@pytest.mark.parametrize("name", [n for n, c in CONFIGS.items() if c.get("trusted")])
def test_trusted_configurations(name):
    cfg = CONFIGS[name]
    main = api.readR1CS(fixtures.path(cfg["main"]))
    subs = [(cfg["trusted_names"][i], api.readR1CS(fixtures.path(t))) for i, t in enumerate(cfg["trusted"])]
    subs.sort(key=lambda x: -len(x[1]))  # stable, longest first (:527)
    got, reduced = check_chain(main, subs)
    assert len(got) >= 1  # bench/bench_abstraction.jl:24


def test_bench_abstraction_pair():
    # bench/bench_abstraction.jl:13-17: exactly one bigmultshortlong inside bigmultmodp86_3
    main = api.readR1CS(fixtures.path("bigmultmodp86_3.r1cs"))
    sub = api.readR1CS(fixtures.path("bigmultshortlong86_3.r1cs"))
    got, _ = check_chain(main, [("bigmultmodp", sub)])
    assert len(got) == 1


# ---- hand-built systems --------------------------------------------------------------------------
def mk(rows, n_wires, pub_out=1, pub_in=1, prv_in=0):
    st, r = _read_mem(_mk_r1cs(rows, n_wires=n_wires, pub_out=pub_out, pub_in=pub_in, prv_in=prv_in))
    assert st == 0
    return r


def mul(a, b, c):
    """(1*a) * (1*b) = (1*c) on 0-based wires."""
    return ([(a, 1)], [(b, 1)], [(c, 1)])


def test_overlapping_match_stalls_the_walk():
    """sub = two chained multiplications out = (in*in)*in; main = a chain of 4 multiplications: windows at rows 0, 1, 2
    all verify.  The walk consumes [0,2), then matches[1] starts at row 1 < 2: `i != matches[cur_idx][1]` holds for
    ever, so the perfectly good window at rows 2-3 is NOT abstracted (:368-388)."""
    sub = mk([mul(2, 2, 3), mul(3, 2, 1)], n_wires=4)             # wires: 0 one, 1 out, 2 in, 3 tmp
    main = mk([mul(2, 2, 3), mul(3, 2, 4), mul(4, 2, 5), mul(5, 2, 1), mul(1, 1, 6)], n_wires=7)
    got, reduced = check_chain(main, [("cube", sub)])
    assert len(got) == 1 and reduced.n_rows == 3
    assert got[0] == ("cube", [3], [5])  # 1-based: in = wire 3, out = the second product (wire index 4 -> 5)


def test_non_overlapping_matches_are_all_taken():
    sub = mk([mul(2, 2, 3), mul(3, 3, 1)], n_wires=4)             # out = (in^2)^2
    main = mk([mul(2, 2, 3), mul(3, 3, 4), mul(1, 2, 7), mul(4, 4, 5), mul(5, 5, 6)], n_wires=8)
    got, reduced = check_chain(main, [("pow4", sub)])
    assert [g[1:] for g in got] == [([3], [5]), ([5], [7])] and reduced.n_rows == 1


def test_last_row_is_verified_although_not_hashed():
    """The candidate scan compares the first n-1 hashes only (:261); the n-th row is still checked by
    checkNonZeroValues (:301-312): a window whose last row has another coefficient is no match."""
    sub = mk([mul(2, 2, 3), ([(3, 1)], [(2, 1)], [(1, 1)])], n_wires=4)
    main = mk([mul(2, 2, 3), ([(3, 1)], [(2, 1)], [(4, 5)]), mul(2, 2, 5), ([(5, 1)], [(2, 1)], [(1, 1)])], n_wires=6)
    got, reduced = check_chain(main, [("cube", sub)])
    assert len(got) == 1 and reduced.n_rows == 2 and got[0][1:] == ([3], [2])


def test_wires_with_identical_signatures():
    """Two wires that appear in exactly the same slots with the same coefficients (a + b in one form): the reference's
    tie order is its Dict's hash order (:334-335, unpinned); both sides here break ties by wire id."""
    sub = mk([([(2, 1), (3, 1)], [(0, 1)], [(1, 1)])], n_wires=4, pub_in=2)            # out = a + b
    main = mk([mul(4, 4, 5), ([(5, 1), (4, 1)], [(0, 1)], [(6, 1)]), mul(6, 6, 1)], n_wires=7)
    got, reduced = check_chain(main, [("add", sub)])
    assert got == [("add", [5, 6], [7])] and reduced.n_rows == 2


def test_keyerror_when_a_trusted_input_never_appears():
    """:381 indexes the wire map with every known input of the trusted circuit: an input with no non-zero term raises."""
    sub = mk([mul(2, 2, 1)], n_wires=4, pub_in=2)     # wire 3 (0-based) is a declared input that no row mentions
    main = mk([mul(2, 2, 3), mul(3, 3, 1)], n_wires=4)
    with pytest.raises(ref.KeyErrorAt):
        ref.abstraction("sq", rows_of(main), sub.known.tolist(), rows_of(sub), sub.targets.tolist())
    with pytest.raises(KeyError):
        api.abstraction("sq", main, sub)


def test_explicit_zero_terms_do_not_count():
    """A stored zero coefficient is dropped by the hash (:231), by checkNonZeroValues (:213) and by the appearance
    map (:284): a window that differs from the trusted circuit only by explicit zeros matches."""
    sub = mk([mul(2, 2, 3), mul(3, 2, 1)], n_wires=4)
    main = mk([([(2, 1), (4, 0)], [(2, 1)], [(3, 1)]), ([(3, 1)], [(2, 1), (0, 0)], [(1, 1)])], n_wires=5)
    got, reduced = check_chain(main, [("cube", sub)])
    assert got == [("cube", [3], [2])] and reduced.n_rows == 0


def _random_case(rng):
    """A random trusted circuit (n rows over a few wires, small coefficient alphabet so that signatures tie and
    unrelated rows collide in hash) planted 0-3 times under random wire renamings into a main circuit between random
    filler rows; sometimes two copies overlap by construction of the filler (identical shapes)."""
    coefs = [1, 1, 1, P - 1, 2, 0]
    nw_sub = int(rng.integers(3, 7))               # wires 0..nw_sub-1 (0 = constant one)
    n = int(rng.integers(1, 4))

    def rand_form(nw, maxlen=3):
        k = int(rng.integers(0, maxlen + 1))
        ws = rng.choice(nw, size=min(k, nw), replace=False).tolist()
        return [(int(w), int(coefs[int(rng.integers(0, len(coefs)))])) for w in ws]
    sub_rows = [(rand_form(nw_sub), rand_form(nw_sub), rand_form(nw_sub)) for _ in range(n)]
    pub_out = 1
    pub_in = int(rng.integers(1, nw_sub - 1))
    main_rows, next_wire = [], 1
    n_main_wires = 1
    for _ in range(int(rng.integers(1, 6))):
        kind = int(rng.integers(0, 3))
        if kind == 0:                               # filler row over the wires allocated so far (+ a fresh one)
            n_main_wires += 1
            main_rows.append((rand_form(n_main_wires), rand_form(n_main_wires), rand_form(n_main_wires)))
        else:                                       # a planted copy: wire w of the sub -> fresh or reused main wire
            ren = {0: 0}
            for w in range(1, nw_sub):
                if n_main_wires > 1 and rng.random() < 0.3:
                    cand = int(rng.integers(1, n_main_wires))
                    if cand in ren.values():
                        cand = n_main_wires
                        n_main_wires += 1
                    ren[w] = cand
                else:
                    ren[w] = n_main_wires
                    n_main_wires += 1
            for forms in sub_rows:
                main_rows.append(tuple([(ren[w], c) for w, c in form] for form in forms))
    return sub_rows, nw_sub, pub_out, pub_in, main_rows, n_main_wires + 1


def test_random_planted_circuits():
    rng = np.random.default_rng(20261017)
    n_matches = n_keyerr = 0
    for case in range(300):
        sub_rows, nw_sub, pub_out, pub_in, main_rows, nw_main = _random_case(rng)
        sub = mk(sub_rows, n_wires=nw_sub, pub_out=pub_out, pub_in=pub_in, prv_in=0)
        main = mk(main_rows, n_wires=nw_main)
        try:
            want_sp, want_rows = ref.abstraction("f", rows_of(main), sub.known.tolist(), rows_of(sub), sub.targets.tolist())
        except ref.KeyErrorAt:
            n_keyerr += 1
            with pytest.raises(KeyError):
                api.abstraction("f", main, sub)
            continue
        sp, red = api.abstraction("f", main, sub)
        got = [(n, [int(x) for x in i], [int(x) for x in o]) for n, i, o in sp.as_list()]
        assert got == [(n, list(i), list(o)) for n, i, o in want_sp], case
        assert rows_of(red) == want_rows, case
        n_matches += len(got)
    assert n_matches > 100 and n_keyerr > 0  # the generator does exercise matches and the KeyError path
