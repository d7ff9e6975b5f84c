"""abstraction() on the device (csrc/abstraction.cu) against the two independent restatements of
R1CSConstraintSolver.jl:237-395: oracle/abstraction_ref.py (plain Python on per-row dictionaries) on the hand-built
quirk cases and 300 random planted circuits, and the host library's ecne_abstraction (itself pinned to the oracle and
to tests/golden/abstraction_goldens.json) on every trusted-function configuration, ecdsa included — special
constraints and reduced rows must be identical; then the solve on the device-abstracted system must reproduce the
oracle goldens."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import abstraction_ref as ref  # noqa: E402
from configs import CONFIGS  # noqa: E402
from ecneproject_b200 import _abi, api, fixtures  # noqa: E402
from test_abstraction_parity import _random_case, mk, mul, rows_of  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_goldens.json")))


class Flat:
    """(seg, col, coef) triple with the attributes rows_of() reads."""

    def __init__(self, seg, col, coef):
        self.seg_ptr, self.col, self.coef = seg, col, coef
        self.n_rows = (len(seg) - 1) // 3


def device_chain(main, subs):
    da = api.DeviceAbstraction(main)
    try:
        counts = [da.apply(name, sub) for name, sub in subs]
        (seg, col, coef), sp = da.export()
    finally:
        da.free()
    return Flat(seg, col, coef), sp, counts


def check_against_oracle(main, subs):
    want_rows, want_sp = rows_of(main), []
    for name, sub in subs:
        sp, want_rows = ref.abstraction(name, want_rows, sub.known.tolist(), rows_of(sub), sub.targets.tolist())
        want_sp += sp
    red, sp, _ = device_chain(main, subs)
    assert sp == [(n, list(i), list(o)) for n, i, o in want_sp]
    assert red.n_rows == len(want_rows)
    assert rows_of(red) == want_rows
    return sp, red


@pytest.mark.parametrize("name", [n for n, c in CONFIGS.items() if c.get("trusted")])
def test_trusted_configurations_match_the_host_library(name):
    cfg = CONFIGS[name]
    main = api.readR1CS(fixtures.path(cfg["main"]))
    subs = [(cfg["trusted_names"][i], api.readR1CS(fixtures.path(t))) for i, t in enumerate(cfg["trusted"])]
    subs.sort(key=lambda x: -len(x[1]))  # stable, longest first (:527)
    reduced, specials = main, api.Specials()
    for nm, sub in subs:
        specials, reduced = api.abstraction(nm, reduced, sub, specials)
    red, sp, counts = device_chain(main, subs)
    assert sp == [(n, [int(x) for x in i], [int(x) for x in o]) for n, i, o in specials.as_list()]
    assert sum(counts) == len(sp) >= 1
    assert red.n_rows == reduced.n_rows
    assert np.array_equal(red.seg_ptr, reduced.seg_ptr)
    assert np.array_equal(red.col, reduced.col)
    assert np.array_equal(red.coef, reduced.coef)


@pytest.mark.parametrize("name", ["secp256k1+bmmp+blt", "tornado/withdraw+pedersen"])
def test_small_trusted_configurations_match_the_oracle(name):
    cfg = CONFIGS[name]
    main = api.readR1CS(fixtures.path(cfg["main"]))
    subs = [(cfg["trusted_names"][i], api.readR1CS(fixtures.path(t))) for i, t in enumerate(cfg["trusted"])]
    subs.sort(key=lambda x: -len(x[1]))
    check_against_oracle(main, subs)


def test_quirks_of_the_pass():
    # the match that starts inside a consumed window stalls the walk (:370)
    sub = mk([mul(2, 2, 3), mul(3, 2, 1)], n_wires=4)
    main = mk([mul(2, 2, 3), mul(3, 2, 4), mul(4, 2, 5), mul(5, 2, 1), mul(1, 1, 6)], n_wires=7)
    sp, red = check_against_oracle(main, [("cube", sub)])
    assert sp == [("cube", [3], [5])] and red.n_rows == 3
    # non-overlapping matches are all taken
    sub = mk([mul(2, 2, 3), mul(3, 3, 1)], n_wires=4)
    main = mk([mul(2, 2, 3), mul(3, 3, 4), mul(1, 2, 7), mul(4, 4, 5), mul(5, 5, 6)], n_wires=8)
    sp, red = check_against_oracle(main, [("pow4", sub)])
    assert [g[1:] for g in sp] == [([3], [5]), ([5], [7])] and red.n_rows == 1
    # the last row is not hashed (:261) but verified (:301-312)
    sub = mk([mul(2, 2, 3), ([(3, 1)], [(2, 1)], [(1, 1)])], n_wires=4)
    main = mk([mul(2, 2, 3), ([(3, 1)], [(2, 1)], [(4, 5)]), mul(2, 2, 5), ([(5, 1)], [(2, 1)], [(1, 1)])], n_wires=6)
    sp, red = check_against_oracle(main, [("cube", sub)])
    assert len(sp) == 1 and red.n_rows == 2 and sp[0][1:] == ([3], [2])
    # wires with identical signatures are paired by wire id
    sub = mk([([(2, 1), (3, 1)], [(0, 1)], [(1, 1)])], n_wires=4, pub_in=2)
    main = mk([mul(4, 4, 5), ([(5, 1), (4, 1)], [(0, 1)], [(6, 1)]), mul(6, 6, 1)], n_wires=7)
    sp, red = check_against_oracle(main, [("add", sub)])
    assert sp == [("add", [5, 6], [7])] and red.n_rows == 2
    # explicit zeros do not count (:213, :231, :284)
    sub = mk([mul(2, 2, 3), mul(3, 2, 1)], n_wires=4)
    main = mk([([(2, 1), (4, 0)], [(2, 1)], [(3, 1)]), ([(3, 1)], [(2, 1), (0, 0)], [(1, 1)])], n_wires=5)
    sp, red = check_against_oracle(main, [("cube", sub)])
    assert sp == [("cube", [3], [2])] and red.n_rows == 0
    # KeyError of :381 when a trusted input never appears
    sub = mk([mul(2, 2, 1)], n_wires=4, pub_in=2)
    main = mk([mul(2, 2, 3), mul(3, 3, 1)], n_wires=4)
    with pytest.raises(KeyError):
        device_chain(main, [("sq", sub)])


def test_random_planted_circuits_on_the_device():
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
                device_chain(main, [("f", sub)])
            continue
        red, sp, _ = device_chain(main, [("f", sub)])
        assert sp == [(n, list(i), list(o)) for n, i, o in want_sp], case
        assert rows_of(red) == want_rows, case
        n_matches += len(sp)
    assert n_matches > 100 and n_keyerr > 0


@pytest.mark.parametrize("name", ["secp256k1+bmmp+blt", "tornado/withdraw+pedersen", "ecdsa+secp256k1"])
def test_solve_on_the_device_abstracted_system_matches_the_goldens(name):
    """file -> device abstraction -> classification in place -> solve: verdict, counts and the `unique` bitmap of the
    oracle goldens (which were minted on the host-abstracted system)."""
    cfg = CONFIGS[name]
    ok, res, sizes = api.solve_with_device_abstraction(
        fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg["trusted"]], cfg["trusted_names"],
        secp_solve=cfg.get("secp_solve", False))
    g = GOLD[name]
    assert ok == g["verdict"]
    assert hashlib.sha256(res.unique_bytes()).hexdigest() == g["sha_unique"]
    assert (res.c.n_unique_nontrivial, res.c.n_nontrivial, res.c.n_targets_unique, res.c.n_unique) == \
        (g["uniq"], g["nontriv"], g["tgt"], g["n_unique"])
    assert sizes[0] == g["reduced_rows"] and sizes[2] == g["n_specials"]


def test_prepared_trusted_circuit_gives_the_same_abstraction():
    """ecne_abstract_prepare + ecne_abstract_apply_prepared (the trusted circuit prepared ahead, on another thread while the
    main circuit is read: api.read_and_prepare) against the plain apply."""
    from configs import CONFIGS
    from ecneproject_b200 import fixtures
    for name in ("secp256k1+bmmp+blt", "tornado/withdraw+pedersen"):
        cfg = CONFIGS[name]
        paths = [fixtures.path(t) for t in cfg["trusted"]]
        main, prepared = api.read_and_prepare(fixtures.path(cfg["main"]), paths, cfg["trusted_names"])
        plain = sorted([(cfg["trusted_names"][i], api.readR1CS(p)) for i, p in enumerate(paths)], key=lambda x: -len(x[1]))
        out = []
        for subs in (prepared, plain):
            da = api.DeviceAbstraction(main)
            try:
                for nm, sub in subs:
                    da.apply(nm, sub)
                (seg, col, coef), sp = da.export()
                out.append((seg.tobytes(), col.tobytes(), coef.tobytes(), sp))
            finally:
                da.free()
        assert out[0] == out[1]
        # a prepared circuit handed over with another circuit's rows is refused
        other = api.readR1CS(fixtures.path(cfg["main"]))
        da = api.DeviceAbstraction(main)
        try:
            wrong = prepared[0][1]
            ph = api.ProblemHandle(other, None, other.known, other.targets, other.n_vars)
            n = C.c_uint64(0)
            st = api._engine().ecne_abstract_apply_prepared(da.handle, 0, C.byref(ph.c), wrong.handle, C.byref(n))
            assert st == _abi.ECNE_E_BADARG
        finally:
            da.free()
