"""The oracle is pinned before it is trusted (prompt ③):
  * every Boolean the reference itself asserts (test/runtests.jl, examples/*.jl, README.md:106);
  * the survey's independent Python restatement (SURVEY.md Appendix B.1 -> golden/survey_appendix_b.json);
  * the committed oracle goldens (golden/oracle_goldens.json) that the GPU parity tests diff against.
"""
import hashlib
import json
import os

import pytest

from configs import CONFIGS
from ecneproject_b200 import api, fixtures
import oracle_lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SURVEY = json.load(open(os.path.join(GOLD, "survey_appendix_b.json")))
ORACLE = json.load(open(os.path.join(GOLD, "oracle_goldens.json")))

SMALL = [n for n, c in CONFIGS.items() if not c.get("big")]


def run_oracle(name):
    cfg = CONFIGS[name]
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]),
                                          [fixtures.path(t) for t in cfg.get("trusted", [])],
                                          cfg.get("trusted_names", []))
    return oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars,
                            cfg.get("secp_solve", False)), reduced, main


@pytest.mark.parametrize("name", SMALL)
def test_oracle_matches_goldens_and_pins(name):
    res, reduced, main = run_oracle(name)
    g = ORACLE[name]
    assert g["status"] == 0
    assert bool(res.c.verdict) == g["verdict"]
    assert hashlib.sha256(res.unique_bytes()).hexdigest() == g["sha_unique"]
    assert hashlib.sha256(res.known_bytes()).hexdigest() == g["sha_known"]
    assert (res.c.n_unique_nontrivial, res.c.n_nontrivial, res.c.n_targets_unique) == (g["uniq"], g["nontriv"], g["tgt"])
    assert res.c.outer_rounds == g["rounds"] and int(res.oracle_counters[0]) == g["pops"]
    assert reduced.n_rows == g["reduced_rows"]
    pinned = CONFIGS[name].get("pinned")
    if pinned is not None:  # the reference's own assertion
        assert bool(res.c.verdict) == pinned[0], f"reference pins {pinned}"
    if name.startswith("circomlib/"):
        s = SURVEY[name[len("circomlib/"):]]
        assert bool(res.c.verdict) == s["verdict"]
        assert (res.c.n_unique_nontrivial, res.c.n_nontrivial) == (s["uniq"], s["nontriv"])
        assert (res.c.n_targets_unique, len(main.targets)) == (s["tgt"], s["ntgt"])
        assert res.c.n_unique == s["n_unique"]
        assert hashlib.sha256(res.unique_bytes()).hexdigest()[:12] == s["sha12"]
        assert (res.c.outer_rounds, int(res.oracle_counters[0])) == (s["rounds"], s["pops"])


def test_big_goldens_present_and_pinned():
    # ecdsa takes ~10 s of oracle time + 20 s prep per config: minted by tests/golden/make_goldens.py --with-ecdsa
    g = ORACLE["ecdsa+secp256k1"]
    assert g["verdict"] is True  # examples/ecdsa_secp_abstraction.jl:4
    assert (g["uniq"], g["nontriv"], g["tgt"], g["rounds"], g["pops"]) == (694285, 694311, 6, 28, 1602505)
    assert g["reduced_rows"] == 694264 and g["n_specials"] == 25
    g = ORACLE["ecdsa"]
    assert g["verdict"] is False and (g["uniq"], g["nontriv"], g["rounds"], g["pops"]) == (700612, 1089136, 4, 2747850)


def test_secp_without_secp_solve_throws():
    # UndefVarError(:dsu) at :762 when BigMultModP + BigLessThan specials exist and secp_solve=false
    cfg = CONFIGS["secp256k1+bmmp+blt"]
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg["trusted"]],
                                          cfg["trusted_names"])
    with pytest.raises(oracle_lib.OracleError) as e:
        oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, False)
    assert e.value.status == -4


def test_oracle_disjoint_sets_of_equal_wires():
    """The oracle's restatement of :634-678 and of its one observable use (:760-768): same sets + a BigLessThan without
    outputs is a BoundsError; building the sets raises BoundsErrors of its own on rows with stored zeros."""
    from helpers import MiniR1CS, dsu_system
    import oracle_lib

    def status(m, sp, secp):
        try:
            oracle_lib.solve(m, sp, m.known, m.targets, m.n_vars, secp)
            return 0
        except oracle_lib.OracleError as e:
            return e.status

    for link in ("xy", "chain", "const"):
        assert status(*dsu_system(link, True), True) == -3
        assert status(*dsu_system(link, False), True) == 0
    assert status(*dsu_system("none", True), True) == 0
    assert status(*dsu_system("xy", True), False) == -4      # UndefVarError(:dsu) without secp_solve (:762)
    base = [({2: 1}, {3: 1}, {4: 1})]
    for row, want in ((({}, {}, {1: 4, 6: 0}), -3), (({}, {}, {5: 3, 6: 0}), 0), (({}, {}, {1: 4, 6: 2}), 0)):
        m = MiniR1CS(base + [row], n_vars=8, known=[1, 2, 3], targets=[4])
        assert status(m, [], True) == want
        assert status(m, [], False) == 0


def test_odd_permutation_sum_both_ways():
    """slow_det (:1389-1400) sums the ODD permutations only.  The oracle enumerates them for k <= 8 (as the reference does)
    and uses a subset recurrence above; both against a plain-Python evaluation of the definition, and — char != 2 —
    against (permanent - determinant) / 2."""
    import itertools
    import random
    import oracle_lib
    from helpers import P
    rng = random.Random(11)

    def by_definition(m):
        k, tot = len(m), 0
        for perm in itertools.permutations(range(k)):
            inv = sum(1 for x in range(k) for y in range(x + 1, k) if perm[x] > perm[y])
            if inv & 1:
                t = 1
                for j in range(k):
                    t = t * m[j][perm[j]] % P
                tot = (tot + t) % P
        return tot

    for k in (1, 2, 3, 4, 5, 6, 7):
        for kind in ("random", "small", "rank1"):
            if kind == "random":
                m = [[rng.randrange(P) for _ in range(k)] for _ in range(k)]
            elif kind == "small":
                m = [[rng.randrange(-3, 4) % P for _ in range(k)] for _ in range(k)]
            else:
                u, v = [rng.randrange(1, P) for _ in range(k)], [rng.randrange(1, P) for _ in range(k)]
                m = [[u[i] * v[j] % P for j in range(k)] for i in range(k)]
            e, d = oracle_lib.odd_permutation_sum(m)
            assert e == d == by_definition(m), (k, kind)
    m = [[rng.randrange(P) for _ in range(9)] for _ in range(9)]
    e, d = oracle_lib.odd_permutation_sum(m)
    assert e == d
    # a rank-1 matrix is singular, yet its odd-permutation sum is (k!/2) * prod(u) * prod(v) != 0: the reference's
    # rule fires on it
    k = 12
    u, v = [rng.randrange(1, P) for _ in range(k)], [rng.randrange(1, P) for _ in range(k)]
    m = [[u[i] * v[j] % P for j in range(k)] for i in range(k)]
    _, d = oracle_lib.odd_permutation_sum(m, enumerate_too=False)
    want = 1
    for x in u + v:
        want = want * x % P
    import math
    assert d == want * (math.factorial(k) // 2) % P
