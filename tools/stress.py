#!/usr/bin/env python3
"""Repeat resident solves of a few configs many times and compare every result with the first one and with
the committed golden (flushes out races in the solve kernel's barriers / solo modes).  Usage: stress.py [reps]"""
import ctypes as C, hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_goldens.json")))
lib = api._engine()
bad = 0
for name in ["ecdsa+secp256k1", "ecdsa", "secp256k1+bmmp+blt", "tornado/withdraw+pedersen", "root/poseidon",
             "circomlib/EdDSAPoseidonVerifier@eddsaposeidon", "root/bigmult86_3", "tornado/merkleTree"]:
    cfg = CONFIGS[name]
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])],
                                          cfg.get("trusted_names", []))
    ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
    h = C.c_void_p()
    assert lib.ecne_upload(C.byref(ph.c), C.byref(h)) == 0, lib.ecne_last_error()
    first = None
    for i in range(reps):
        res = api.SolveResult(main.n_vars)
        st = lib.ecne_solve_resident(h, C.byref(res.c))
        cur = (st, hashlib.sha256(res.unique_bytes()).hexdigest(), hashlib.sha256(res.known_bytes()).hexdigest(),
               bool(res.c.verdict), int(res.c.n_unique), int(res.c.outer_rounds), int(res.c.inner_rounds))
        if first is None:
            first = cur
            ok = st == 0 and cur[1] == gold[name]["sha_unique"] and cur[3] == gold[name]["verdict"]
            if not ok:
                bad += 1
                print("MISMATCH vs golden", name, cur)
        elif cur != first:
            bad += 1
            print("NONDETERMINISTIC", name, i, cur, first)
    lib.ecne_free_resident(h)
    print(f"{name}: {reps} solves, {'stable' if bad == 0 else 'PROBLEMS'}: outer={first[5]} inner={first[6]}", flush=True)
print("stress:", "OK" if bad == 0 else f"{bad} problems")
sys.exit(1 if bad else 0)
