#!/usr/bin/env python3
"""Run every config through the CUDA engine and diff against tests/golden/oracle_goldens.json.
Usage (on a GPU box): python tools/gpu_parity.py [--with-ecdsa] [name-substring ...]"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ctypes as C  # noqa: E402
from ecneproject_b200 import api, fixtures, _abi  # noqa: E402
from configs import CONFIGS  # noqa: E402


def run(name, cfg, gold):
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]),
                                          [fixtures.path(t) for t in cfg.get("trusted", [])],
                                          cfg.get("trusted_names", []))
    lib = api._engine()
    ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
    res = api.SolveResult(main.n_vars, full_state=False)
    t0 = time.time()
    st = lib.ecne_solve(C.byref(ph.c), C.byref(res.c))
    dt = time.time() - t0
    if st != 0:
        return False, f"status {st}: {lib.ecne_last_error().decode()}"
    g = gold[name]
    su = hashlib.sha256(res.unique_bytes()).hexdigest()
    sk = hashlib.sha256(res.known_bytes()).hexdigest()
    c = res.c
    ok = (su == g["sha_unique"] and bool(c.verdict) == g["verdict"] and c.n_unique_nontrivial == g["uniq"]
          and c.n_nontrivial == g["nontriv"] and c.n_targets_unique == g["tgt"])
    okk = sk == g["sha_known"]
    msg = (f"verdict={bool(c.verdict)} uniq={c.n_unique_nontrivial}/{c.n_nontrivial} (gold {g['uniq']}/{g['nontriv']}) "
           f"tgt={c.n_targets_unique} outer={c.outer_rounds} (gold {g['rounds']}) inner={c.inner_rounds} "
           f"evals={c.constraint_evals} solve={c.ms_solve:.3f}ms sweep={c.ms_sweep:.3f}ms h2d={c.ms_h2d:.2f} "
           f"classify={c.ms_classify:.2f} total={dt*1e3:.1f}ms known_ok={okk}")
    return ok and okk, msg


def main():
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_goldens.json")))
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    with_big = "--with-ecdsa" in sys.argv
    bad = 0
    n = 0
    for name, cfg in CONFIGS.items():
        if cfg.get("big") and not with_big:
            continue
        if args and not any(a in name for a in args):
            continue
        if gold[name].get("status", 0) != 0:
            continue
        ok, msg = run(name, cfg, gold)
        n += 1
        if not ok:
            bad += 1
        print(("OK   " if ok else "FAIL ") + name + " :: " + msg, flush=True)
    print(f"{n - bad}/{n} configs bit-identical to the oracle")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
