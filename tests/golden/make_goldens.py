#!/usr/bin/env python3
"""Mint tests/golden/*.json.

1. tests/golden/survey_appendix_b.json — the per-fixture table of SURVEY.md Appendix B.1 (verdict,
   uniq/nontriv, targets, #U, sha12 of the packed `unique` bitmap, outer rounds, pops), produced
   by the survey's independent Python restatement of the Julia.  Parsed from SURVEY.md verbatim.
2. tests/golden/oracle_goldens.json — what oracle/ecne_oracle.cpp produces on every run config
   (circomlib corpus + root / tornado / trusted-function configs): verdict, counts, SHA-256 of the
   unique and known bitmaps, per-rule firing counters.  The GPU engine is diffed against these.

Run here (CPU):  python tests/golden/make_goldens.py [--with-ecdsa]
(lives under tests/: it executes the oracle, which only test infrastructure may do)
"""
import hashlib
import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ecneproject_b200 import api, fixtures  # noqa: E402
import oracle_lib  # noqa: E402
from configs import CONFIGS  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def parse_survey():
    rows = {}
    pat = re.compile(r"^\| `([^`]+)` \| (\d+) \| (\d+) \| ([TF]) \| (\d+)/(\d+) \| (\d+)/(\d+) \| (\d+) \| `([0-9a-f]{12})` \| (\d+) \| (\d+) \|")
    for line in open(os.path.join(ROOT, "SURVEY.md")):
        m = pat.match(line)
        if m:
            g = m.groups()
            rows[g[0]] = {"rows": int(g[1]), "n_vars": int(g[2]), "verdict": g[3] == "T",
                          "uniq": int(g[4]), "nontriv": int(g[5]), "tgt": int(g[6]),
                          "ntgt": int(g[7]), "n_unique": int(g[8]), "sha12": g[9],
                          "rounds": int(g[10]), "pops": int(g[11])}
    return rows


def run_config(name, cfg):
    t0 = time.time()
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]),
                                          [fixtures.path(t) for t in cfg.get("trusted", [])],
                                          cfg.get("trusted_names", []))
    t1 = time.time()
    rec = {"rows": main.n_rows, "reduced_rows": reduced.n_rows, "n_vars": main.n_vars,
           "nnz": reduced.nnz, "n_specials": len(specials)}
    try:
        res = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars,
                               cfg.get("secp_solve", False))
    except oracle_lib.OracleError as e:
        rec["status"] = e.status
        return rec
    t2 = time.time()
    c = res.c
    rec.update({
        "status": 0, "verdict": bool(c.verdict), "uniq": c.n_unique_nontrivial,
        "nontriv": c.n_nontrivial, "tgt": c.n_targets_unique, "ntgt": len(main.targets),
        "n_unique": c.n_unique, "sha_unique": hashlib.sha256(res.unique_bytes()).hexdigest(),
        "sha_known": hashlib.sha256(res.known_bytes()).hexdigest(),
        "rounds": c.outer_rounds, "pops": int(res.oracle_counters[0]),
        "sweep_visits": int(res.oracle_counters[1]),
        "fired": [int(x) for x in res.oracle_counters[3:19]],
        "prep_s": round(t1 - t0, 3), "solve_s": round(t2 - t1, 3),
    })
    return rec


def main():
    os.makedirs(GOLD, exist_ok=True)
    survey = parse_survey()
    with open(os.path.join(GOLD, "survey_appendix_b.json"), "w") as f:
        json.dump(survey, f, indent=1, sort_keys=True)
    print("survey rows:", len(survey))
    with_ecdsa = "--with-ecdsa" in sys.argv
    out = {}
    path = os.path.join(GOLD, "oracle_goldens.json")
    if os.path.exists(path):
        out = json.load(open(path))
    for name, cfg in CONFIGS.items():
        if cfg.get("big") and not with_ecdsa:
            continue
        rec = run_config(name, cfg)
        out[name] = rec
        print(name, {k: rec.get(k) for k in ("status", "verdict", "uniq", "nontriv", "tgt", "rounds", "pops", "solve_s")}, flush=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    # compare with the survey table
    bad = 0
    for k, s in survey.items():
        name = "circomlib/" + k
        if name not in out:
            print("missing", name)
            bad += 1
            continue
        o = out[name]
        for fld in ("verdict", "uniq", "nontriv", "tgt", "n_unique", "rounds", "pops"):
            if o.get(fld) != s[fld]:
                print("MISMATCH", name, fld, o.get(fld), s[fld])
                bad += 1
        if o.get("sha_unique", "")[:12] != s["sha12"]:
            print("MISMATCH", name, "sha", o.get("sha_unique", "")[:12], s["sha12"])
            bad += 1
    print("survey mismatches:", bad)


if __name__ == "__main__":
    main()
