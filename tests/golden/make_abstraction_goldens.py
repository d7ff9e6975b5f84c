#!/usr/bin/env python3
"""Mint tests/golden/abstraction_goldens.json: per trusted-function run configuration, the size of the reduced
system, the number of special constraints and a SHA-256 over the reduced arrays + the (name, inputs, outputs)
list that ecne_abstraction (libecne_host.so) produces.

These are REGRESSION pins of the host fast path (minted by it; its output is identical to that of
oracle/abstraction_ref.py, the plain-Python restatement, on every configuration: tests/test_abstraction_parity.py),
not reference outputs — Julia cannot run here.  What ties them to the reference: the match counts bench/bench_abstraction.jl:16,24
asserts, the reduced sizes of SURVEY.md §8a, and the verdicts the reference asserts for the abstracted
configurations (test/runtests.jl:25,30,35, examples/ecdsa_secp_abstraction.jl:4), which the oracle and the
engine reproduce from exactly these special constraints."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from configs import CONFIGS  # noqa: E402
from ecneproject_b200 import api, fixtures  # noqa: E402


def digest(red, sp):
    h = hashlib.sha256()
    for a in (red.seg_ptr, red.col, red.coef, red.known, red.targets):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(json.dumps([(n, [int(x) for x in i], [int(x) for x in o]) for n, i, o in sp.as_list()]).encode())
    return h.hexdigest()


def entry(red, sp):
    return {"rows": int(red.n_rows), "nnz": int(red.nnz), "specials": len(sp), "sha256": digest(red, sp)}


def mint():
    out = {}
    for name, cfg in CONFIGS.items():
        if not cfg.get("trusted"):
            continue
        red, sp, _ = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg["trusted"]],
                                 cfg["trusted_names"])
        out[name] = entry(red, sp)
    main = api.readR1CS(fixtures.path("bigmultmodp86_3.r1cs"))
    sub = api.readR1CS(fixtures.path("bigmultshortlong86_3.r1cs"))
    sp, red = api.abstraction("bigmultmodp", main, sub)
    out["bench/bigmultmodp86_3<-bigmultshortlong86_3"] = entry(red, sp)
    return out


if __name__ == "__main__":
    json.dump(mint(), open(os.path.join(HERE, "abstraction_goldens.json"), "w"), indent=1, sort_keys=True)
