#!/usr/bin/env python3
"""Mint tests/golden/report_goldens.json: for the two full-size run configurations, what the ORACLE's final state
implies for the "Bad Constraints" listing (R1CSConstraintSolver.jl:1609-1633): number of listed rows and wires,
SHA-256 of the packed row bitmap (bit i&7 of byte i>>3 = 0-based row i) and of the ascending uint32 wire list.
Runs the oracle (13 s of CPU); tests/ only."""
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
import oracle_lib  # noqa: E402
from test_report import _expected  # noqa: E402


def digest(n_rows, rows, wires):
    bits = np.zeros(n_rows, dtype=np.uint8)
    bits[rows] = 1
    return {"n_bad_rows": int(len(rows)), "n_wires": int(len(wires)),
            "sha_rows": hashlib.sha256(np.packbits(bits, bitorder="little").tobytes()).hexdigest(),
            "sha_wires": hashlib.sha256(np.ascontiguousarray(wires, dtype=np.uint32).tobytes()).hexdigest()}


if __name__ == "__main__":
    out = {}
    for name in ("ecdsa+secp256k1", "ecdsa"):
        cfg = CONFIGS[name]
        reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])],
                                              cfg.get("trusted_names", []))
        o = oracle_lib.solve(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
        rows, wires, _ = _expected(reduced, o.unique_bits, main.n_vars)
        out[name] = digest(reduced.n_rows, rows, wires)
        print(name, out[name], flush=True)
    json.dump(out, open(os.path.join(HERE, "report_goldens.json"), "w"), indent=1, sort_keys=True)
