#!/usr/bin/env python3
"""Time the report path of one config on the GPU: resident solve, then ecne_report_resident (two passes: bitmap +
counts, then the compacted state), against the full per-wire export.  Usage: report_one.py <config>"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
name = sys.argv[1]
cfg = CONFIGS[name]
reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
lib = api._engine()
ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
h = C.c_void_p()
assert lib.ecne_upload(C.byref(ph.c), C.byref(h)) == 0, lib.ecne_last_error()
for rep in range(3):
    res = api.SolveResult(main.n_vars)
    t0 = time.perf_counter(); assert lib.ecne_solve_resident(h, C.byref(res.c)) == 0; t1 = time.perf_counter()
    bad = api.BadConstraints(h, reduced.n_rows); t2 = time.perf_counter()
    full = api.SolveResult(main.n_vars, full_state=True)
    t3 = time.perf_counter(); assert lib.ecne_solve_resident(h, C.byref(full.c)) == 0; t4 = time.perf_counter()
    d2h_compact = reduced.n_rows // 8 + 138 * len(bad.wire)
    d2h_full = 133 * main.n_vars
    print(f"{name} rep{rep}: solve {1e3*(t1-t0):.2f} ms | report {1e3*(t2-t1):.2f} ms ({bad.n_bad_rows} rows, {len(bad.wire)} wires, "
          f"{d2h_compact/1e6:.2f} MB D2H) | solve + full per-wire export {1e3*(t4-t3):.2f} ms ({d2h_full/1e6:.1f} MB D2H)", flush=True)
lib.ecne_free_resident(h)
