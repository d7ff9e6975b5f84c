#!/usr/bin/env python3
"""Time ecne_solve (host buffers in/out) on one config.  Usage: e2e_one.py <config> [reps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = CONFIGS[name]
reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
lib = api._engine()
ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
for i in range(reps):
    res = api.SolveResult(main.n_vars)
    t0 = time.perf_counter()
    st = lib.ecne_solve(C.byref(ph.c), C.byref(res.c))
    dt = time.perf_counter() - t0
    c = res.c
    print(f"{name} rep{i} st={st} wall={dt*1e3:.2f}ms h2d={c.ms_h2d:.2f} classify={c.ms_classify:.2f} solve={c.ms_solve:.2f} d2h={c.ms_d2h:.2f}", flush=True)
