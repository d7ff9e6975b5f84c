#!/usr/bin/env python3
"""Solve one config N times on the GPU (resident) and print timings.  Usage: run_one.py <config> [reps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = CONFIGS[name]
reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
lib = api._engine()
for kv in sys.argv[3:]:  # engine knobs: key=value (ecne_set_option)
    k, v = kv.split('=')
    assert lib.ecne_set_option(k.encode(), int(v)) == 0, lib.ecne_last_error()
ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
h = C.c_void_p()
assert lib.ecne_upload(C.byref(ph.c), C.byref(h)) == 0, lib.ecne_last_error()
for i in range(reps):
    res = api.SolveResult(main.n_vars)
    t0 = time.perf_counter()
    st = lib.ecne_solve_resident(h, C.byref(res.c))
    dt = time.perf_counter() - t0
    c = res.c
    print(f"{name} rep{i} st={st} wall={dt*1e3:.3f}ms solve={c.ms_solve:.3f} sweep={c.ms_sweep:.3f} outer={c.outer_rounds} inner={c.inner_rounds} "
          f"us/round={1e3*c.ms_sweep/max(1,c.inner_rounds):.2f} evals={c.constraint_evals} rule_evals={c.rule_evals} launches={c.sweep_launches}", flush=True)
lib.ecne_free_resident(h)
