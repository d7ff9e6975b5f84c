#!/usr/bin/env python3
"""Solve the K-times block-diagonal tiling of a config (SURVEY.md §8d "S-K") resident on the GPU and print the
timings and the dense-round figures.  Usage: run_tiled.py <K> [reps] [config] [key=value engine knobs ...]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
import bench
K = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
name = sys.argv[3] if len(sys.argv) > 3 else "ecdsa+secp256k1"
cfg = CONFIGS[name]
reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
lib = api._engine()
for kv in sys.argv[4:]:
    k, v = kv.split('=')
    assert lib.ecne_set_option(k.encode(), int(v)) == 0, lib.ecne_last_error()
if K > 1:
    t, sp, known, targets, nv = bench.tile_problem(np, reduced, specials, main, K)
else:
    t, sp, known, targets, nv = reduced, specials, main.known, main.targets, main.n_vars
ph = api.ProblemHandle(t, sp, known, targets, nv, cfg.get("secp_solve", False))
h = C.c_void_p()
assert lib.ecne_upload(C.byref(ph.c), C.byref(h)) == 0, lib.ecne_last_error()
nnz_nz = int(np.count_nonzero(reduced.coef.any(axis=1)))
b_eval = 32.0 + nnz_nz / float(reduced.n_rows)
for i in range(reps):
    res = api.SolveResult(nv)
    t0 = time.perf_counter()
    st = lib.ecne_solve_resident(h, C.byref(res.c))
    dt = time.perf_counter() - t0
    c = res.c
    dms = c.dense_cycles / 1.965e6
    print(f"tile{K} {name} rep{i} st={st} rows={t.n_rows} wall={dt*1e3:.3f}ms solve={c.ms_solve:.3f} sweep={c.ms_sweep:.3f} outer={c.outer_rounds} "
          f"inner={c.inner_rounds} evals={c.constraint_evals} n_unique={c.n_unique} verdict={c.verdict} | dense: rounds={c.dense_rounds} "
          f"evals={c.dense_evals} ms={dms:.3f} GB/s={b_eval*c.dense_evals/max(dms,1e-9)/1e6:.1f} | whole kernel GB/s={b_eval*c.constraint_evals/max(c.ms_sweep,1e-9)/1e6:.1f}",
          flush=True)
lib.ecne_free_resident(h)
