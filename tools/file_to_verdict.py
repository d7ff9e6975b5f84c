#!/usr/bin/env python3
"""file -> verdict with abstraction() on the device: stage times of every step.  Usage: file_to_verdict.py [config] [reps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
name = sys.argv[1] if len(sys.argv) > 1 else "ecdsa+secp256k1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = CONFIGS[name]
api._engine()
main = subs = da = res = None
for rep in range(reps):
    main = subs = da = res = None   # the circuits of the run before (hundreds of megabytes) are released outside the timed region
    import gc; gc.collect()
    t0 = time.perf_counter()
    # the main circuit is parsed on a worker thread while this one reads and prepares the trusted circuits
    main, subs = api.read_and_prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])],
                                      cfg.get("trusted_names", []))
    t1 = time.perf_counter()
    da = api.DeviceAbstraction(main)
    t2 = time.perf_counter()
    for nm, sub in subs:
        da.apply(nm, sub)
    t3 = time.perf_counter()
    h = da.upload(cfg.get("secp_solve", False))
    t4 = time.perf_counter()
    res = api.SolveResult(main.n_vars)
    st = api._engine().ecne_solve_resident(h, C.byref(res.c))
    t5 = time.perf_counter()
    api._engine().ecne_free_resident(h); da.free()
    print(f"{name} rep{rep}: read {1e3*(t1-t0):.1f} ms | H2D of the unreduced system {1e3*(t2-t1):.1f} | abstraction on the device {1e3*(t3-t2):.1f} | "
          f"classification in place {1e3*(t4-t3):.1f} | solve {1e3*(t5-t4):.1f} | verdict {bool(res.c.verdict)} st {st} | file->verdict {1e3*(t5-t0):.1f} ms", flush=True)
