#!/usr/bin/env python3
"""Time ecne_solve with pinned host buffers in the compact form on one config.  Usage: e2e_compact.py <config> [reps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from ecneproject_b200 import api, fixtures, _abi
from configs import CONFIGS
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = CONFIGS[name]
reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
lib = api._engine()
ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False), compact=True)
pins = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); pins.append(t); return t.data_ptr()
cls = next(a for a in ph.keep if a.dtype == np.uint8 and a.size == reduced.nnz)
ph.c.col = C.cast(pin(reduced.col.view(np.int32)), _abi.u32p)
ph.c.seg_ptr32 = C.cast(pin(np.asarray(reduced.seg_ptr).astype(np.uint32).view(np.int32)), _abi.u32p)
ph.c.coef_class = C.cast(pin(cls), _abi.u8p)
other = next(a for a in ph.keep if a.dtype == np.uint64 and a.size == 4 * ph.c.n_coef_other)
term = next(a for a in ph.keep if a.dtype == np.uint32 and a.size == ph.c.n_coef_other and a is not cls)
ph.c.coef_other = C.cast(pin(other.view(np.int64)), _abi.u64p)
ph.c.coef_other_term = C.cast(pin(term.view(np.int32)), _abi.u32p)
for i in range(reps):
    res = api.SolveResult(main.n_vars, full_state=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = lib.ecne_solve(C.byref(ph.c), C.byref(res.c))
    dt = time.perf_counter() - t0
    c = res.c
    print(f"{name} rep{i} st={st} wall={dt*1e3:.3f}ms h2d={c.ms_h2d:.3f} classify={c.ms_classify:.3f} solve={c.ms_solve:.3f} device={c.ms_device:.3f} d2h={c.ms_d2h:.3f} total={c.ms_total:.3f}", flush=True)
