#!/usr/bin/env python3
"""Upload + classify (ecne_upload) without solving: the set-up kernels alone, for compute-sanitizer's racecheck — the
persistent solve kernel's spin barriers do not finish under that tool on the full grid.  Synthetic rows that reach every
shared-memory stage of the set-up: long linear rows whose weights are not in order (both length classes of the sorting
kernel, with dropped zero terms), a long bit decomposition beyond 2^253 (block-per-row classifier), thousands of
distinct and of repeated constants (sample sort forced by ECNE_SAMPLE_SORT=1); then a fixture with trusted circuits.
Usage: setup_only.py [config ...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ECNE_SAMPLE_SORT", "1")
import numpy as np
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
from helpers import MiniR1CS, P

lib = api._engine()
def upload(problem_like, specials, known, targets, n_vars, secp=False, compact=False):
    ph = api.ProblemHandle(problem_like, specials, known, targets, n_vars, secp, compact=compact)
    h = C.c_void_p()
    st = lib.ecne_upload(C.byref(ph.c), C.byref(h))
    assert st == 0, lib.ecne_last_error()
    lib.ecne_free_resident(h)

rng = np.random.default_rng(11)
for nbits in (100, 600, 1100):
    bits = list(range(3, 3 + nbits))
    out, extra, dead = 2, 3 + nbits, 4 + nbits
    perm = rng.permutation(nbits)
    rows = [({b: 1}, {b: 1, 1: -1}, {}) for b in bits]
    c1 = {out: 1, **{b: -pow(2, int(perm[i]), P) for i, b in enumerate(bits)}}
    c2 = {extra: 1, **{b: int(3 + perm[i] % 7) for i, b in enumerate(bits)}, dead: 0}
    c3 = {out: 1, **{b: -pow(2, i, P) for i, b in enumerate(bits)}}          # in order: the shortcut
    for b in bits[::5]:
        c2[b] = 0
    rows += [({}, {}, c1), ({}, {}, c2), ({}, {}, c3)]
    for i in range(3000):
        v = int(rng.integers(1, 2**62)) * int(rng.integers(1, 2**62)) if i % 2 else 12345
        rows.append(({}, {}, {5 + nbits + i: 1, 1: -v % P}))
    m = MiniR1CS(rows, n_vars=5 + nbits + 3000, known=[1, out], targets=[extra])
    for compact in (False, True):
        upload(m, [], m.known, m.targets, m.n_vars, compact=compact)
    print(f"synthetic rows with {nbits}-term sums: uploaded and classified (full and compact form)", flush=True)
for name in sys.argv[1:]:
    cfg = CONFIGS[name]
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
    upload(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False), compact=True)
    print(f"{name}: uploaded and classified", flush=True)
