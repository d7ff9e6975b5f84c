#!/usr/bin/env python3
"""A/B of two builds of the host reader on one box: read ecdsa.r1cs N times with each."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ecneproject_b200 import fixtures
path = fixtures.path("ecdsa.r1cs").encode()
libs = {n: C.CDLL(os.path.join(ROOT, "ecneproject_b200", n)) for n in sys.argv[1:]}
for rep in range(6):
    for n, lib in libs.items():
        for flags in (0, 1):
            if not hasattr(lib, "ecne_read_r1cs_opts") and flags: continue
            out = C.c_void_p()
            t0 = time.perf_counter()
            if hasattr(lib, "ecne_read_r1cs_opts"):
                lib.ecne_read_r1cs_opts.argtypes = [C.c_char_p, C.c_uint, C.POINTER(C.c_void_p)]
                st = lib.ecne_read_r1cs_opts(path, flags, C.byref(out))
            else:
                lib.ecne_read_r1cs.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
                st = lib.ecne_read_r1cs(path, C.byref(out))
            dt = time.perf_counter() - t0
            lib.ecne_r1cs_free.argtypes = [C.c_void_p]
            lib.ecne_r1cs_free(out)
            print(f"rep{rep} {n} flags={flags} st={st} {dt*1e3:.1f} ms", flush=True)
