#!/usr/bin/env python3
"""Diagnose a parity failure: run engine + oracle with full state and print differing wires.
(Lives under tests/: it executes the oracle, which only test infrastructure may do.)  Usage: python tests/diag_gpu_diff.py <config> [limit]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import numpy as np
from ecneproject_b200 import api, fixtures
from configs import CONFIGS
import oracle_lib

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
def toint(a): return sum(int(a[i]) << (64 * i) for i in range(4))
def bits(arr, n): return np.unpackbits(arr.view(np.uint8), bitorder="little")[:n].astype(bool)

def main():
    name = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    cfg = CONFIGS[name]
    reduced, specials, main_ = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
    o = oracle_lib.solve(reduced, specials, main_.known, main_.targets, main_.n_vars, cfg.get("secp_solve", False))
    lib = api._engine()
    ph = api.ProblemHandle(reduced, specials, main_.known, main_.targets, main_.n_vars, cfg.get("secp_solve", False))
    g = api.SolveResult(main_.n_vars, full_state=True)
    st = lib.ecne_solve(C.byref(ph.c), C.byref(g.c))
    print("status", st, lib.ecne_last_error())
    V = main_.n_vars
    ou, gu, ok, gk = bits(o.unique_bits, V), bits(g.unique_bits, V), bits(o.known_bits, V), bits(g.known_bits, V)
    print("U diff", int((ou != gu).sum()), "K diff", int((ok != gk).sum()))
    lbd = [w for w in range(V) if toint(o.lb[w]) != toint(g.lb[w]) or toint(o.ub[w]) != toint(g.ub[w])]
    print("bounds diff", len(lbd), "abz diff", int((o.abz != g.abz).sum()), "nvalues diff", int((o.nvalues != g.nvalues).sum()))
    seg = reduced.seg_ptr.astype(np.int64); col = reduced.col
    # wire -> rows
    def rows_of(w1):
        idx = np.nonzero(col == w1)[0]
        segs = np.searchsorted(seg, idx, side="right") - 1
        return sorted(set((segs // 3).tolist()))
    def show_row(r):
        out = []
        for f in range(3):
            s, e = seg[3 * r + f], seg[3 * r + f + 1]
            out.append("{" + ", ".join(f"{col[t]}:{(lambda v: v if v < P//2 else v-P)(toint(reduced.coef[t]))}" for t in range(s, e)) + "}")
        return " * ".join(out[:2]) + " = " + out[2]
    def show_w(w1):
        w = w1 - 1
        return (f"wire {w1}: oracle U={ou[w]} K={ok[w]} lb={toint(o.lb[w])} ub={toint(o.ub[w]) if toint(o.ub[w]) != P-1 else 'p-1'} nv={o.nvalues[w]} abz={o.abz[w]} | "
                f"gpu U={gu[w]} K={gk[w]} lb={toint(g.lb[w])} ub={toint(g.ub[w]) if toint(g.ub[w]) != P-1 else 'p-1'} nv={g.nvalues[w]} abz={g.abz[w]}")
    diff = [w + 1 for w in range(V) if ou[w] != gu[w] or ok[w] != gk[w]] or [w + 1 for w in lbd]
    for w1 in diff[:limit]:
        print(show_w(w1))
        for r in rows_of(w1)[:6]:
            print("    row", r, show_row(r))
            n_other = 0
            for f in range(3):
                for t in range(seg[3*r+f], seg[3*r+f+1]):
                    if col[t] != w1 and col[t] != 1 and n_other < 4:
                        n_other += 1
                        print("        ", show_w(int(col[t])))

main()
