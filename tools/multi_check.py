#!/usr/bin/env python3
"""Sharded parity check in ONE process (ecne_init_multi): python tools/multi_check.py <n_gpus> [--all] [--with-ecdsa] [configs...]
Every configuration is solved on n GPUs driven by this one process and compared with the committed oracle goldens:
verdict, counts, SHA-256 of the `unique` (and, outside the two documented circuits, `is_known`) bitmaps, and the
round counters must equal those of a one-GPU solve of the same process (row visits within 5 %)."""
import ctypes as C, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ecneproject_b200 import api, fixtures
from configs import CONFIGS

n_gpus = int(sys.argv[1])
flags = [a for a in sys.argv[2:] if a.startswith("--")]
names = [a for a in sys.argv[2:] if not a.startswith("--")]
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_goldens.json")))
if "--all" in flags:
    names = [n for n, c in CONFIGS.items() if gold[n].get("status", 0) == 0 and (not c.get("big") or "--with-ecdsa" in flags)]
names = names or ["root/trivial_mult", "circomlib/Poseidon@poseidon", "tornado/merkleTree", "root/bigmult86_3",
                  "root/multiplexer_33", "secp256k1+bmmp+blt", "tornado/withdraw+pedersen", "root/poseidon",
                  "circomlib/Num2Bits_strict@bitify", "circomlib/EdDSAPoseidonVerifier@eddsaposeidon"]
SCHED = {"circomlib/Bits2Point_Strict@pointbits", "circomlib/EdDSAVerifier@eddsa"}
from ecneproject_b200 import _abi
lib = _abi.engine_lib()
# the fixtures are far below "shard_min_rows": force the sharded path, which is what this tool checks
assert lib.ecne_set_option(b"shard_min_rows", 0) == 0
if os.environ.get("ECNE_GRID_BLOCKS"):  # sanitizer runs squeeze the solve kernels into a few blocks
    assert lib.ecne_set_option(b"grid_blocks", int(os.environ["ECNE_GRID_BLOCKS"])) == 0


def solve(n, ph, n_vars):
    api.init_multi(n) if n > 1 else (lib.ecne_init(0) == 0 or sys.exit(lib.ecne_last_error()))
    res = api.SolveResult(n_vars)
    t0 = time.perf_counter()
    st = lib.ecne_solve(C.byref(ph.c), C.byref(res.c))
    return st, res, time.perf_counter() - t0


bad = 0
for name in names:
    cfg = CONFIGS[name]
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
    ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
    st1, r1, _ = solve(1, ph, main.n_vars)
    st, res, dt = solve(n_gpus, ph, main.n_vars)
    g = gold[name]
    ok = st == 0 and st1 == 0
    why = ""
    if ok:
        checks = {
            "sha_unique": hashlib.sha256(res.unique_bytes()).hexdigest() == g["sha_unique"],
            "sha_known": name in SCHED or hashlib.sha256(res.known_bytes()).hexdigest() == g["sha_known"],
            "verdict": bool(res.c.verdict) == g["verdict"], "n_unique": res.c.n_unique == g["n_unique"],
            "same bitmaps as 1 GPU": res.unique_bytes() == r1.unique_bytes() and res.known_bytes() == r1.known_bytes(),
            "same rounds as 1 GPU": (res.c.outer_rounds, res.c.inner_rounds) == (r1.c.outer_rounds, r1.c.inner_rounds),
            # (the row visits of a frontier-driven round depend on which of two racing rows logs a wire first — one
            # record or two — so the count moves by a few per cent from run to run, on one GPU as well)
            # A sharded round counts a wire that rows of two ranks changed twice when it sizes the next round, so a
            # round near the dense / frontier threshold can be swept densely on N GPUs and not on one: at most one
            # extra sweep of the rows per such round, never a different result.
            "same evals as 1 GPU": abs(int(res.c.constraint_evals) - int(r1.c.constraint_evals)) <= 0.05 * r1.c.constraint_evals + n_gpus * reduced.n_rows,
            "gpus_used": res.c.gpus_used == n_gpus, "sharded": res.c.sharded == (1 if n_gpus > 1 else 0),
        }
        why = ",".join(k for k, v in checks.items() if not v)
        ok = not why
    print(("OK   " if ok else "FAIL ") + f"{name} gpus={res.c.gpus_used} st={st} verdict={bool(res.c.verdict)} n_unique={res.c.n_unique} "
          f"(gold {g['n_unique']}) outer={res.c.outer_rounds} inner={res.c.inner_rounds} dense={res.c.dense_rounds} (1 GPU {r1.c.dense_rounds}) evals={res.c.constraint_evals} (1 GPU {r1.c.constraint_evals}) "
          f"sweep={res.c.ms_sweep:.3f}ms (1 GPU {r1.c.ms_sweep:.3f}ms) call={dt*1e3:.2f}ms {why} {lib.ecne_last_error().decode() if st else ''}", flush=True)
    bad += 0 if ok else 1
print(f"{len(names) - bad}/{len(names)} configs bit-identical to the oracle and to the one-GPU run on {n_gpus} GPUs driven by one process")
lib.ecne_shutdown()
sys.exit(1 if bad else 0)
