#!/usr/bin/env python3
"""Sharded parity check, run under torchrun on N GPUs:
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [configs...]"""
import ctypes as C, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from ecneproject_b200 import api, fixtures, dist as edist
from configs import CONFIGS

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = edist.init_from_torch(local)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_goldens.json")))
names = sys.argv[1:] or ["root/trivial_mult", "circomlib/Poseidon@poseidon", "tornado/merkleTree", "root/bigmult86_3",
                         "root/multiplexer_33", "secp256k1+bmmp+blt", "tornado/withdraw+pedersen", "root/poseidon",
                         "circomlib/Num2Bits_strict@bitify", "circomlib/EdDSAPoseidonVerifier@eddsaposeidon"]
lib = api._engine()
# the fixtures are far below "shard_min_rows": force the sharded path, which is what this tool checks
assert lib.ecne_set_option(b"shard_min_rows", 0) == 0
bad = 0
for name in names:
    cfg = CONFIGS[name]
    reduced, specials, main = api.prepare(fixtures.path(cfg["main"]), [fixtures.path(t) for t in cfg.get("trusted", [])], cfg.get("trusted_names", []))
    ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars, cfg.get("secp_solve", False))
    lo, hi = edist.shard_rows(ph, rank, world)
    h = C.c_void_p()
    st = lib.ecne_upload(C.byref(ph.c), C.byref(h))
    assert st == 0, lib.ecne_last_error()
    res = api.SolveResult(main.n_vars)
    dist.barrier()
    t0 = time.perf_counter()
    st = lib.ecne_solve_resident(h, C.byref(res.c))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = st == 0
    g = gold[name]
    if ok:
        ok = (hashlib.sha256(res.unique_bytes()).hexdigest() == g["sha_unique"] and bool(res.c.verdict) == g["verdict"]
              and res.c.n_unique == g["n_unique"])
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    if rank == 0:
        print(("OK   " if flag.item() == 0 else "FAIL ") + f"{name} world={world} rows[{lo},{hi}) st={st} verdict={bool(res.c.verdict)} "
              f"n_unique={res.c.n_unique} (gold {g['n_unique']}) outer={res.c.outer_rounds} inner={res.c.inner_rounds} "
              f"solve={res.c.ms_solve:.3f}ms sweep={res.c.ms_sweep:.3f}ms wall={dt*1e3:.2f}ms err={lib.ecne_last_error().decode() if st else ''}", flush=True)
    bad += int(flag.item() != 0)
    lib.ecne_free_resident(h)
if rank == 0:
    print(f"{len(names) - bad}/{len(names)} sharded configs bit-identical to the oracle on {world} GPUs")
dist.destroy_process_group()
sys.exit(1 if bad else 0)
