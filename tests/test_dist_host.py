"""Host-side logic of the N>1 path, run on CPU with world_size-2/4 `gloo` groups:
row-range sharding by stored-term balance (ecne_shard_rows) and the communicator bootstrap that
ecneproject_b200.dist performs over torch.distributed.  No GPU, no compute call."""
import os
import socket

import numpy as np
import pytest

from ecneproject_b200 import api, fixtures
from ecneproject_b200 import dist as edist
from configs import CONFIGS


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, names, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) bootstrap: rank 0's 128-byte id reaches every rank unchanged
        uid = edist.broadcast_unique_id(lambda: bytes(range(128)), rank)
        assert uid == bytes(range(128)), "unique id was not broadcast intact"
        out = []
        for name in names:
            cfg = CONFIGS[name]
            reduced, specials, main = api.prepare(fixtures.path(cfg["main"]),
                                                  [fixtures.path(t) for t in cfg.get("trusted", [])],
                                                  cfg.get("trusted_names", []))
            ph = api.ProblemHandle(reduced, specials, main.known, main.targets, main.n_vars,
                                   cfg.get("secp_solve", False))
            lo, hi = edist.shard_rows(ph, rank, world)
            # (2) every rank learns every range through the process group
            mine = torch.tensor([lo, hi], dtype=torch.int64)
            allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(allr, mine)
            ranges = [(int(t[0]), int(t[1])) for t in allr]
            seg = reduced.seg_ptr.astype(np.int64)
            terms = [int(seg[3 * b] - seg[3 * a]) for a, b in ranges]
            longest = int(np.max(seg[3::3] - seg[0:-1:3])) if reduced.n_rows else 0
            out.append((name, reduced.n_rows, ranges, terms, int(seg[-1]), longest))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_row_range_sharding_over_gloo(world):
    import torch.multiprocessing as mp
    names = ["tornado/merkleTree", "tornado/withdraw+pedersen", "root/multiplexer_33", "secp256k1+bmmp+blt"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, names, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, out = q.get(timeout=180)
        got[rank] = out
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for i, name in enumerate(names):
        ref = got[0][i]
        for r in range(1, world):
            assert got[r][i] == ref, f"{name}: ranks disagree on the partition"
        _, n_rows, ranges, terms, total, longest = ref
        # contiguous, ordered, covering [0, N)
        assert ranges[0][0] == 0 and ranges[-1][1] == n_rows
        for a, b in zip(ranges, ranges[1:]):
            assert a[1] == b[0] and a[0] <= a[1]
        assert sum(terms) == total
        # balanced by stored terms to within one row
        for t in terms:
            assert abs(t - total / world) <= longest + 1, (name, terms, total)


def test_shard_rows_edge_cases():
    from helpers import MiniR1CS
    # empty system: every rank gets the empty range
    empty = MiniR1CS([], 1, [1], [])
    ph = api.ProblemHandle(empty, [], empty.known, empty.targets, 1, False)
    for w in (1, 2, 8):
        for r in range(w):
            assert edist.shard_rows(ph, r, w) == (0, 0)
    # fewer rows than ranks: ranges still tile [0, N) and stay ordered
    rows = [({2: 1}, {3: 1}, {4: 1}), ({2: 1}, {2: 1}, {3: 1})]
    m = MiniR1CS(rows, 4, [1], [4])
    ph = api.ProblemHandle(m, [], m.known, m.targets, 4, False)
    cuts = [edist.shard_rows(ph, r, 8) for r in range(8)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 2
    for a, b in zip(cuts, cuts[1:]):
        assert a[1] == b[0]
    # bad rank / world are argument errors, not crashes
    with pytest.raises(ValueError):
        edist.shard_rows(ph, 3, 2)
    with pytest.raises(ValueError):
        edist.shard_rows(ph, 0, 0)
