"""Sharded parity on a box with at least two GPUs: one process per GPU under torchrun, rows sharded by
stored-term balance, update records exchanged over NVLink inside the solve kernel; verdict, counts and the
SHA-256 of the `unique` bitmap must equal the committed oracle goldens on every rank (tools/dist_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs on one box")
def test_two_rank_sharded_solve_matches_goldens():
    names = ["root/trivial_mult", "tornado/merkleTree", "root/bigmult86_3", "secp256k1+bmmp+blt",
             "circomlib/Num2Bits_strict@bitify", "root/poseidon"]
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(ROOT, "tools", "dist_check.py")] + names,
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    assert f"{len(names)}/{len(names)} sharded configs bit-identical to the oracle on 2 GPUs" in p.stdout
