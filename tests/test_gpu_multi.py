"""One process, several GPUs (ecne_init_multi, SURVEY.md §8b "Threading"): the rows are sharded over the GPUs of the
box by the one host thread that calls ecne_solve, and the result must be what one GPU gives.  Runs in a child
process so that the engine state of the test session (bound to one device) is left alone."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["root/trivial_mult", "tornado/merkleTree", "root/bigmult86_3", "secp256k1+bmmp+blt",
         "circomlib/Num2Bits_strict@bitify", "root/poseidon", "tornado/withdraw+pedersen"]


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(n, names):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_check.py"), str(n)] + names,
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    return p.stdout


def test_init_multi_with_one_gpu_is_the_single_gpu_engine():
    out = _run(1, NAMES[:3])
    assert "3/3 configs bit-identical" in out


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs on one box")
def test_one_process_drives_every_gpu_of_the_box():
    n = min(_n_gpus(), 8)
    out = _run(n, NAMES)
    assert f"{len(NAMES)}/{len(NAMES)} configs bit-identical to the oracle and to the one-GPU run on {n} GPUs" in out


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs on one box")
def test_c_example_on_all_gpus(tmp_path):
    """examples/solve_r1cs.c --gpus N: the C ABI from plain C, one process, N GPUs."""
    from ecneproject_b200 import fixtures
    pkg = os.path.join(ROOT, "ecneproject_b200")
    exe = str(tmp_path / "solve_r1cs")
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "solve_r1cs.c"),
                    "-L", pkg, "-lecne_host", "-lecne_b200", "-Wl,-rpath," + pkg, "-o", exe], check=True)
    n = min(_n_gpus(), 8)
    p = subprocess.run([exe, "--gpus", str(n), "--secp-solve", fixtures.path("secp256k1.r1cs"), fixtures.path("bigmultmodp.r1cs"),
                        "BigMultModP", fixtures.path("biglessthan.r1cs"), "BigLessThan"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "sound constraints" in p.stdout and f"{n} GPU(s)" in p.stdout, (p.stdout, p.stderr)
