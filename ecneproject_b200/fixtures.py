"""Locate the staged circuit corpus (data/, see tools/stage_fixtures.py).

`path("ecne_circomlib_tests/Poseidon@poseidon.r1cs")` returns a real file path, unpacking
data/r1cs_corpus.tar.xz / data/ecdsa.r1cs.xz into data/_cache/ on first use.  Nothing here reads
/root/reference: that tree does not exist on the GPU box.
"""
import json
import lzma
import os
import tarfile
import io
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")
CACHE = os.path.join(DATA, "_cache")

_manifest = None


def manifest():
    global _manifest
    if _manifest is None:
        with open(os.path.join(DATA, "MANIFEST.json")) as f:
            _manifest = json.load(f)
    return _manifest


def _atomic_write(dst, data):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    fd, tmp = tempfile.mkstemp(dir=os.path.dirname(dst))
    with os.fdopen(fd, "wb") as f:
        f.write(data)
    os.replace(tmp, dst)


def _unpack_corpus():
    marker = os.path.join(CACHE, ".corpus_ok")
    if os.path.exists(marker):
        return
    with open(os.path.join(DATA, "r1cs_corpus.tar.xz"), "rb") as f:
        raw = lzma.decompress(f.read())
    with tarfile.open(fileobj=io.BytesIO(raw)) as tf:
        for m in tf.getmembers():
            _atomic_write(os.path.join(CACHE, m.name), tf.extractfile(m).read())
    _atomic_write(marker, b"ok")


def path(rel):
    """Absolute path of a staged fixture, e.g. 'poseidon.r1cs' or 'ecdsa.r1cs'."""
    m = manifest()
    if rel not in m:
        raise FileNotFoundError(f"{rel} is not in data/MANIFEST.json")
    dst = os.path.join(CACHE, rel)
    if os.path.exists(dst) and os.path.getsize(dst) == m[rel]["bytes"]:
        return dst
    if rel == "ecdsa.r1cs":
        with open(os.path.join(DATA, "ecdsa.r1cs.xz"), "rb") as f:
            _atomic_write(dst, lzma.decompress(f.read()))
    else:
        _unpack_corpus()
    return dst


def circomlib():
    """Sorted fixture names under ecne_circomlib_tests/ (without directory and extension)."""
    pre = "ecne_circomlib_tests/"
    return sorted(k[len(pre):-5] for k in manifest() if k.startswith(pre) and k.endswith(".r1cs"))
