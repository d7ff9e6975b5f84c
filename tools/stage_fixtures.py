#!/usr/bin/env python3
"""Stage the reference's .r1cs circuit corpus into data/ (run once, in the build container).

The circuits are the *inputs* of the hot path (SURVEY.md Appendix B / §8d "Concrete inputs"):
compiled circom constraint systems, not source code.  /root/reference does not exist on the GPU
box, so the corpus is re-packed here (xz, our own container layout) and committed:

  data/r1cs_corpus.tar.xz   every *.r1cs under /root/reference (80-odd small files, paths kept)
  data/ecdsa.r1cs.xz        ecdsa.r1cs from ecdsa_r1cs.tar.gz (142 507 804 B; Artifacts.toml sha256
                            of the tarball 522eba8d...)
  data/MANIFEST.json        path -> {bytes, sha256} of every staged file

`ecneproject_b200.fixtures` unpacks them lazily into data/_cache/ (git-ignored).
"""
import hashlib, io, json, lzma, os, sys, tarfile

REF = os.environ.get("ECNE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "data")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    os.makedirs(DATA, exist_ok=True)
    manifest = {}
    paths = []
    for d, _, files in os.walk(REF):
        for f in files:
            if f.endswith(".r1cs") or f.endswith(".sym"):
                paths.append(os.path.relpath(os.path.join(d, f), REF))
    paths.sort()
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as tf:
        for p in paths:
            b = open(os.path.join(REF, p), "rb").read()
            manifest[p] = {"bytes": len(b), "sha256": sha(b)}
            ti = tarfile.TarInfo(p)
            ti.size = len(b)
            tf.addfile(ti, io.BytesIO(b))
    with open(os.path.join(DATA, "r1cs_corpus.tar.xz"), "wb") as f:
        f.write(lzma.compress(buf.getvalue(), preset=9 | lzma.PRESET_EXTREME))
    with tarfile.open(os.path.join(REF, "ecdsa_r1cs.tar.gz")) as tf:
        b = tf.extractfile("ecdsa.r1cs").read()
    manifest["ecdsa.r1cs"] = {"bytes": len(b), "sha256": sha(b)}
    with open(os.path.join(DATA, "ecdsa.r1cs.xz"), "wb") as f:
        f.write(lzma.compress(b, preset=6))
    with open(os.path.join(DATA, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("staged", len(manifest), "files")


if __name__ == "__main__":
    sys.exit(main())
