"""In-tree native builds (no JIT cache: the built .so files travel to the GPU box with the repo).

  libecne_host.so   g++   csrc/host_r1cs.cpp                (parser + abstraction, no CUDA)
  libecne_b200.so   nvcc  csrc/engine.cu csrc/abi.cu ...    (sm_100a only; the product)
  oracle/_build/liboracle.so  g++ oracle/ecne_oracle.cpp     (test infrastructure, built here too so
                                                              the GPU box needs no compiler run)
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")

HOST_SO = os.path.join(PKG, "libecne_host.so")
ENGINE_SO = os.path.join(PKG, "libecne_b200.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (p.stdout or p.stderr):
        print(p.stdout + p.stderr)


def _sources(dirname, exts):
    out = []
    for f in sorted(os.listdir(dirname)):
        if f.endswith(exts):
            out.append(os.path.join(dirname, f))
    return out


def build_host(force=False, verbose=False):
    src = [os.path.join(CSRC, "host_r1cs.cpp")]
    deps = src + _sources(INC, (".h",))
    if force or _newer(HOST_SO, deps):
        _run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-I", INC, "-o", HOST_SO] + src, verbose)
    return HOST_SO


def build_oracle(force=False, verbose=False):
    odir = os.path.join(ROOT, "oracle")
    src = [os.path.join(odir, "ecne_oracle.cpp")]
    deps = src + _sources(odir, (".h",)) + _sources(INC, (".h",))
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    if force or _newer(ORACLE_SO, deps):
        _run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", INC, "-o", ORACLE_SO] + src, verbose)
    return ORACLE_SO


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


def build_engine(force=False, verbose=False):
    src = _sources(CSRC, (".cu",))
    deps = src + _sources(CSRC, (".cuh", ".h")) + _sources(INC, (".h",))
    if force or _newer(ENGINE_SO, deps):
        nvcc = nvcc_path()
        if nvcc is None:
            raise RuntimeError("nvcc not found: libecne_b200.so cannot be built (no CPU fallback exists)")
        cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-shared", "--cudart", "shared", "-Xcompiler", "-fPIC",
               "-Xptxas", "-v", "-I", INC, "-I", CSRC] + NVCC_ARCH
        cmd += ["-o", ENGINE_SO] + src + ["-ldl"]
        _run(cmd, verbose)
    return ENGINE_SO


def build_all(force=False, verbose=False):
    return build_host(force, verbose), build_engine(force, verbose), build_oracle(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
