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


# libecne_b200.so carries the solve kernel twice: kernels.cu as it stands (512 threads x 128 registers per block) and a
# second build for bandwidth-bound problems (1024 x 64) in its own namespace (csrc/kernels.cu, "the second build")
SOLVE_VARIANT_1024 = ["-DP1_THREADS=1024", "-DP1_MAX_KS=6", "-DP1_INFLIGHT=2", "-Decne=ecne_v1024",
                      "-DECNE_VARIANT_SUFFIX=v1024"]


def build_engine(force=False, verbose=False, extra_flags=(), out=None, objdir_name="_obj"):
    """extra_flags / out: side builds (tools/build_prof.sh: -DECNE_PROFILE into libecne_b200_prof.so)."""
    out_so = out or ENGINE_SO
    src = _sources(CSRC, (".cu",))
    deps = src + _sources(CSRC, (".cuh", ".h")) + _sources(INC, (".h",))
    if force or _newer(out_so, deps):
        nvcc = nvcc_path()
        if nvcc is None:
            raise RuntimeError("nvcc not found: libecne_b200.so cannot be built (no CPU fallback exists)")
        objdir = os.path.join(PKG, objdir_name)
        os.makedirs(objdir, exist_ok=True)
        base = [nvcc, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INC, "-I", CSRC] + NVCC_ARCH + \
            list(extra_flags)
        jobs = [(f, os.path.join(objdir, os.path.basename(f)[:-3] + ".o"), []) for f in src]
        jobs.append((os.path.join(CSRC, "kernels.cu"), os.path.join(objdir, "kernels_v1024.o"), SOLVE_VARIANT_1024))
        procs = []
        for f, o, extra in jobs:   # the translation units compile side by side
            cmd = base + extra + ["-c", f, "-o", o]
            if verbose:
                print("+", " ".join(cmd), flush=True)
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for cmd, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                sys.stderr.write(out)
                raise RuntimeError("build failed: " + " ".join(cmd))
            if verbose and out:
                print(out)
        _run([nvcc, "-shared", "--cudart", "shared", "-Xcompiler", "-fPIC"] + NVCC_ARCH + ["-o", out_so] +
             [o for _, o, _ in jobs] + ["-ldl"], verbose)
    return out_so


def build_all(force=False, verbose=False):
    return build_host(force, verbose), build_engine(force, verbose), build_oracle(force, verbose)


if __name__ == "__main__":
    if "--side" in sys.argv:   # python -m ecneproject_b200.build --side <out.so> [nvcc flags...]
        i = sys.argv.index("--side")
        build_engine(force=True, verbose=False, extra_flags=sys.argv[i + 2:], out=os.path.abspath(sys.argv[i + 1]),
                     objdir_name="_obj_side")
    else:
        build_all(force="--force" in sys.argv, verbose=True)
