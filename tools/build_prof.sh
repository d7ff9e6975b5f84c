#!/bin/bash
# Build the profile variant of the engine (-DECNE_PROFILE: clock64 stamps per round, per phase, per dense-round stage
# and per long-row evaluation stage) into ecneproject_b200/libecne_b200_prof.so.  Use it with
#   ECNE_ENGINE_SO=$PWD/ecneproject_b200/libecne_b200_prof.so ECNE_DEBUG_PROF=3 python tools/run_one.py ecdsa+secp256k1 2
# (profiles/r01_per_round_cycles_ecdsa.txt is such an output).  Its timings are not bench values.
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -lineinfo -shared --cudart shared -Xcompiler -fPIC -DECNE_PROFILE "$@" -I include -I ecneproject_b200/csrc \
  -gencode arch=compute_100a,code=sm_100a -o ecneproject_b200/libecne_b200_prof.so ecneproject_b200/csrc/*.cu -ldl
ls -la ecneproject_b200/libecne_b200_prof.so
