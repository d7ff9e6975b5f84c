#!/bin/bash
# Build the profile variant of the engine (-DECNE_PROFILE: clock64 stamps per round, per phase, per dense-round stage
# and per long-row evaluation stage) into ecneproject_b200/libecne_b200_prof.so.  Use it with
#   ECNE_ENGINE_SO=$PWD/ecneproject_b200/libecne_b200_prof.so ECNE_DEBUG_PROF=3 python tools/run_one.py ecdsa+secp256k1 2
# (profiles/r0*_per_round_cycles_ecdsa.txt are such outputs).  Its timings are not bench values.
set -e
cd "$(dirname "$0")/.."
python -m ecneproject_b200.build --side ecneproject_b200/libecne_b200_prof.so -DECNE_PROFILE "$@"
ls -la ecneproject_b200/libecne_b200_prof.so
