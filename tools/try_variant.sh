#!/bin/bash
# Build libecne_b200 with other solve-kernel launch parameters into a side library and run the GPU suite, the
# parity sweep and the bench on it (ECNE_ENGINE_SO makes every Python entry load it instead of the default).
# Usage:  tools/try_variant.sh [-DP1_THREADS=1024 -DP1_MAX_KS=6 -DP1_INFLIGHT=2]      (build here, then run the printed
#         command under gpurun; DESIGN.md §8 item 0 has the numbers this was first used for)
set -e
cd "$(dirname "$0")/.."
FLAGS=${@:--DP1_THREADS=1024 -DP1_MAX_KS=6 -DP1_INFLIGHT=2}
OUT=ecneproject_b200/libecne_b200_variant.so
nvcc -O3 -std=c++17 -lineinfo -shared --cudart shared -Xcompiler -fPIC -Xptxas -v $FLAGS -I include -I ecneproject_b200/csrc \
  -gencode arch=compute_100a,code=sm_100a -o $OUT ecneproject_b200/csrc/*.cu -ldl 2>&1 | grep -A2 "k_solveEji\|warp_solo\|sparse_round" | grep "spill\|registers"
echo "built $OUT with $FLAGS; on a B200:"
echo "  export ECNE_ENGINE_SO=\$PWD/$OUT; python -m pytest tests -m gpu -x -q; python tools/gpu_parity.py --with-ecdsa | tail -3; python tools/stress.py 100 | tail -2; python bench.py"
