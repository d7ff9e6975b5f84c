#!/bin/bash
# Build libecne_b200 with other compile-time parameters into a side library (ECNE_ENGINE_SO makes every Python entry
# load it instead of the default).  The library already carries both builds of the solve kernel (512 x 128 and
# 1024 x 64, engine knob "solve_variant"); this script is for other experiments, e.g. -DP1_INFLIGHT=2.
# Usage:  tools/try_variant.sh <nvcc flags...>
set -e
cd "$(dirname "$0")/.."
OUT=ecneproject_b200/libecne_b200_variant.so
python -m ecneproject_b200.build --side $OUT "$@"
echo "built $OUT with $@; on a B200:"
echo "  export ECNE_ENGINE_SO=\$PWD/$OUT; python -m pytest tests -m gpu -x -q; python tools/stress.py 100 | tail -2; python bench.py"
