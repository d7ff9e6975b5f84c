#!/bin/bash
# compute-sanitizer over small run configurations (SURVEY.md §5 "race detection / sanitizers").
# The reference is single-threaded and has no sanitizer story; the engine's updates are commutative
# atomics, so what is checked here is addressing (memcheck), shared-memory hazards in the block-level
# stages (racecheck) and barrier divergence (synccheck).
#   memcheck               five configurations at the full grid (148 blocks), plus abstraction on the device,
#                          plus a sharded solve on every GPU of the box when it has more than one
#   racecheck / synccheck  the solve kernel squeezed into TWO blocks (engine knob grid_blocks=2): under the tools'
#                          shared-memory / barrier tracking the spin barriers of the full persistent grid do not finish
#                          a solve within minutes; with two blocks every barrier, solo stretch and phase still runs
# Usage (on a B200 box):  tools/sanitize.sh [outdir]     -- writes <outdir>/sanitize_<tool>_<config>.log + a summary
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SUMMARY="$OUT/sanitize_summary.txt"
: > "$SUMMARY"
run() {  # tool, tag, timeout, command...
  local tool=$1 tag=$2 tmo=$3; shift 3
  local log="$OUT/sanitize_${tool}_${tag}.log"
  timeout "$tmo" compute-sanitizer --tool "$tool" --print-limit 20 "$@" > "$log" 2>&1
  local rc=$?
  local line=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  local solve=$(grep -E "rep0 st=|configs bit-identical|passed|file->verdict" "$log" | head -1 | cut -c1-110)
  echo "$tool | $tag | rc=$rc | ${line:-no summary line} | ${solve:-no result line}" >> "$SUMMARY"
}
for cfg in "root/poseidon" "tornado/merkleTree" "secp256k1+bmmp+blt" "tornado/withdraw+pedersen" "target/division"; do
  run memcheck "$(echo "$cfg" | tr '/+' '__')" 300 python tools/run_one.py "$cfg" 1
done
run memcheck device_abstraction_secp256k1 300 python tools/file_to_verdict.py secp256k1+bmmp+blt 1
for tool in racecheck synccheck; do
  for cfg in "root/multiplexer_33" "root/bigmult86_3" "tornado/merkleTree" "secp256k1+bmmp+blt"; do
    run $tool "2blocks_$(echo "$cfg" | tr '/+' '__')" 420 python tools/run_one.py "$cfg" 1 grid_blocks=2
  done
done
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -gt 1 ]; then
  run memcheck "sharded_${NG}gpus_one_process" 600 python tools/multi_check.py "$NG" root/bigmult86_3 tornado/merkleTree secp256k1+bmmp+blt
  run synccheck "sharded_${NG}gpus_one_process_2blocks" 600 env ECNE_GRID_BLOCKS=2 python tools/multi_check.py "$NG" root/bigmult86_3 tornado/merkleTree
fi
cat "$SUMMARY"
