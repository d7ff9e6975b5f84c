#!/bin/bash
# compute-sanitizer over a few small run configurations (SURVEY.md §5 "race detection / sanitizers").
# The reference is single-threaded and has no sanitizer story; the engine's updates are commutative
# atomics, so what is checked here is addressing (memcheck), shared-memory hazards in the block-level
# stages (racecheck), barrier divergence (synccheck) and reads of uninitialised device memory (initcheck).
# Usage (on a B200 box):  tools/sanitize.sh [outdir]     -- writes <outdir>/sanitize_<tool>_<config>.log
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CONFIGS=("root/poseidon" "tornado/merkleTree" "secp256k1+bmmp+blt" "tornado/withdraw+pedersen" "target/division")
# memcheck only by default: racecheck did not finish a solve of the persistent cooperative kernel within
# 240 s on a B200 (its spin barriers crawl under the tool's shared-memory tracking); ask for it explicitly.
TOOLS=(${SANITIZE_TOOLS:-memcheck})
SUMMARY="$OUT/sanitize_summary.txt"
: > "$SUMMARY"
for tool in "${TOOLS[@]}"; do
  for cfg in "${CONFIGS[@]}"; do
    tag=$(echo "$cfg" | tr '/+' '__')
    log="$OUT/sanitize_${tool}_${tag}.log"
    timeout 240 compute-sanitizer --tool "$tool" --print-limit 20 python tools/run_one.py "$cfg" 1 > "$log" 2>&1
    rc=$?
    line=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    solve=$(grep -E "rep0 st=" "$log" | head -1 | cut -c1-80)
    echo "$tool | $cfg | rc=$rc | ${line:-no summary line} | ${solve:-no solve line}" >> "$SUMMARY"
  done
done
cat "$SUMMARY"
