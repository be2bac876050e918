#!/bin/bash
#  scripts/hunt_diag.sh <rounds of 4 parallel> <tag> [ENV=VAL ...]    (environment of the diagnostic processes)
N=${1:-40}; TAG=${2:-d}; shift 2
OUT=gpurun_out/flake; mkdir -p $OUT
for kv in "$@"; do export "$kv"; done
export C2V_POISON=1
fail=0; total=0; t0=$(date +%s)
for r in $(seq 1 $N); do
    pids=()
    for j in 1 2 3 4; do ( python scripts/flake_diag.py > $OUT/$TAG.$r.$j.log 2>&1 || { echo "FAIL $TAG.$r.$j $(grep -o 'FIRST.*' $OUT/$TAG.$r.$j.log | head -1)"; exit 1; }; rm -f $OUT/$TAG.$r.$j.log ) & pids+=($!); done
    for p in "${pids[@]}"; do wait $p || fail=$((fail+1)); total=$((total+1)); done
done
echo "== $TAG [$*]: $fail / $total failed, $(( $(date +%s) - t0 )) s"
