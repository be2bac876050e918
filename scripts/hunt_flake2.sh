#!/bin/bash
#   scripts/hunt_flake2.sh <lib.so> <rounds of 4 parallel single-test processes> <full-suite runs> <tag>
LIB=${1:-code2vec_b200/libc2v_b200.so}; NPAR=${2:-50}; NFULL=${3:-10}; TAG=${4:-h2}
OUT=gpurun_out/flake; mkdir -p $OUT
export C2V_LIB=$PWD/$LIB C2V_POISON=1
fail=0; total=0; t0=$(date +%s)
for r in $(seq 1 $NPAR); do
    pids=()
    for j in 1 2 3 4; do ( python scripts/${FLAKE_SCRIPT:-flake_once.py} > $OUT/$TAG.p${r}_$j.log 2>&1 || { echo "FAIL $TAG.p${r}_$j"; exit 1; }; rm -f $OUT/$TAG.p${r}_$j.log ) & pids+=($!); done
    for p in "${pids[@]}"; do wait $p || fail=$((fail+1)); total=$((total+1)); done
done
echo "$TAG single-test: $fail / $total failed, $(( $(date +%s) - t0 )) s"
ffail=0
for i in $(seq 1 $NFULL); do
    python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/$TAG.full$i.log 2>&1 || { ffail=$((ffail+1)); echo "FAIL full suite run $i: $(grep -E "FAILED|failed" $OUT/$TAG.full$i.log | head -3)"; continue; }
    rm -f $OUT/$TAG.full$i.log
done
echo "$TAG full suite: $ffail / $NFULL failed, $(( $(date +%s) - t0 )) s"
