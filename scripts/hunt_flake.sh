#!/bin/bash
# Fresh-process hunt for the intermittent gradient mismatch.
#   scripts/hunt_flake.sh <lib.so> <n_sequential> <n_parallel_rounds> <tag>
# Every run is its own python process executing the backward parity file the way the suite does; C2V_POISON=1 fills
# every uninitialised buffer with NaN bytes first.  Failures keep their full pytest output under gpurun_out/flake/.
LIB=${1:-code2vec_b200/libc2v_b200.so}; NSEQ=${2:-40}; NPAR=${3:-20}; TAG=${4:-run}
OUT=gpurun_out/flake; mkdir -p $OUT
export C2V_LIB=$PWD/$LIB C2V_POISON=1
fail=0; total=0
one() {  # $1 = id
    python -m pytest tests/test_backward_parity_gpu.py -x -q -p no:cacheprovider > $OUT/$TAG.$1.log 2>&1
    rc=$?
    if [ $rc -ne 0 ]; then echo "FAIL $TAG.$1 rc=$rc"; else rm -f $OUT/$TAG.$1.log; fi
    return $rc
}
t0=$(date +%s)
for i in $(seq 1 $NSEQ); do one s$i || fail=$((fail+1)); total=$((total+1)); done
echo "$TAG sequential: $fail / $total failed, $(( $(date +%s) - t0 )) s"
for r in $(seq 1 $NPAR); do
    pids=()
    for j in 1 2 3 4; do one p${r}_$j & pids+=($!); done
    for p in "${pids[@]}"; do wait $p || fail=$((fail+1)); total=$((total+1)); done
done
echo "$TAG total: $fail / $total failed, $(( $(date +%s) - t0 )) s"
