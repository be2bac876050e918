#!/bin/bash
# compute-sanitizer passes over the training path: racecheck / initcheck / synccheck / memcheck.
#   scripts/sanitize.sh <lib.so> "<pytest -k expression>" <tag>
LIB=${1:-code2vec_b200/libc2v_san.so}; KEXPR=${2:-attention_in_the_loss}; TAG=${3:-san}
OUT=gpurun_out/sanitizer; mkdir -p $OUT
export C2V_LIB=$PWD/$LIB
for tool in ${TOOLS:-racecheck initcheck synccheck memcheck}; do
    extra=""
    [ $tool = racecheck ] && extra="--racecheck-report all --racecheck-memcpy-async no"
    [ $tool = initcheck ] && extra=""
    timeout 900 compute-sanitizer --tool $tool $extra --print-limit 200 --log-file $OUT/$TAG.$tool.log \
        python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$KEXPR" > $OUT/$TAG.$tool.pytest.log 2>&1
    echo "$tool rc=$? : $(grep -c 'ERROR\|Error\|error' $OUT/$TAG.$tool.log) error lines; $(tail -1 $OUT/$TAG.$tool.log)"
    tail -2 $OUT/$TAG.$tool.pytest.log
done
