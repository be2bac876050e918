#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_backward_parity_gpu.py tests/test_dropin_loop_gpu.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/s29_tests.log
for wl in cfg2 cfg4; do WORKLOAD=$wl timeout 100 python scripts/time_train.py 2>&1 | grep -v "Warn\|_warn" > gpurun_out/s29_time_train_$wl.txt; done
tail -4 gpurun_out/s29_tests.log; head -6 gpurun_out/s29_time_train_cfg2.txt; head -4 gpurun_out/s29_time_train_cfg4.txt
