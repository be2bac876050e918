#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
B=1024 C=195299 H=100 ONLY="logits+argmax" timeout 200 $NCU -k regex:label_gemm_v2 -s 3 -c 1 -f -o gpurun_out/r2_label_cfg3_lanes python scripts/time_label.py > gpurun_out/s22_a.log 2>&1
WORKLOAD=cfg4 timeout 300 $NCU -k 'regex:backward_rows_lite|backward_dc_tc|backward_dw_tc' -s 6 -c 3 -f -o gpurun_out/r2_backward_cfg4 python scripts/time_train.py > gpurun_out/s22_b.log 2>&1
timeout 200 $NCU -k 'regex:backward_rows_lite' -s 2 -c 1 -f -o gpurun_out/r2_rows_lite_cfg2 python scripts/time_train.py > gpurun_out/s22_c.log 2>&1
WORKLOAD=cfg4 timeout 120 python scripts/time_train.py > gpurun_out/r2_time_train_cfg4.txt 2>&1
WORKLOAD=cfg3 timeout 120 python scripts/time_train.py > gpurun_out/r2_time_train_cfg3.txt 2>&1
tail -3 gpurun_out/s22_a.log gpurun_out/s22_b.log gpurun_out/s22_c.log; cat gpurun_out/r2_time_train_cfg4.txt
