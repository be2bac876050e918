set -x
python scripts/time_label.py 2>&1 | tail -8
C=195299 H=100 python scripts/time_label.py 2>&1 | tail -8
python scripts/time_train.py 2>&1 | tail -22
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench_steps4.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-eager --train-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_tm -s 3 -c 1 -o gpurun_out/r2_encode_tm python scripts/time_encode.py tm > /dev/null 2>&1
ONLY="logits+argmax" C=195299 H=100 ncu --set full --clock-control none --import-source on -k regex:label_gemm_v2 -s 2 -c 1 -o gpurun_out/r2_label_cfg3 python scripts/time_label.py > /dev/null 2>&1
ONLY="no logits" C=195299 H=100 ncu --set full --clock-control none --import-source on -k regex:label_gemm_v2 -s 2 -c 1 -o gpurun_out/r2_label_loss_cfg3 python scripts/time_label.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
for wl in cfg3 cfg4 cfg5; do python bench.py --workload $wl --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_${wl}_n1.json; python -c "
import json; j=json.load(open('gpurun_out/r2_bench_${wl}_n1.json')); print('$wl', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['kernel_ms'], j['e2e']['value'], j['train']['ms_per_step'] if j['train'] else None, j['clocks'])"; done
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1.json; python -c "
import json; j=json.load(open('gpurun_out/r2_bench_n1.json')); print('cfg2', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['kernel_ms'], j['e2e']['value'], j['train'], j['gpu_eager_baseline']['value'], j['cpu_baseline']['value'], j['clocks'])"
