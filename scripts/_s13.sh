timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python scripts/time_encode.py tm 2>&1 | tail -1
WORKLOAD=cfg3 python scripts/time_encode.py tm 2>&1 | tail -1
WORKLOAD=cfg5 python scripts/time_encode.py tm 2>&1 | tail -1
python scripts/time_train.py 2>&1 | tail -24
for C in 195299 195296; do ONLY="logits+argmax" C=$C H=100 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:label_gemm_v2 -s 2 -c 1 python scripts/time_label.py 2>&1 | grep -E "dram__|gpu__time|us per call"; done
for wl in cfg2 cfg3 cfg4 cfg5; do python bench.py --workload $wl --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_${wl}_n1.json; python -c "
import json; j=json.load(open('gpurun_out/r2_bench_${wl}_n1.json')); print('$wl', 'value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'kms', j['roofline']['kernel_ms'], 'e2e', j['e2e']['value'], 'train', j['train']['ms_per_step'], 'eager', j['gpu_eager_baseline']['value'], 'cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'], j['clocks'])"; done
