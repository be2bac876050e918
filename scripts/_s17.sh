TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29601 scripts/check_sharded_adam.py 2>&1 | grep -E "TIMING|FAIL|unavailable|rank 0\]" | head -12
$TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-eager 2>/dev/null | tail -1 > gpurun_out/r2_bench_n8.json
$TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-eager --no-e2e --early-reduce 2>/dev/null | tail -1 > gpurun_out/r2_bench_n8_early.json
for f in r2_bench_n8 r2_bench_n8_early; do python -c "
import json; j=json.load(open('gpurun_out/$f.json')); print('$f', 'value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'] if j['e2e'] else None, 'train', {k: j['train'][k] for k in ('ms_per_step','transport','early_region','transport_calibration_ms')}, j['clocks'])"; done
