TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29601 scripts/check_sharded_adam.py 2>&1 | grep -E "TIMING|FAIL|PASS" | head -6
$TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-eager 2>/dev/null | tail -1 > gpurun_out/r2_bench_n8b.json
python -c "
import json; j=json.load(open('gpurun_out/r2_bench_n8b.json')); print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], j['roofline'].get('frac_back_to_back'), 'e2e', j['e2e']['value'], 'train', {k: j['train'][k] for k in ('ms_per_step','transport','transport_calibration_ms')}, j['clocks'])"
