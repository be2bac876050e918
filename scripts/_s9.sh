TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m 2>/dev/null | head -12
$TR --nproc-per-node 8 --master-port 29601 scripts/check_sharded_adam.py 2>&1 | grep -E "TIMING|PASS|FAIL|unavailable|rank 0\]" | head -20
$TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-eager 2>/dev/null | tail -1 > gpurun_out/r2_bench_n8.json
$TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 --steps 20 --warmup 5 --no-gpu-eager 2>/dev/null | tail -1 > gpurun_out/r2_bench_n4.json
$TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --steps 20 --warmup 5 --no-gpu-eager 2>/dev/null | tail -1 > gpurun_out/r2_bench_n2.json
$TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --workload cfg4 --steps 20 --warmup 5 --no-gpu-eager 2>/dev/null | tail -1 > gpurun_out/r2_bench_cfg4_n8.json
for f in r2_bench_n8 r2_bench_n4 r2_bench_n2 r2_bench_cfg4_n8; do python -c "
import json; j=json.load(open('gpurun_out/$f.json')); print('$f', 'value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e'].get('cpu_binding'), 'train', {k: j['train'][k] for k in ('ms_per_step','transport','transport_calibration_ms')}, j['clocks'])"; done
