#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout -s KILL 120 python bench.py --steps 5 --warmup 3 --cpu-reps 0 --no-solver > $O/r2z_bench_1gpu_quick.json 2> $O/r2z_bench_1gpu_quick.err
python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench_1gpu_quick.json').read().strip().splitlines()[-1]); print('1 GPU', round(d['value']), d['solver_kernels'])"
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2z_bench_2gpu.json 2> $O/r2z_bench_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench_2gpu.json').read().strip().splitlines()[-1]); s=d['solver_kernels'] or {}
print('2 GPU', round(d['value']), round(d['ms_per_step'],4), 'vcycle', d['vcycle_ms'], 'substep', (s.get('hot_substep') or {}).get('steady_ms'), s.get('error'))"
tail -2 $O/r2z_bench_2gpu.err | cut -c1-300
