set -x
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_transfer.py -m gpu -x -q 2>&1 | tail -5
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --no-solver > $O/r2c_bench.json 2> $O/r2c_bench.err
tail -3 $O/r2c_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["cpu_baseline"], d["roofline"]["frac"], d["config"])
PY
