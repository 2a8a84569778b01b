# round 2, call A: parity of the new persistent scatter + first timing
set -x
O=gpurun_out
mkdir -p $O
(timeout -s KILL 900 python -m pytest tests -m gpu -x -q) > $O/r2a_pytest.log 2>&1
tail -5 $O/r2a_pytest.log
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --cpu-reps 0 --no-solver > $O/r2a_bench_ws.json 2> $O/r2a_bench_ws.err
HOT_SCATTER=plane timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --cpu-reps 0 --no-solver > $O/r2a_bench_plane.json 2> $O/r2a_bench_plane.err
tail -3 $O/r2a_bench_ws.err
python - <<'PY'
import json
for n in ("ws","plane"):
    try:
        d=json.load(open(f"gpurun_out/r2a_bench_{n}.json"))
        print(n, d["value"], d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["per_kernel"].items()}, d["roofline"]["frac"])
    except Exception as e: print(n, "failed", e)
PY
