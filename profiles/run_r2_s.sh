#!/bin/bash
O=gpurun_out; mkdir -p $O
for cfg in 0 1; do
HOT_GS_COOP=0 HOT_GX_BLOCK=$cfg timeout -s KILL 600 python -m pytest tests/test_gpu_matrix.py -m gpu -x -q -k "smoother_parity or vcycle_parity" 2>&1 | tail -1
HOT_GX_BLOCK=$cfg timeout -s KILL 600 python bench.py --cpu-reps 0 > $O/r2s_bench_$cfg.json 2> $O/r2s_bench_$cfg.err
done
HOT_GS_COOP=0 timeout -s KILL 600 python bench.py --cpu-reps 0 > $O/r2s_bench_2.json 2> $O/r2s_bench_2.err
python - <<'PY'
import json
for f in ("r2s_bench_0.json","r2s_bench_1.json","r2s_bench_2.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1]); s = d["solver_kernels"]
        print(f, "vcycle", s["vcycle"]["ms"], "gs", [round(g["ms"], 4) for g in s["gs_smooth"]], "substep", s["hot_substep"]["steady_ms"])
    except Exception as e:
        print(f, "failed", e)
PY
