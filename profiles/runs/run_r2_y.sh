#!/bin/bash
# final code: whole GPU suite, smoke, the C2 bench line (both arms)
O=gpurun_out; mkdir -p $O
(time timeout -s KILL 900 python -m pytest tests -m gpu -q) > $O/r2y_pytest_gpu.log 2>&1
tail -4 $O/r2y_pytest_gpu.log
timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout -s KILL 400 python bench.py > $O/r2y_bench_c2.json 2> $O/r2y_bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2y_bench_c2.json").read().strip().splitlines()[-1]); s=d["solver_kernels"]
print(round(d["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), round(d["e2e"]["value"],1), "cpu", round(d["cpu_baseline"]["value"],1), "vcycle", s["vcycle"]["ms"], s["vcycle"]["frac"], "gs", [(round(g["ms"],3), round(g["frac"],2)) for g in s["gs_smooth"]], "substep", s["hot_substep"]["steady_ms"], [x["lbfgs_iterations"] for x in s["hot_substep"]["substeps"]])
PY
