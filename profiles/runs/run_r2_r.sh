#!/bin/bash
# full single-GPU test suite + C2 bench line after the Gauss-Seidel rework
O=gpurun_out; mkdir -p $O
(time timeout -s KILL 1200 python -m pytest tests -m gpu -q -x) > $O/r2r_pytest_gpu.log 2>&1
tail -8 $O/r2r_pytest_gpu.log
timeout -s KILL 600 python bench.py --cpu-reps 0 > $O/r2r_bench.json 2> $O/r2r_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1]); s = d["solver_kernels"]
print("value", d["value"], "e2e", d["e2e"]["value"], "vcycle", s["vcycle"]["ms"], "gs", [round(g["ms"], 4) for g in s["gs_smooth"]], "coarse", s["coarse_pcg_alone"]["ms"], s["coarse_pcg_alone"]["cg_iters"],
      "build", s["build_matrix_ms"], s["build_mg_ms"], "substep", s["hot_substep"]["steady_ms"])
PY
