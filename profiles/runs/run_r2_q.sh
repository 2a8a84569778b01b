#!/bin/bash
# block-inverse GS: parity (default + per-phase form), stamps, then the C2 bench line
O=gpurun_out; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_gpu_matrix.py tests/test_gpu_solver.py -m gpu -x -q > $O/r2q_pytest.log 2>&1
tail -15 $O/r2q_pytest.log
HOT_GS_COOP=0 timeout -s KILL 600 python -m pytest tests/test_gpu_matrix.py -m gpu -x -q -k "smoother_parity or vcycle_parity" > $O/r2q_pytest_nocoop.log 2>&1
tail -15 $O/r2q_pytest_nocoop.log
HOT_GS_DEBUG=3 timeout -s KILL 600 python profiles/prof_gs.py 2>&1 | grep "dbg\|vcycle" | head -4 > $O/r2q_gx_stamps.txt
cut -c1-700 $O/r2q_gx_stamps.txt
timeout -s KILL 600 python bench.py --cpu-reps 0 > $O/r2q_bench.json 2> $O/r2q_bench.err
tail -3 $O/r2q_bench.err
python - <<'PY'
import json
for f in ("r2q_bench.json",):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1]); s = d["solver_kernels"]
        print(f, "vcycle", s["vcycle"]["ms"], "gs", [round(g["ms"], 4) for g in s["gs_smooth"]], "table", s["vcycle"]["per_level_ms[smooth,restrict,prolongate,merge]"],
              "build_mg", s["build_mg_ms"], "substep", s["hot_substep"]["steady_ms"], [x["lbfgs_iterations"] for x in s["hot_substep"]["substeps"]])
    except Exception as e:
        print(f, "failed", e)
PY
