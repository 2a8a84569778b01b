#!/bin/bash
# after the update-stream change: parity, smoke, smoother timing, then the C3 / C5 stand-in bench lines
O=gpurun_out; mkdir -p $O
timeout -s KILL 300 python -m pytest tests/test_gpu_matrix.py tests/test_gpu_solver.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
HOT_GS_COOP=0 timeout -s KILL 150 python -m pytest tests/test_gpu_matrix.py -m gpu -x -q -k "smoother_parity or vcycle_parity" 2>&1 | tail -1
timeout -s KILL 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 90 python profiles/prof_smooth.py 2>&1 | tail -1
timeout -s KILL 300 python bench.py --cpu-reps 0 > $O/r2w_bench_c2.json 2> $O/r2w_bench_c2.err
timeout -s KILL 300 python bench.py --workload c3 --cpu-reps 0 > $O/r2w_bench_c3.json 2> $O/r2w_bench_c3.err
timeout -s KILL 400 python bench.py --workload c5 --cpu-reps 0 > $O/r2w_bench_c5.json 2> $O/r2w_bench_c5.err
python - <<'PY'
import json
for f in ("r2w_bench_c2.json","r2w_bench_c3.json","r2w_bench_c5.json"):
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1]); s=d.get("solver_kernels") or {}
        print(f, round(d["value"]), round(d["ms_per_step"],4), d["roofline"]["per_kernel"]["p2g"]["frac"], round(d["e2e"]["value"],1), "vcycle", (s.get("vcycle") or {}).get("ms"), "hess", (s.get("hessian_apply_mf") or {}).get("ms"), (s.get("hessian_apply_mf") or {}).get("frac"), "gs", [round(g["ms"],3) for g in s.get("gs_smooth",[])], "substep", (s.get("hot_substep") or {}).get("steady_ms"))
    except Exception as e: print(f, "failed", e)
PY
tail -3 $O/r2w_bench_c5.err
