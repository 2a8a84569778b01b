set -x
O=gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_matrix.py tests/test_gpu_solver.py -x -q > $O/r2m_pytest.log 2>&1; tail -5 $O/r2m_pytest.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --cpu-reps 0 > $O/r2m_bench.json 2> $O/r2m_bench.err; tail -3 $O/r2m_bench.err
HOT_ASSEMBLE=scatter81 timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --cpu-reps 0 > $O/r2m_bench_s81.json 2> $O/r2m_bench_s81.err
python - <<'PY'
import json
for f in ("r2m_bench","r2m_bench_s81"):
    d=json.load(open(f"gpurun_out/{f}.json")); s=d["solver_kernels"]
    print(f, round(d["value"]), "asm", s["build_matrix_ms"], "mg", s["build_mg_ms"], "vc", s["vcycle"]["ms"], "substep", s["hot_substep"]["steady_ms"])
PY
