set -x
O=gpurun_out
run() { # n workload scaling extra
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 20 --warmup 5 --workload $2 --scaling $3 $4 > $O/r2j_$3_$2_$1gpu.json 2> $O/r2j_$3_$2_$1gpu.err
  tail -2 $O/r2j_$3_$2_$1gpu.err
}
run 8 c2 weak ""
run 4 c2 weak "--no-solver"
run 8 c4 strong "--no-solver"
run 4 c4 strong "--no-solver"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j_*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), {k:round(v["ms"],4) for k,v in d["roofline"]["per_kernel"].items()}, d.get("sort_ms"))
        if d.get("solver_kernels"): print("  solver", json.dumps(d["solver_kernels"])[:600])
    except Exception as e: print(f,"failed",e)
PY
