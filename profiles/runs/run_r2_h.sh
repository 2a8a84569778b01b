set -x
O=gpurun_out
nvidia-smi -L
for mode in weak strong; do
  wl=c2; [ $mode = strong ] && wl=c4
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl --scaling $mode > $O/r2h_${mode}_2gpu.json 2> $O/r2h_${mode}_2gpu.err
  tail -5 $O/r2h_${mode}_2gpu.err
done
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 --workload c4 --no-solver --cpu-reps 0 > $O/r2h_c4_1gpu.json 2> $O/r2h_c4_1gpu.err
python - <<'PY'
import json
for n in ("weak_2gpu","strong_2gpu","c4_1gpu"):
    try:
        d=json.load(open(f"gpurun_out/r2h_{n}.json"))
        print(n, round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["roofline"]["per_kernel"], d["config"]["parallelism"], d.get("sort_ms"))
    except Exception as e: print(n,"failed",e)
PY
