#!/bin/bash
# 8-GPU confirmation after the bench fix of b577d74 (clock sampling off the timed ranks): weak C2 and strong C4, default transport
O=gpurun_out; mkdir -p $O
run() { # tag n workload scaling
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $2 --steps 20 --warmup 5 --workload $3 --scaling $4 --no-solver > $O/r2t_$1.json 2> $O/r2t_$1.err
}
run weak_c2_8 8 c2 weak
run strong_c4_8 8 c4 strong
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), {k:round(v["ms"],4) for k,v in d["roofline"]["per_kernel"].items()}, d["sort_ms"], d["clocks"])
    except Exception as e: print(f,"failed",e)
PY
