set -x
O=gpurun_out
nvidia-smi topo -m > $O/r2k_topo.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_dist.py -x -q -k peer 2>&1 | tail -15
run() { # tag n workload scaling env
  env $5 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $2 --steps 20 --warmup 5 --workload $3 --scaling $4 --no-solver > $O/r2k_$1.json 2> $O/r2k_$1.err
  tail -3 $O/r2k_$1.err
}
run weak_c2_2_peer 2 c2 weak HOT_XCHG=peer
run weak_c2_2_nccl 2 c2 weak HOT_XCHG=nccl
run strong_c4_2_peer 2 c4 strong HOT_XCHG=peer
run strong_c4_2_nccl 2 c4 strong HOT_XCHG=nccl
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k_*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), {k:round(v["ms"],4) for k,v in d["roofline"]["per_kernel"].items()}, d["config"]["parallelism"][-60:])
    except Exception as e: print(f,"failed",e)
PY
