set -x
O=gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_dist.py -x -q -k "multigrid or peer" > $O/r2p_pytest.log 2>&1; tail -15 $O/r2p_pytest.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --workload c2 --scaling weak > $O/r2p_weak_c2_2.json 2> $O/r2p_weak_c2_2.err
tail -5 $O/r2p_weak_c2_2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_weak_c2_2.json"))
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"])
print(json.dumps(d["solver_kernels"], indent=0)[:3000])
PY
