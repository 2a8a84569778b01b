set -x
O=gpurun_out
nvidia-smi topo -m > $O/r2l_topo.txt 2>&1
nproc > $O/r2l_nproc.txt
run() { # tag n workload scaling env...
  tag=$1; n=$2; wl=$3; sc=$4; shift 4
  env HOT_BENCH_VERBOSE=1 "$@" timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $n --steps 20 --warmup 5 --workload $wl --scaling $sc --no-solver > $O/r2l_$tag.json 2> $O/r2l_$tag.err
  grep "^\[rank" $O/r2l_$tag.err | cut -c1-400
}
run weak_c2_8_peer 8 c2 weak HOT_XCHG=peer
run weak_c2_8_nccl 8 c2 weak HOT_XCHG=nccl NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P
run weak_c2_8_peer_bar 8 c2 weak HOT_XCHG=peer HOT_BENCH_STEP_BARRIER=1
run strong_c4_8_peer 8 c4 strong HOT_XCHG=peer
grep -i " via \|channel" $O/r2l_weak_c2_8_nccl.err | head -30 > $O/r2l_nccl_via.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d=json.load(open(f))
        print(f, round(d["value"]), round(d["ms_per_step"],4), round(d["e2e"]["value"],1), {k:round(v["ms"],4) for k,v in d["roofline"]["per_kernel"].items()}, d["wall_s_timed_loop"])
    except Exception as e: print(f,"failed",e)
PY
