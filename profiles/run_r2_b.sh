set -x
O=gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_transfer.py tests/test_gpu_force.py -m gpu -x -q 2>&1 | tail -3
for k in jitter poisson; do
HOT_WS_DEBUG=2 timeout -s KILL 300 python profiles/ws_debug.py $k 2>&1 | grep -E "ws dbg|us per launch|particles"
timeout -s KILL 300 python profiles/ws_debug.py $k 2>&1 | grep -E "us per launch"
done
