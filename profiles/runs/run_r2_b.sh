set -x
timeout -s KILL 600 python -m pytest tests/test_gpu_transfer.py tests/test_gpu_force.py tests/test_gpu_solver.py -m gpu -x -q 2>&1 | tail -3
for k in jitter poisson; do
timeout -s KILL 300 python profiles/ws_debug.py $k 2>&1 | grep -E "us per launch"
HOT_PF_DIST=0 timeout -s KILL 300 python profiles/ws_debug.py $k 2>&1 | grep -E "us per launch"
HOT_SCATTER=plane timeout -s KILL 300 python profiles/ws_debug.py $k 2>&1 | grep -E "us per launch"
done
