set -x
O=gpurun_out
nproc
timeout -s KILL 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --durations=10 2>&1 | tail -25
