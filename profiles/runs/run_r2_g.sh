set -x
timeout -s KILL 1200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -30
