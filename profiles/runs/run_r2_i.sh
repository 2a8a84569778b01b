set -x
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
