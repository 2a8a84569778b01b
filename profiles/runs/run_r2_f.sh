set -x
timeout -s KILL 900 python -m pytest tests/test_oracle_ziran_ref.py tests/test_plugin.py tests/test_gpu_host_cpp.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -25
HOT_HOST_COLLIDERS=1 timeout -s KILL 600 python -m pytest tests/test_gpu_host_cpp.py -m gpu -x -q -k time_steps 2>&1 | tail -3
