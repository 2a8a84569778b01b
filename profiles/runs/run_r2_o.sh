set -x
O=gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_dist.py -x -q -k "multigrid" > $O/r2o_pytest.log 2>&1; tail -30 $O/r2o_pytest.log
