#!/bin/bash
HOT_GX_COOP=3 timeout -s KILL 150 python -m pytest tests/test_gpu_matrix.py -m gpu -x -q -k "smoother_parity or vcycle_parity" 2>&1 | tail -2
timeout -s KILL 150 python -m pytest tests/test_gpu_matrix.py tests/test_gpu_fullsize.py -m gpu -x -q -k "smoother_parity or vcycle" 2>&1 | tail -2
HOT_GX_COOP0=0 timeout -s KILL 90 python profiles/prof_smooth.py 2>&1 | tail -1
HOT_GX_COOP0=1 timeout -s KILL 90 python profiles/prof_smooth.py 2>&1 | tail -1
HOT_GX_COOP=3 timeout -s KILL 90 python profiles/prof_smooth.py 2>&1 | tail -1
