#!/bin/bash
# programmatic dependent launch of the GS colour phases: parity (per-phase path forced on the small scenes), full-size V-cycle parity, timing
HOT_GS_COOP=0 timeout -s KILL 150 python -m pytest tests/test_gpu_matrix.py tests/test_gpu_solver.py -m gpu -x -q 2>&1 | tail -1
timeout -s KILL 200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -1
HOT_GX_PDL=0 timeout -s KILL 90 python profiles/prof_smooth.py 2>&1 | tail -1
HOT_GX_PDL=1 timeout -s KILL 90 python profiles/prof_smooth.py 2>&1 | tail -1
