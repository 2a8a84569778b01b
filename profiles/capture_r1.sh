#!/bin/bash
# Round-1 evidence capture (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are copied to profiles/ afterwards)
#   gpurun --timeout 1500 -- 'bash profiles/capture_r1.sh'
set -x
mkdir -p gpurun_out
(time timeout -s KILL 900 python -m pytest tests -m gpu -q) > gpurun_out/pytest_gpu.log 2>&1
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/r1_bench_c2_reference.json 2> gpurun_out/bench_ref.err
timeout -s KILL 600 python bench.py > gpurun_out/r1_bench_c2.json 2> gpurun_out/bench.err
# launch list of the same command (shares, not absolutes: cold-cache, serialised launches)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --cpu-reps 0 > gpurun_out/b_launches.log 2>&1
# full capture of the transfer / force kernels (source-level stall sampling needs -lineinfo + --import-source)
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_plane2_scatter|k_column_scatter|k_g2p|k_hessian_gather|k_update_state|k_page_masks|k_number_and_normalise|k_tile_dof" -c 24 \
    -o gpurun_out/r1_full_transfer -f python bench.py --steps 1 --warmup 3 --cpu-reps 0 > gpurun_out/b_full1.log 2>&1
# full capture of the solver-side kernels: one V-cycle's worth of GS phases, SpMV, update
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_gs_block|k_gs_sweep|k_spmv|k_restrict|k_prolong|k_assemble|k_galerkin" -c 44 \
    -o gpurun_out/r1_full_solver -f python profiles/prof_gs.py > gpurun_out/b_full2.log 2>&1
ls -la gpurun_out
