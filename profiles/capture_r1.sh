#!/bin/bash
# Round-1 evidence capture (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are copied to profiles/ afterwards)
#   gpurun --timeout 1800 -- 'bash profiles/capture_r1.sh'
# The .ncu-rep files are summarised ON the box (gpurun merges at most 64 MiB back) and then deleted.
set -x
O=gpurun_out
mkdir -p $O
(time timeout -s KILL 900 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1
timeout -s KILL 600 python bench.py --impl reference > $O/r1_bench_c2_reference.json 2> $O/bench_ref.err
timeout -s KILL 600 python bench.py > $O/r1_bench_c2.json 2> $O/bench.err
# launch list of the same command (shares, not absolutes: cold-cache, serialised launches)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r1_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --cpu-reps 0 > $O/b_launches.log 2>&1
# full capture of the transfer / force kernels (source-level stall sampling needs -lineinfo + --import-source)
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_plane2_scatter|k_column_scatter|k_g2p|k_hessian_gather|k_update_state|k_page_masks|k_number_and_normalise|k_tile_dof" -c 16 \
    -o $O/r1_full_transfer -f python bench.py --steps 1 --warmup 3 --cpu-reps 0 > $O/b_full1.log 2>&1
python profiles/ncu_summary.py $O/r1_full_transfer.ncu-rep $O/r1_ncu_full_transfer.md > /dev/null
python profiles/ncu_traffic.py $O/r1_full_transfer.ncu-rep $O/r1_traffic.json > /dev/null
python - <<'PY'
import csv, io, subprocess
rep = "gpurun_out/r1_full_transfer.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ik = rows[0].index("Kernel Name")
seen = {}
for n, r in enumerate(rows[2:]):
    key = r[ik].split("(")[0][-60:]
    if key not in seen:
        seen[key] = n
for key, n in seen.items():
    tag = "".join(ch if ch.isalnum() else "_" for ch in key)[-40:]
    out = subprocess.run(["python", "profiles/ncu_source.py", rep, str(n), "30"], capture_output=True, text=True).stdout
    open(f"gpurun_out/r1_ncu_source_{tag}.txt", "w").write(out)
PY
rm -f $O/r1_full_transfer.ncu-rep
# full capture of the solver-side kernels: one V-cycle's worth of GS phases, SpMV, update
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_gs_block|k_gs_sweep|k_gs_stream_update|k_spmv|k_assemble|k_galerkin" -c 30 \
    -o $O/r1_full_solver -f python profiles/prof_gs.py > $O/b_full2.log 2>&1
python profiles/ncu_summary.py $O/r1_full_solver.ncu-rep $O/r1_ncu_full_solver.md > /dev/null
python profiles/ncu_source.py $O/r1_full_solver.ncu-rep 4 30 > $O/r1_ncu_source_gs_block.txt
rm -f $O/r1_full_solver.ncu-rep
HOT_GS_DEBUG=5 python profiles/prof_gs.py > $O/r1_gs_phase_stamps.txt 2>&1
HOT_CS_DEBUG=2 HOT_SCATTER=column python bench.py --cpu-reps 0 --no-solver --steps 3 > /dev/null 2> $O/r1_scatter_phase_stamps.txt
ls -la $O
