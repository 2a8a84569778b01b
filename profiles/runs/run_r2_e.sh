set -x
O=gpurun_out
for mode in plane item; do
  if [ $mode = plane ]; then export HOT_SCATTER=plane; else unset HOT_SCATTER; fi
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_plane2_scatter|k_item_scatter" -s 3 -c 2 \
      -o $O/r2e_$mode -f python profiles/ws_debug.py poisson 1 > $O/r2e_ncu_$mode.log 2>&1
  python profiles/ncu_summary.py $O/r2e_$mode.ncu-rep $O/r2e_ncu_$mode.md > /dev/null
  python profiles/ncu_source.py $O/r2e_$mode.ncu-rep 0 45 > $O/r2e_ncu_source_$mode.txt
  ncu -i $O/r2e_$mode.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY' > $O/r2e_raw_$mode.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ["smsp__warp_issue_stalled", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__inst_executed_pipe", "smsp__inst_executed.sum", "sm__warps_active", "smsp__issue_active", "sm__pipe_fp64", "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__average_warp", "smsp__warps_issue_stalled", "sm__cycles_elapsed.max", "smsp__pcsamp"]
for i, h in enumerate(hdr):
    if any(w in h for w in want):
        print(h, [r[i] for r in rows[2:4]])
PY
  rm -f $O/r2e_$mode.ncu-rep
done
