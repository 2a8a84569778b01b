#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"k_gx_sweep|k_gx_block" -c 40 -o $O/r2v_gx -f python profiles/prof_smooth.py > $O/r2v_ncu.log 2>&1
python profiles/ncu_summary.py $O/r2v_gx.ncu-rep $O/r2v_ncu_gx.md > /dev/null
python - <<'PY'
import csv, io, subprocess
rep = "gpurun_out/r2v_gx.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ik = rows[0].index("Kernel Name")
names = [r[ik] for r in rows[2:]]
# the slowest k_gx_block launch and the first k_gx_sweep launch
it = rows[0].index("gpu__time_duration.sum")
best = max((n for n, r in enumerate(rows[2:]) if "k_gx_block" in r[ik]), key=lambda n: float(rows[2 + n][it].replace(",", "")), default=None)
sweep = next((n for n, r in enumerate(rows[2:]) if "k_gx_sweep" in r[ik]), None)
for tag, n in (("block", best), ("sweep", sweep)):
    if n is None: continue
    out = subprocess.run(["python", "profiles/ncu_source.py", rep, str(n), "40"], capture_output=True, text=True).stdout
    open(f"gpurun_out/r2v_ncu_source_gx_{tag}.txt", "w").write(out)
PY
rm -f $O/r2v_gx.ncu-rep
tail -5 $O/r2v_ncu.log
