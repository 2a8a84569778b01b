#!/bin/bash
# Round-2 evidence capture (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are copied to profiles/ afterwards)
#   gpurun --timeout 2400 -- 'bash profiles/capture_r2.sh'
# The .ncu-rep files are summarised ON the box (gpurun merges at most 64 MiB back) and then deleted.
set -x
O=gpurun_out
mkdir -p $O
(time timeout -s KILL 1500 python -m pytest tests -m gpu -q) > $O/r2_pytest_gpu.log 2>&1
timeout -s KILL 600 python bench.py --impl reference > $O/r2_bench_c2_reference.json 2> $O/r2_bench_ref.err
timeout -s KILL 900 python bench.py > $O/r2_bench_c2.json 2> $O/r2_bench.err
timeout -s KILL 900 python bench.py --workload c4 --cpu-reps 0 > $O/r2_bench_c4.json 2> $O/r2_bench_c4.err
# launch list of the same command (shares, not absolutes: cold-cache, serialised launches)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --cpu-reps 0 > $O/r2_b_launches.log 2>&1
# full capture of the transfer / force kernels (source-level stall sampling needs -lineinfo + --import-source)
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_plane2_scatter|k_g2p|k_hessian_gather|k_update_state|k_page_masks|k_number_and_normalise|k_tile_dof" -c 14 \
    -o $O/r2_full_transfer -f python bench.py --steps 1 --warmup 3 --cpu-reps 0 > $O/r2_b_full1.log 2>&1
python profiles/ncu_summary.py $O/r2_full_transfer.ncu-rep $O/r2_ncu_full_transfer.md > /dev/null
python profiles/ncu_traffic.py $O/r2_full_transfer.ncu-rep $O/r2_traffic.json > /dev/null
python - <<'PY'
import csv, io, subprocess
rep = "gpurun_out/r2_full_transfer.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ik = rows[0].index("Kernel Name")
seen = {}
for n, r in enumerate(rows[2:]):
    key = r[ik].split("(")[0][-60:]
    if key not in seen:
        seen[key] = n
for key, n in seen.items():
    if not any(t in key for t in ("P2GPolicy", "k_g2p", "k_hessian_gather")):
        continue
    tag = "".join(ch if ch.isalnum() else "_" for ch in key)[-40:]
    out = subprocess.run(["python", "profiles/ncu_source.py", rep, str(n), "30"], capture_output=True, text=True).stdout
    open(f"gpurun_out/r2_ncu_source_{tag}.txt", "w").write(out)
PY
rm -f $O/r2_full_transfer.ncu-rep
# full capture of the solver-side kernels: assembly, hierarchy, one V-cycle's worth of GS phases, SpMV
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_gx_block|k_gx_sweep|k_gx_update|k_gx_inverse|k_gx_stream|k_spmv|k_assemble|k_mirror|k_galerkin" -c 45 \
    -o $O/r2_full_solver -f python profiles/prof_gs.py > $O/r2_b_full2.log 2>&1
python profiles/ncu_summary.py $O/r2_full_solver.ncu-rep $O/r2_ncu_full_solver.md > /dev/null
rm -f $O/r2_full_solver.ncu-rep
# SASS evidence: TMA bulk copies (UBLKCP), mbarrier waits (SYNCS), cp.async (LDGSTS), peer-memory release / acquire in the exchange kernels
cuobjdump -sass hot_b200/lib/libhot_b200.so > $O/all.sass 2>/dev/null
python - <<'PY'
import re, collections
fn = None; counts = collections.defaultdict(collections.Counter)
for line in open("gpurun_out/all.sass", errors="replace"):
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    for op in ("UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "RED.E", "ATOM", "ST.E.64.STRONG.SYS", "LD.E.64.STRONG.SYS", "DFMA", "DMMA", "HMMA", "UCGABAR", "MAPA"):
        if fn and (" " + op) in line: counts[fn][op] += 1
with open("gpurun_out/r2_sass_opcodes.md", "w") as f:
    f.write("| kernel (mangled, shortened) | opcode counts |\n|---|---|\n")
    for k, c in sorted(counts.items()):
        if any(t in k for t in ("k_g2p", "k_hessian_gather", "k_plane2_scatter", "k_scatter_ws", "k_assemble", "k_pack_peer", "k_unpack_shared", "k_gs_block", "k_gx_block", "k_gx_sweep", "k_gx_update", "k_spmv", "k_update_state")):
            f.write(f"| {k[-70:]} | " + ", ".join(f"{o} {n}" for o, n in sorted(c.items())) + " |\n")
PY
rm -f $O/all.sass
ls -la $O | tail -30
