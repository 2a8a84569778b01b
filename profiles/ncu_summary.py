"""Summarise an .ncu-rep (raw page) into a compact per-launch table: python profiles/ncu_summary.py rep [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
def conv(v, u, name):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if "time_duration" in name:
        return f"{x / 1e3 if u in ('ns', 'nsecond') else (x if u in ('us','usecond') else x * 1e3):.1f}"
    if "bytes" in name:
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
        return f"{x * scale:.1f}"
    return f"{x:.1f}" if x != int(x) else str(int(x))
lines = ["| # | kernel | " + " | ".join(w[1] for w in want) + " |", "|" + "---|" * (len(want) + 2)]
ik = col("Kernel Name")
for n, r in enumerate(rows[2:]):
    name = r[ik].split("(")[0].replace("hot::<unnamed>::", "").replace("<unnamed>::", "")[:60]
    vals = []
    for w, _ in want:
        c = col(w)
        vals.append(conv(r[c], units[c], w) if c is not None else "-")
    lines.append(f"| {n} | {name} | " + " | ".join(vals) + " |")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
