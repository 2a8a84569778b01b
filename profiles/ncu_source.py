"""Aggregate the source page of one launch of an .ncu-rep by CUDA source line:
   python profiles/ncu_source.py rep launch_index [top]
prints samples, instructions and the dominant stall reasons per source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io, collections
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
print(rows[1][1][:150])
c_samp = hdr.index("# Samples"); c_inst = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.OrderedDict()
cur = None
seen = set()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    if r[0] == "":  # SASS row of the current source line: opcode statistics only
        if cur is not None and r[3] not in ("...", "-", "") and r[2] not in seen:
            seen.add(r[2])
            t = r[3].split()
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            try:
                cur[3][op.split(".")[0]] += int(r[c_inst] or 0)
                cur[4] += int(r[c_inst + 1] or 0)
            except ValueError: pass
        continue
    key = (r[0], r[1].strip()[:110])
    cur = a = agg.setdefault(key, [0, 0, collections.Counter(), collections.Counter(), 0])
    try:
        a[0] += int(r[c_samp] or 0); a[1] += int(r[c_inst] or 0)
    except ValueError:
        continue
    for i in stalls:
        try: a[2][hdr[i][6:]] += int(r[i] or 0)
        except ValueError: pass
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"total samples {tot}, warp instructions {toti}")
allst = collections.Counter()
for a in agg.values(): allst.update(a[2])
print("stalls overall:", ", ".join(f"{k} {v * 100 // max(1, sum(allst.values()))}%" for k, v in allst.most_common(8)))
ops = collections.Counter()
for a in agg.values(): ops.update(a[3])
print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in a[2].most_common(3))
    print(f"{a[0] * 100 / max(tot, 1):5.1f}% smp {a[1] * 100 / max(toti, 1):5.1f}% ins  L{ln:>4} {src[:90]:90s} | {st}")
