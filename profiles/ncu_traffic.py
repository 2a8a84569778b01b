"""profiles/r1_traffic.json from an `ncu --set full` report: DRAM bytes and duration per launch of the transfer kernels
   python profiles/ncu_traffic.py rep.ncu-rep profiles/r1_traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
def val(r, name):
    c = hdr.index(name)
    x = float(r[c].replace(",", ""))
    u = units[c]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    return x * scale
keys = {"p2g": "P2GPolicy", "g2p": "k_g2p", "force_scatter": "ForcePolicy", "hessian_gather": "k_hessian_gather", "update_state": "k_update_state"}
res = {}
for k, pat in keys.items():
    sel = [r for r in rows[2:] if pat in r[ik]]
    if not sel: continue
    rd = sum(val(r, "dram__bytes_read.sum") for r in sel) / len(sel)
    wr = sum(val(r, "dram__bytes_write.sum") for r in sel) / len(sel)
    us = sum(val(r, "gpu__time_duration.sum") for r in sel) / len(sel)
    res[k] = {"kernel": sel[0][ik].split("(")[0][:90], "launches": len(sel), "dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "us_under_ncu": us}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
