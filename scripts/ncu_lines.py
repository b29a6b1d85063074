"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line.

usage: python scripts/ncu_lines.py export.csv [top_n]
Prints, for the lines with most stall samples: samples, share, instructions executed and the
dominant stall reasons.
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
col = {name: i for i, name in enumerate(hdr)}
src_i = 1
samp_i = col["# Samples"]
inst_i = col["Instructions Executed"]
stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(lambda: {"samples": 0, "inst": 0, "src": "", "stalls": defaultdict(int)})
for r in rows[hdr_i + 1:]:
    if len(r) <= samp_i or not r[0].strip().isdigit():
        continue
    a = agg[int(r[0])]
    a["src"] = r[src_i].strip()
    try:
        a["samples"] += int(r[samp_i] or 0)
        a["inst"] += int(r[inst_i] or 0)
    except ValueError:
        continue
    for n, i in stall_cols:
        try:
            a["stalls"][n] += int(r[i] or 0)
        except ValueError:
            pass
tot = sum(a["samples"] for a in agg.values()) or 1
tot_i = sum(a["inst"] for a in agg.values()) or 1
print(f"total samples {tot}, instructions {tot_i}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    st_s = " ".join(f"{n[6:]}={v}" for n, v in st if v)
    print(f"{ln:5d} {100 * a['samples'] / tot:5.1f}% inst {100 * a['inst'] / tot_i:5.1f}%  {st_s:45s} | {a['src'][:70]}")
