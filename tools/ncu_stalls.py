"""Summarises the per-instruction warp-stall samples of an `ncu --page source --csv` export.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print(f"kernel: {rows[0][1]}\ntotal samples: {tot}")
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stall_cols}
print("stall reasons (all samples):")
for s, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {s:28s} {v:8d}  {100.0 * v / max(tot, 1):5.1f}%")
print(f"top {top} instructions by samples:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][col['# Samples']] or 0))[:top]
for i in sorted(order):
    r = body[i]
    n = int(r[col["# Samples"]] or 0)
    why = sorted(((int(r[col[s]] or 0), s) for s in stall_cols), reverse=True)[:2]
    print(f"  #{i:4d} {100.0 * n / max(tot, 1):5.1f}%  {r[col['Source']][:60]:60s} " +
          ", ".join(f"{s[6:]}={v}" for v, s in why if v))
