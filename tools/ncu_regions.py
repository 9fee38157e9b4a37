"""Where do the warps of a kernel spend their time?  Buckets the SASS of an `ncu --page source --csv` export and prints,
per bucket with samples, the share of all warp samples, the executions per instruction, the MUFU count and the top stall
reasons; then the totals per opcode.
usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv [bucket=40] [min_samples=60]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
MIN = int(sys.argv[3]) if len(sys.argv) > 3 else 60
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
S = lambda r: int(r[col["# Samples"]] or 0)
E = lambda r: int(r[col["Instructions Executed"]] or 0)
tot = sum(S(r) for r in body)
print(f"kernel: {rows[0][1][:100]}\ninstructions: {len(body)}  samples: {tot}  warp-instructions executed: {sum(E(r) for r in body)}")
agg = collections.Counter()
for r in body:
    for s in stall:
        agg[s[6:]] += int(r[col[s]] or 0)
print("stalls:", ", ".join(f"{k}={100 * v / tot:.1f}%" for k, v in agg.most_common(9)))
for a in range(0, len(body), B):
    rs = body[a:a + B]
    n = sum(S(r) for r in rs)
    if n < MIN:
        continue
    c = collections.Counter()
    for r in rs:
        for s in stall:
            c[s[6:]] += int(r[col[s]] or 0)
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip()).split()[0].split(".")[0] for r in rs if r[col["Source"]].strip())
    print(f"  [{a:5d}] {100 * n / tot:5.1f}%  exec/instr {sum(E(r) for r in rs) // len(rs):8d}  " +
          ", ".join(f"{k}={v}" for k, v in c.most_common(3)) + "   | " + " ".join(f"{k}:{v}" for k, v in ops.most_common(4)))
ex = collections.Counter()
for r in body:
    src = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
    if src:
        ex[src.split()[0].split(".")[0]] += E(r)
t = sum(ex.values())
print("executed by opcode:", ", ".join(f"{k} {100 * v / t:.1f}%" for k, v in ex.most_common(14)))
