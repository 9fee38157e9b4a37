"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total us, share.
usage: python tools/ncu_launch_summary.py launches.csv [--last-step MARKER] > profiles/rNN_ncu_launches_step.txt
--last-step adamw keeps only the launches of the last complete training step: those after the second-to-last launch whose
name contains MARKER, up to and including the last one (every step ends with exactly one AdamW launch)."""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = {h: i for i, h in enumerate(rows[hdr])}
body = [r for r in rows[hdr + 1:] if r[col["Metric Name"]] == "gpu__time_duration.sum"]
if "--last-step" in sys.argv:
    marker = sys.argv[sys.argv.index("--last-step") + 1]
    marks = [i for i, r in enumerate(body) if marker in r[col["Kernel Name"]]]
    if len(marks) >= 2:
        body = body[marks[-2] + 1:marks[-1] + 1]
agg = OrderedDict()
for r in body:
    name = r[col["Kernel Name"]]
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    val = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    us = val / 1e3 if unit in ("ns", "nsecond") else val * 1e3 if unit in ("ms", "msecond") else val
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + us)
total = sum(t for _, t in agg.values())
print(f"# {sum(n for n, _ in agg.values())} launches captured, {total / 1e3:.2f} ms total (cold-cache, serialised under ncu: compare SHARES)")
print(f"{'kernel':90s} {'launches':>8s} {'us':>10s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:90]:90s} {n:8d} {t:10.1f} {100 * t / total:6.1f}%")
