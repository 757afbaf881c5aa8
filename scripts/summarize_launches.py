"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel name.

usage: python scripts/summarize_launches.py launches.csv "header line(s)" > profiles/rNN_launches_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event numbers, not absolutes."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
hdr = sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
head = next(rd)
ix = {n: i for i, n in enumerate(head)}
agg = defaultdict(lambda: [0.0, 0])
for r in rd:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]])[:110]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    agg[name][0] += us
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
n = sum(v[1] for v in agg.values())
if hdr:
    print(hdr)
print(f"launches {n}, total {tot / 1e3:.2f} ms (cold-cache, serialised: compare shares)")
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{us:10.1f} us {100 * us / tot:5.1f}%  x{c:<5d} {name}")
