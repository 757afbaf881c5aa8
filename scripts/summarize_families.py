"""Per-kernel-family summary of an `ncu --csv --page raw` capture (SpeedOfLight + MemoryWorkloadAnalysis sections):
launches, total time, DRAM bytes, achieved DRAM GB/s against the measured HBM peak, SM / L2 throughput, registers.

usage: python scripts/summarize_families.py capture.csv "header" > profiles/r02_families_<what>_summary.txt"""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, hdr = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6455.0
lines = [l for l in open(path, newline="") if l.startswith('"')]
rd = csv.reader(lines)
head = next(rd)
units = next(rd)
ix = {n: i for i, n in enumerate(head)}


def val(r, name, default=0.0):
    i = ix.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3,
             "second": 1e6}.get(u, 1.0)
    return v * scale


agg = defaultdict(lambda: dict(n=0, us=0.0, rd=0.0, wr=0.0, sm=0.0, l2=0.0, regs=0, occ=0.0, tensor=0.0))
for r in rd:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("mk::", "")
    a = agg[name]
    us = val(r, "gpu__time_duration.sum")
    a["n"] += 1
    a["us"] += us
    if "dram__bytes_read.sum" in ix:
        a["rd"] += val(r, "dram__bytes_read.sum")
        a["wr"] += val(r, "dram__bytes_write.sum")
    else:  # section captures carry the rate only: bytes = rate x duration
        a["rd"] += val(r, "dram__bytes.sum.per_second") * us * 1e-6
    a["sm"] += us * val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    a["l2"] += us * val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    a["occ"] += us * val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["regs"] = max(a["regs"], int(val(r, "launch__registers_per_thread")))
tot = sum(a["us"] for a in agg.values())
if hdr:
    print(hdr)
print(f"HBM peak (MEASURED_PEAKS.json): {peak:.0f} GB/s.  Per-launch times under ncu are cold-cache and serialised: compare shares and GB/s, not absolutes.")
print(f"{'kernel':34s} {'launches':>8s} {'time us':>10s} {'share':>6s} {'DRAM MB':>10s} {'GB/s':>7s} {'of HBM':>7s} {'SM %':>6s} {'L2 %':>6s} {'warps %':>8s} {'regs':>5s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] > 0 else 0.0
    print(f"{name[:34]:34s} {a['n']:8d} {a['us']:10.1f} {100 * a['us'] / tot:5.1f}% {(a['rd'] + a['wr']) / 1e6:10.1f} {gbs:7.0f} {gbs / peak:7.2f} "
          f"{a['sm'] / max(a['us'], 1e-9):6.1f} {a['l2'] / max(a['us'], 1e-9):6.1f} {a['occ'] / max(a['us'], 1e-9):8.1f} {a['regs']:5d}")
