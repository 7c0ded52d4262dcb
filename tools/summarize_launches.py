"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of the captured window).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((int(r["ID"]), name, v))
rows = rows[skip:]
tot = sum(v for _, _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for _, n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
print(f"# launch list summary: {len(rows)} launches, {tot / 1e3:.3f} ms total (ncu per-launch times are cold-cache and serialised: compare shares)\n")
print("| kernel | launches | total us | share |")
print("|---|---:|---:|---:|")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n[:90]}` | {c} | {v:.1f} | {100 * v / tot:.1f}% |")
