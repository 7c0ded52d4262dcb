#!/bin/bash
# final launch list of one eager detection step (ncu durations + DRAM bytes), summarised per kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras --no-cpu-baseline > gpurun_out/r02n_launch_bench.log 2>&1
python - <<'PY'
import csv, re, collections
lines = [l for l in open("gpurun_out/r02n_launches.csv") if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    k = int(r["ID"])
    per.setdefault(k, {"name": re.sub(r"\(.*", "", r["Kernel Name"])})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
ids = sorted(per)
# the last complete forward: from the last stem kernel on
last = max(i for i in ids if "stem_s2d4" in per[i]["name"])
sel = [per[i] for i in ids if i >= last]
tot = sum(p["gpu__time_duration.sum"] for p in sel)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for p in sel:
    a = agg[p["name"]]; a[0] += 1; a[1] += p["gpu__time_duration.sum"]; a[2] += p.get("dram__bytes_read.sum", 0) + p.get("dram__bytes_write.sum", 0)
out = [f"({len(sel)} launches, {tot / 1e3:.3f} ms summed -- cold-cache and serialised under ncu: compare shares)\n", "| kernel | launches | total us | share | DRAM MB |", "|---|---:|---:|---:|---:|"]
for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% | {b / 1e6:.1f} |")
open("gpurun_out/r02n_launches.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
PY
