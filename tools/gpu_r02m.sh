#!/bin/bash
# launch list of the ViT front end (1024 crops), summed per kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m_vit.csv python tools/vit_prof.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02m_vit.csv") if l.startswith('"')))
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik][:90]].append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in d.values())
print("total ms (all 5 forwards incl. warm-up):", tot / 1e6)
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:92s} n={len(v):4d} sum {sum(v)/1e6:8.2f} ms  mean {sum(v)/len(v)/1e3:8.1f} us  {100*sum(v)/tot:5.1f}%")
PY
