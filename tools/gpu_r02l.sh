#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -x -q -k "rpn or golden_e2e or stagewise" -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rpn_" -c 40 --csv --log-file gpurun_out/r02l_rpn.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-graph > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/r02l_rpn.csv") if l.startswith('"')))
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ik].split("(")[0]].append(float(r[iv].replace(",", "")))
for k, v in d.items():
    print(f"{k:40s} n={len(v):3d} mean {sum(v)/len(v)/1e3:8.1f} us")
PY
