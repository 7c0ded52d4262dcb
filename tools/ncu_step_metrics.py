"""Summarise an ncu CSV (--page raw or --metrics list, --csv) of the dense kernels of one detection step into the JSON bench.py reads
for roofline.traffic.  usage: python tools/ncu_step_metrics.py gpurun_out/gemm_step.csv profiles/rNN_gemm_step_metrics.json "<source cmd>" """
import csv
import json
import sys
from collections import defaultdict

path, out, source = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
per = defaultdict(dict)
names = {}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    per[int(r["ID"])][r["Metric Name"]] = v * scale
    names[int(r["ID"])] = r["Kernel Name"].split("(")[0]
ids = sorted(per)
tot_us = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in ids)
dram = sum(per[i].get("dram__bytes_read.sum", 0.0) + per[i].get("dram__bytes_write.sum", 0.0) for i in ids)
l2 = sum(per[i].get("lts__t_bytes.sum", 0.0) for i in ids)
tp_key = next((k for k in per[ids[0]] if k.startswith("sm__pipe_tensor")), None)
tp = sum(per[i].get(tp_key, 0.0) * per[i].get("gpu__time_duration.sum", 0.0) for i in ids) / tot_us if tp_key else None
by = defaultdict(lambda: [0, 0.0, 0.0])
for i in ids:
    b = by[names[i]]
    b[0] += 1; b[1] += per[i].get("gpu__time_duration.sum", 0.0)
    b[2] += per[i].get("dram__bytes_read.sum", 0.0) + per[i].get("dram__bytes_write.sum", 0.0)
res = {"launches": len(ids), "total_us": tot_us, "dram_bytes_per_step": dram, "dram_bytes_per_launch": dram / len(ids),
       "l2_bytes_per_step": l2, "tensor_pipe_active_pct_time_weighted": tp,
       "by_kernel": {k: {"launches": v[0], "us": v[1], "dram_bytes": v[2]} for k, v in by.items()}, "source": source}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
