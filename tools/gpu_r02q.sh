#!/bin/bash
# ncu metrics over the dense launches of one step (final engine) -> gpurun_out/r02q_gemm_step_metrics.json
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm --csv --log-file gpurun_out/r02q_gemm.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras --no-cpu-baseline > gpurun_out/r02q_bench.log 2>&1
python - <<'PY'
import csv
lines = [l for l in open("gpurun_out/r02q_gemm.csv") if l.startswith('"')]
rows = list(csv.DictReader(lines))
ids = sorted({int(r["ID"]) for r in rows})
keep = set(ids[-35:])
with open("gpurun_out/r02q_gemm_last35.csv", "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=rows[0].keys(), quoting=csv.QUOTE_ALL)
    w.writeheader()
    for r in rows:
        if int(r["ID"]) in keep:
            w.writerow(r)
print(len(ids), "gemm launches captured; kept the last 35")
PY
python tools/ncu_step_metrics.py gpurun_out/r02q_gemm_last35.csv gpurun_out/r02q_gemm_step_metrics.json "tools/gpu_r02q.sh: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm python bench.py --steps 1 --warmup 3 --no-graph --no-extras --no-cpu-baseline; the 35 dense launches of the last replay of bench.py's dense-only CUDA graph (final engine of round 2)"
python -c "
import json; d=json.load(open('gpurun_out/r02q_gemm_step_metrics.json')); print({k:v for k,v in d.items() if k not in ('by_kernel','source')})"
