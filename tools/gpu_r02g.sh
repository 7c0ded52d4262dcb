#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02g_smoke.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc=$?" >> gpurun_out/r02g_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02g_step_metrics.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras --no-cpu-baseline > gpurun_out/r02g_step_metrics.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:'roi_pool|rpn_|det_|knn_|nms_' -s 18 -c 18 -o /tmp/r02g_ops python tools/ops_prof.py > gpurun_out/r02g_ncu_ops.log 2>&1
ncu -i /tmp/r02g_ops.ncu-rep --page raw --csv > gpurun_out/r02g_ops_raw.csv 2>> gpurun_out/r02g_ncu_ops.log
tail -3 gpurun_out/r02g_pytest.log gpurun_out/r02g_smoke.log
