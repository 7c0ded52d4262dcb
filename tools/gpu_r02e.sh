#!/bin/bash
mkdir -p gpurun_out
python tools/knn_prof.py > gpurun_out/r02e_knn_prof.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_bytes.sum --clock-control none -k regex:'knn_|gemm_bf16' --csv --log-file gpurun_out/r02e_knn_ncu.csv python tools/knn_prof.py > gpurun_out/r02e_knn_ncu.log 2>&1
cat gpurun_out/r02e_knn_prof.log
