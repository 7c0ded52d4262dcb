#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k "knn" > gpurun_out/r02f_knn.log 2>&1
echo "rc=$?" >> gpurun_out/r02f_knn.log
python tools/knn_prof.py > gpurun_out/r02f_knn_prof.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_bytes.sum --clock-control none -k regex:'knn_' -c 8 --csv --log-file gpurun_out/r02f_knn_ncu.csv python tools/knn_prof.py > gpurun_out/r02f_knn_ncu.log 2>&1
tail -3 gpurun_out/r02f_knn.log; cat gpurun_out/r02f_knn_prof.log
