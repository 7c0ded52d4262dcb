#!/bin/bash
# GPU call A of round 2: full -m gpu suite, bench, ncu --set full of the HBM-class kernels, compute-sanitizer on the dense tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 420 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 500 ncu --set full --clock-control none -k regex:'roi_pool|rpn_|det_|knn_|nms_' -s 16 -c 16 -o /tmp/r02a_ops python tools/ops_prof.py > gpurun_out/r02a_ncu_ops.log 2>&1
ncu -i /tmp/r02a_ops.ncu-rep --page raw --csv > gpurun_out/r02a_ops_raw.csv 2>> gpurun_out/r02a_ncu_ops.log
ls -la /tmp/r02a_ops.ncu-rep >> gpurun_out/r02a_ncu_ops.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_dense.py -q -p no:cacheprovider -k "chain or conv3x3 or split_conv or residual" > gpurun_out/r02a_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02a_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_dense.py -q -p no:cacheprovider -k "test_layer_chain_matches_per_layer_launches and 25-42" > gpurun_out/r02a_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02a_racecheck.log
tail -5 gpurun_out/r02a_pytest.log
