#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_engine.py tests/test_gpu_dense.py -q -p no:cacheprovider -k "shape_lru or lateral_conv" > gpurun_out/r02c_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/r02c_memcheck.log
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-table gpurun_out/r02c_gemm_table.txt --mining-images 2000 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?" >> gpurun_out/r02c_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-extras --no-cpu-baseline > gpurun_out/r02c_launch_bench.log 2>&1
tail -3 gpurun_out/r02c_pytest.log gpurun_out/r02c_memcheck.log
