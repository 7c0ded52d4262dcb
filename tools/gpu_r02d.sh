#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k "knn" > gpurun_out/r02d_knn.log 2>&1
echo "rc=$?" >> gpurun_out/r02d_knn.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 300 python - > gpurun_out/r02d_knn_bench.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
print(json.dumps(bench.knn_section(0, 1, torch.device("cuda"), None, with_cpu=False), indent=1))
PY
tail -4 gpurun_out/r02d_knn.log gpurun_out/r02d_pytest.log; tail -30 gpurun_out/r02d_knn_bench.log
