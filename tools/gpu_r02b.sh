#!/bin/bash
mkdir -p gpurun_out
run() { # name, -k expr
  CUDA_LAUNCH_BLOCKING=1 timeout 400 python -m pytest tests/test_gpu_engine.py -q -s -p no:cacheprovider -k "$2" > gpurun_out/r02b_$1.log 2>&1
  echo "rc=$?" >> gpurun_out/r02b_$1.log
}
run strict_oracle "strict_engine_vs_fp32_oracle"
run gold_cos_strict "reference_golden_e2e and cosine and strict"
run gold_cos_bf16 "reference_golden_e2e and cosine and bf16"
run gold_b8_strict "reference_golden_e2e and b8 and strict"
run gold_b8_bf16 "reference_golden_e2e and b8 and bf16"
run rest_engine "not strict_engine_vs_fp32_oracle and not reference_golden_e2e"
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_engine.py --ignore=tests/test_gpu_dense.py > gpurun_out/r02b_rest.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_rest.log
# the in-sequence crash of call A: golden e2e tests in one process, under memcheck (small image first)
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_engine.py -q -p no:cacheprovider -k "reference_golden_e2e and cosine" > gpurun_out/r02b_memcheck_seq.log 2>&1
echo "rc=$?" >> gpurun_out/r02b_memcheck_seq.log
tail -3 gpurun_out/r02b_*.log
