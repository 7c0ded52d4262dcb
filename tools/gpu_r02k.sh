#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 2
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_vit.py -x -q -k "attention_kernel or attention_tc" -p no:cacheprovider > gpurun_out/r02_san_attn.log 2>&1
echo "rc=$?" >> gpurun_out/r02_san_attn.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_mining.py -x -q -k "subsample or miner_equals_stagewise_composition or no_candidates" -p no:cacheprovider > gpurun_out/r02_san_misc.log 2>&1
echo "rc=$?" >> gpurun_out/r02_san_misc.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_vit.py -x -q -k "attention_kernel_vs_torch and tc and 257" -p no:cacheprovider > gpurun_out/r02_san_attn_race.log 2>&1
echo "rc=$?" >> gpurun_out/r02_san_attn_race.log
for f in gpurun_out/r02_san_attn.log gpurun_out/r02_san_misc.log gpurun_out/r02_san_attn_race.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $f | tail -5; done
