#!/bin/bash
# full round-end check on one B200: -m gpu suite, smoke, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/full_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/full_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
echo "bench rc=$?" >> gpurun_out/full_bench.err
tail -n 3 gpurun_out/full_pytest.log; tail -n 2 gpurun_out/full_smoke.log; tail -n 1 gpurun_out/full_bench.err
