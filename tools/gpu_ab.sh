#!/bin/bash
# same-box A/B of two builds of the library: A = lvc_b200/liblvcb200.so, B = $1 (LVCB200_LIB); three alternating short bench runs each
B=$1
for i in 1 2 3; do
  for lib in "" "$B"; do
    LVCB200_LIB=$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lib=${lib:-default}', round(d['ms_per_step'], 4), round(d['roofline']['gemm_ms_per_step'], 4))"
  done
done
