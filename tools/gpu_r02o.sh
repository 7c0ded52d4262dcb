#!/bin/bash
# two-stage separable pooler: tests, then op bench (COCO-shaped boxes) and step time per row threshold (0 = per-bin-row kernel only)
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -x -q -k "roi_pool or level or golden_e2e or stagewise or corrector or deterministic" -p no:cacheprovider 2>&1 | tail -3
for f in 0 4 6 8 12 24; do
  LVCB200_POOL_SEP2=$f python -c "
import json, torch, bench
r = bench.ops_section(torch.device('cuda'))
print('SEP2=$f op ms', round(r['roi_pool_fpn']['ms'], 4))
"
  LVCB200_POOL_SEP2=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SEP2=$f step', round(d['ms_per_step'], 4), round(d['value'], 1))"
done
