#!/bin/bash
# two-stage separable pooler: tests, op bench (both kernels), step time
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_engine.py -x -q -k "roi_pool or level or golden_e2e or stagewise or corrector or deterministic" -p no:cacheprovider 2>&1 | tail -5
python - <<'PY'
import json, os, torch, bench
for flag in ("1", "0"):
    os.environ["LVCB200_POOL_SEP2"] = flag
PY
for f in 1 0; do LVCB200_POOL_SEP2=$f python -c "
import json, torch, bench
r = bench.ops_section(torch.device('cuda'))
print('SEP2=$f', json.dumps(r.get('roi_pool_fpn')))
"; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
