#!/bin/bash
# miner tests + the pipeline bench section alone
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mining.py -x -q 2>&1 | tail -25
python - <<'PY' 2>&1 | tail -20
import json, torch, bench
from lvc_b200.modeling import GeneralizedRCNN
from lvc_b200.weights import synthetic_state_dict
cfg = bench.bench_cfg(); sd = synthetic_state_dict(cfg, 0)
model = GeneralizedRCNN(cfg, sd, "cuda", use_cuda_graph=True)
r = bench.pipeline_section(model, sd, torch.device("cuda"))
print(json.dumps(r, indent=1))
json.dump(r, open("gpurun_out/r02_pipeline.json", "w"), indent=1)
PY
