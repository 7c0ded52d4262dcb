#!/bin/bash
python -m pytest tests/test_gpu_ops.py -x -q -k "subsample" 2>&1 | tail -15
python - <<'PY'
import torch, numpy as np
from lvc_b200.modeling import subsample_labels_batched
lab = torch.from_numpy(np.random.default_rng(0).choice(np.array([-1,0,1],np.int8), size=(8,268569), p=[0.3,0.699,0.001])).cuda()
keys = torch.randint(0, 2**32, lab.shape, dtype=torch.int64, device="cuda").to(torch.uint32)
for _ in range(3): subsample_labels_batched(lab, 256, 0.5, 0, keys=keys)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): subsample_labels_batched(lab, 256, 0.5, 0, keys=keys)
e1.record(); torch.cuda.synchronize()
print("subsample_labels 8 x 268569 anchors: %.1f us per call" % (e0.elapsed_time(e1)/20*1e3))
PY
