"""Run the bench workload's detection step a few times eagerly (for ncu -k regex:<kernel> captures)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lvc_b200.modeling import DetectorEngine  # noqa: E402
from lvc_b200.weights import synthetic_state_dict  # noqa: E402

cfg = bench.bench_cfg()
eng = DetectorEngine(cfg, synthetic_state_dict(cfg, 0))
ims = bench.make_images(0, bench.BATCH, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = eng.run(ims)
torch.cuda.synchronize()
if eng.debug is None:
    eng.debug = {}
    eng.run(ims)
    torch.cuda.synchronize()
    c = eng.debug["prop_counts"]
    b = eng.debug["props"][0, : int(c[0])]
    wh = (b[:, 2:] - b[:, :2])
    print("proposal counts", c.tolist(), "img0 w/h quantiles", torch.quantile(wh[:, 0], torch.tensor([0.1, 0.5, 0.9], device="cuda")).tolist(),
          torch.quantile(wh[:, 1], torch.tensor([0.1, 0.5, 0.9], device="cuda")).tolist())
print("detections", out[4].tolist())
