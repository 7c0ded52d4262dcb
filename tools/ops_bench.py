"""Op-level timing of the fused RoI pooler on COCO-shaped and on the detector's own (sliver) proposals."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

out = bench.ops_section(torch.device("cuda"))
for k, v in out.items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "hbm_gbs", "frac_of_measured_hbm")})
