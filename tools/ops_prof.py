"""One or two launches of every HBM / latency-class kernel of the path on the bench's op-level workloads, for `ncu --set full`
(profiles/r02_ops_ncu.md).  usage (on the GPU box): ncu --set full --clock-control none -k regex:'roi_pool|rpn_|det_|knn_|nms_' ... python tools/ops_prof.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from lvc_b200 import ops  # noqa: E402
from lvc_b200.testing import coco_like_boxes  # noqa: E402

dev = torch.device("cuda")
rng = np.random.default_rng(0)
N, P = 8, 1000
g = torch.Generator(device=dev).manual_seed(0)
sizes = [(200, 336), (100, 168), (50, 84), (25, 42)]
planes = []
for h, w in sizes:
    t = torch.zeros((N, h + 2, w + 2, 256), dtype=torch.bfloat16, device=dev)
    t[:, 1:h + 1, 1:w + 1] = torch.randn((N, h, w, 256), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    planes.append(ops.Plane(t, h, w, 256))
boxes = np.concatenate([np.concatenate([np.full((P, 1), i, np.float32), coco_like_boxes(rng, P)], 1) for i in range(N)])
rois = torch.from_numpy(boxes).to(dev)
scales = [0.25, 0.125, 0.0625, 0.03125]
lv = []
for (h, w) in sizes + [(13, 21)]:
    n = h * w * 3
    lv.append(ops.rpn_level_dense(torch.randn((N, n), generator=g, device=dev), torch.randn((N, n, 4), generator=g, device=dev) * 0.3, h, w, 3))
isz = torch.tensor([[800, 1333]] * N, dtype=torch.int32, device=dev)
logits = torch.randn((N * P, 81), generator=g, device=dev) * 2
deltas = torch.randn((N * P, 320), generator=g, device=dev) * 0.5
props = rois[:, 1:].contiguous()
rimg = torch.arange(N, device=dev, dtype=torch.int32).repeat_interleave(P)
S, D, Q, ncls = 600, 1024, 200_000, 20
means = torch.zeros(ncls, D, device=dev)
means[torch.arange(ncls), torch.arange(ncls)] = 4.0
cls_all = torch.arange(ncls, device=dev).repeat_interleave(S // ncls)
bank = torch.randn(S, D, generator=g, device=dev) + means[cls_all]
qcls = torch.randint(0, ncls, (Q,), generator=g, device=dev)
queries = torch.randn(Q, D, generator=g, device=dev) + means[qcls]
for rep in range(2):
    ops.roi_pool_fpn(planes, scales, rois, out_dtype=torch.bfloat16, out_layout=ops.OUT_NHWC)
    ops.rpn_proposals(lv, isz, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0))
    ops.detections(logits, deltas, props, rimg, isz, isz, 80, max_rois_per_image=P)
    kb = ops.KnnBank(bank, cls_all)
    kb.verify(queries, qcls, topk=10, knn=10, path="tc")
    torch.cuda.synchronize()
print("ok")
