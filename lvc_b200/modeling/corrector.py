"""Box corrector ("Correct" stage): CascadeROIHeads._forward_box_qe with BoxOnlyLayersCascade heads
(lvc/modeling/roi_heads/cascade_rcnn.py:167-203, :329-369; roi_heads_cascade.py:134-138,197-211).

Per stage k: ROIPooler -> fc1..fc3 (+ReLU) -> Linear 1024->4 -> apply_deltas(stage weights) -> clip.  The final
fast_rcnn_inference of the reference runs on one-hot scores with nms_thresh 1.0 and topk 1e10 and restores the input order
(cascade_rcnn.py:190-203): it keeps every box, so the result is the clipped stage-3 boxes in input order with their classes.
"""
from typing import Dict, List

import torch

from .. import _lib, ops
from ..config import DetectorConfig

STRIDES = (4, 8, 16, 32)


class BoxCorrectorHead:
    def __init__(self, cfg: DetectorConfig, state_dict: Dict[str, torch.Tensor], device="cuda", num_fc=3, stages=3):
        _lib.load()
        self.cfg, self.device, self.stages = cfg, torch.device(device), stages
        res = cfg.pooler_resolution
        self.fc, self.pred = [], []
        for k in range(stages):
            layers = []
            for i in range(num_fc):
                w = state_dict[f"roi_heads.box_head.{k}.fc{i + 1}.weight"].float()
                if i == 0:   # c*49+h*7+w (box_head.py:86-87) -> pooler's (h, w, c)
                    w = w.view(-1, 256, res, res).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
                layers.append((w.contiguous().to(self.device, torch.bfloat16),
                               state_dict[f"roi_heads.box_head.{k}.fc{i + 1}.bias"].float().to(self.device)))
            self.fc.append(layers)
            wp = torch.zeros(16, cfg.fc_dim)
            bp = torch.zeros(16)
            wp[:4] = state_dict[f"roi_heads.box_predictor.{k}.bbox_pred.weight"].float()
            bp[:4] = state_dict[f"roi_heads.box_predictor.{k}.bbox_pred.bias"].float()
            self.pred.append((wp.to(self.device, torch.bfloat16), bp.to(self.device)))

    def head(self, k, pooled):
        """One stage's regression head on pooled features [R, 12544] bf16 (BASELINE config #5) -> deltas [R, 16] fp32 (4 used)."""
        x = pooled
        for w, b in self.fc[k]:
            x = ops.gemm(x, w, bias=b, relu=True)
        return ops.gemm(x, self.pred[k][0], bias=self.pred[k][1], out_dtype=torch.float32)

    def __call__(self, planes: List[ops.Plane], box_lists: List[torch.Tensor], image_sizes):
        """planes: p2..p5 as zero-bordered bf16 planes; box_lists: per image [Ri,4] fp32 CUDA; image_sizes: list of (h, w).
        Returns the corrected boxes per image, input order."""
        cfg = self.cfg
        counts = [len(b) for b in box_lists]
        boxes = torch.cat(box_lists).float().contiguous()
        roi_image = torch.cat([torch.full((c,), i, dtype=torch.int32, device=self.device) for i, c in enumerate(counts)])
        sizes = torch.tensor(image_sizes, dtype=torch.int32, device=self.device)
        img_col = roi_image.float()[:, None]
        for k in range(self.stages):
            rois = torch.cat([img_col, boxes], dim=1)
            pooled = ops.roi_pool_fpn(planes, [1.0 / s for s in STRIDES], rois, cfg.pooler_resolution, cfg.pooler_sampling_ratio,
                                      out_dtype=torch.bfloat16, out_layout=ops.OUT_NHWC)
            deltas = self.head(k, pooled.view(len(boxes), -1))
            # every stage output is clipped: stages 1,2 by _create_proposals_from_boxes, the last by fast_rcnn_inference
            boxes = ops.apply_deltas_clip(boxes, deltas, cfg.cascade_bbox_weights[k], roi_image, sizes)
        return list(boxes.split(counts))
