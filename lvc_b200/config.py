"""Detector configuration for the hot path.

A flat dataclass holding exactly the keys of the reference's yacs tree that the inference hot path
reads (detectron2/config/defaults.py + lvc/config/defaults.py; line cites per field).  Build one with
``DetectorConfig.from_reference_cfg(cfg)`` from a reference ``CfgNode`` (any object with the same
attribute tree works), or directly.
"""
from dataclasses import dataclass, field
from typing import Tuple


@dataclass
class DetectorConfig:
    # MODEL.RESNETS.DEPTH (detectron2/config/defaults.py:458; blocks per stage resnet.py:885-891)
    depth: int = 101
    # MODEL.PIXEL_MEAN / PIXEL_STD (detectron2/config/defaults.py:38-42), BGR
    pixel_mean: Tuple[float, float, float] = (103.530, 116.280, 123.675)
    pixel_std: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    size_divisibility: int = 32  # FPN strides[-1], fpn.py:101
    # MODEL.ANCHOR_GENERATOR.SIZES / ASPECT_RATIOS (configs/Base-RCNN-FPN.yaml:10-12)
    anchor_sizes: Tuple[float, ...] = (32.0, 64.0, 128.0, 256.0, 512.0)
    anchor_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    # MODEL.RPN.* (Base-RCNN-FPN.yaml:13-21; defaults.py:226,243)
    rpn_pre_nms_topk: int = 1000
    rpn_post_nms_topk: int = 1000
    rpn_nms_thresh: float = 0.7
    rpn_min_box_size: float = 0.0
    rpn_bbox_weights: Tuple[float, float, float, float] = (1.0, 1.0, 1.0, 1.0)
    # MODEL.ROI_HEADS / ROI_BOX_HEAD (Base-RCNN-FPN.yaml:22-28; defaults.py:276-304)
    num_classes: int = 80
    roi_bbox_weights: Tuple[float, float, float, float] = (10.0, 10.0, 5.0, 5.0)
    pooler_resolution: int = 7
    pooler_sampling_ratio: int = 0
    fc_dim: int = 1024
    num_fc: int = 2
    # MODEL.ROI_HEADS.OUTPUT_LAYER / COSINE_SCALE (lvc/config/defaults.py:97)
    output_layer: str = "FastRCNNOutputLayers"  # or "CosineSimOutputLayers"
    cosine_scale: float = 20.0
    score_thresh_test: float = 0.05   # defaults.py:276
    nms_thresh_test: float = 0.5      # defaults.py:279
    detections_per_image: int = 100   # TEST.DETECTIONS_PER_IMAGE defaults.py:585
    # box corrector (cascade yaml: 3 stages, NUM_FC 3; defaults.py:325-329)
    cascade_bbox_weights: Tuple[Tuple[float, float, float, float], ...] = (
        (10.0, 10.0, 5.0, 5.0), (20.0, 20.0, 10.0, 10.0), (30.0, 30.0, 15.0, 15.0))

    @property
    def blocks_per_stage(self):
        return {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}[self.depth]

    @staticmethod
    def from_reference_cfg(cfg) -> "DetectorConfig":
        m = cfg.MODEL
        sizes = tuple(float(s[0]) for s in m.ANCHOR_GENERATOR.SIZES)
        return DetectorConfig(
            depth=int(m.RESNETS.DEPTH),
            pixel_mean=tuple(m.PIXEL_MEAN), pixel_std=tuple(m.PIXEL_STD),
            anchor_sizes=sizes, anchor_ratios=tuple(float(r) for r in m.ANCHOR_GENERATOR.ASPECT_RATIOS[0]),
            rpn_pre_nms_topk=int(m.RPN.PRE_NMS_TOPK_TEST), rpn_post_nms_topk=int(m.RPN.POST_NMS_TOPK_TEST),
            rpn_nms_thresh=float(m.RPN.NMS_THRESH), rpn_min_box_size=0.0,
            rpn_bbox_weights=tuple(m.RPN.BBOX_REG_WEIGHTS),
            num_classes=int(m.ROI_HEADS.NUM_CLASSES), roi_bbox_weights=tuple(m.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
            pooler_resolution=int(m.ROI_BOX_HEAD.POOLER_RESOLUTION),
            pooler_sampling_ratio=int(m.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO),
            fc_dim=int(m.ROI_BOX_HEAD.FC_DIM), num_fc=int(m.ROI_BOX_HEAD.NUM_FC),
            output_layer=str(m.ROI_HEADS.OUTPUT_LAYER), cosine_scale=float(m.ROI_HEADS.COSINE_SCALE),
            score_thresh_test=float(m.ROI_HEADS.SCORE_THRESH_TEST), nms_thresh_test=float(m.ROI_HEADS.NMS_THRESH_TEST),
            detections_per_image=int(cfg.TEST.DETECTIONS_PER_IMAGE),
        )
