"""Deterministic synthetic weights with the reference's state-dict names / shapes / init laws.

There is no network for checkpoints, so benchmarks and parity tests use random-init weights of the
reference architecture (SURVEY.md Appendix B for names; SURVEY 8(d) for the calibration: every
bottleneck ``conv3.norm.weight = 0.2`` so activations stay finite and RPN yields 1000 proposals).
The dict loads with ``strict=True`` into the reference's ``GeneralizedRCNN`` (checked by
oracle/make_golden.py) -- so a real checkpoint with the same keys loads into this package too.
Init laws: c2_msra_fill for backbone convs (resnet.py:139-141), c2_xavier_fill for FPN / FC
(fpn.py:72-73, box_head.py:76-79), normal(0.01) for RPN + cls_score, normal(0.001) for bbox_pred
(rpn.py:99-101, fast_rcnn.py:551-554).
"""
import math

import torch

from .config import DetectorConfig


def _msra(g, cout, cin, k):
    std = math.sqrt(2.0 / (cout * k * k))  # kaiming_normal_, fan_out, relu
    return torch.randn(cout, cin, k, k, generator=g) * std


def _xavier_conv(g, cout, cin, k):
    bound = math.sqrt(3.0 / (cin * k * k))  # kaiming_uniform_(a=1): gain 1, fan_in
    return (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound


def _xavier_fc(g, cout, cin):
    bound = math.sqrt(3.0 / cin)
    return (torch.rand(cout, cin, generator=g) * 2 - 1) * bound


def _frozen_bn(sd, prefix, c, gamma=1.0, eps=1e-5):
    sd[prefix + ".weight"] = torch.full((c,), float(gamma))
    sd[prefix + ".bias"] = torch.zeros(c)
    sd[prefix + ".running_mean"] = torch.zeros(c)
    sd[prefix + ".running_var"] = torch.ones(c) - eps


def synthetic_state_dict(cfg: DetectorConfig, seed: int = 0, conv3_gamma: float = 0.2, randomize_bn: bool = False):
    """randomize_bn=True additionally draws non-trivial FrozenBN statistics (used by parity tests so that
    the BN folding is actually exercised)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def bn(prefix, c, gamma=1.0):
        _frozen_bn(sd, prefix, c, gamma)
        if randomize_bn:
            sd[prefix + ".weight"] = sd[prefix + ".weight"] * (0.8 + 0.4 * torch.rand(c, generator=g))
            sd[prefix + ".bias"] = 0.05 * torch.randn(c, generator=g)
            sd[prefix + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
            sd[prefix + ".running_var"] = 0.8 + 0.4 * torch.rand(c, generator=g)

    bu = "backbone.bottom_up."
    sd[bu + "stem.conv1.weight"] = _msra(g, 64, 3, 7)
    bn(bu + "stem.conv1.norm", 64)
    cin = 64
    for si, nblocks in enumerate(cfg.blocks_per_stage):
        stage = si + 2
        bott = 64 * 2 ** si
        cout = 256 * 2 ** si
        for b in range(nblocks):
            p = f"{bu}res{stage}.{b}."
            if b == 0:
                sd[p + "shortcut.weight"] = _msra(g, cout, cin, 1)
                bn(p + "shortcut.norm", cout)
            sd[p + "conv1.weight"] = _msra(g, bott, cin, 1)
            bn(p + "conv1.norm", bott)
            sd[p + "conv2.weight"] = _msra(g, bott, bott, 3)
            bn(p + "conv2.norm", bott)
            sd[p + "conv3.weight"] = _msra(g, cout, bott, 1)
            bn(p + "conv3.norm", cout, conv3_gamma)
            cin = cout
    for lvl, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        sd[f"backbone.fpn_lateral{lvl}.weight"] = _xavier_conv(g, 256, c, 1)
        sd[f"backbone.fpn_lateral{lvl}.bias"] = torch.zeros(256)
        sd[f"backbone.fpn_output{lvl}.weight"] = _xavier_conv(g, 256, 256, 3)
        sd[f"backbone.fpn_output{lvl}.bias"] = torch.zeros(256)
    A = len(cfg.anchor_ratios)
    rp = "proposal_generator.rpn_head."
    sd[rp + "conv.weight"] = torch.randn(256, 256, 3, 3, generator=g) * 0.01
    sd[rp + "conv.bias"] = torch.zeros(256)
    sd[rp + "objectness_logits.weight"] = torch.randn(A, 256, 1, 1, generator=g) * 0.01
    sd[rp + "objectness_logits.bias"] = torch.zeros(A)
    sd[rp + "anchor_deltas.weight"] = torch.randn(4 * A, 256, 1, 1, generator=g) * 0.01
    sd[rp + "anchor_deltas.bias"] = torch.zeros(4 * A)
    res = cfg.pooler_resolution
    d = 256 * res * res
    for i in range(cfg.num_fc):
        sd[f"roi_heads.box_head.fc{i + 1}.weight"] = _xavier_fc(g, cfg.fc_dim, d)
        sd[f"roi_heads.box_head.fc{i + 1}.bias"] = torch.zeros(cfg.fc_dim)
        d = cfg.fc_dim
    K = cfg.num_classes
    sd["roi_heads.box_predictor.cls_score.weight"] = torch.randn(K + 1, d, generator=g) * 0.01
    if cfg.output_layer != "CosineSimOutputLayers":
        sd["roi_heads.box_predictor.cls_score.bias"] = torch.zeros(K + 1)
    sd["roi_heads.box_predictor.bbox_pred.weight"] = torch.randn(4 * K, d, generator=g) * 0.001
    sd["roi_heads.box_predictor.bbox_pred.bias"] = torch.zeros(4 * K)
    return sd


def synthetic_corrector_head(cfg: DetectorConfig, seed: int = 0, num_fc: int = 3, stages: int = 3):
    """Box-corrector (CascadeROIHeads + BoxOnlyLayersCascade) head weights:
    roi_heads.box_head.{k}.fc{1..3}, roi_heads.box_predictor.{k}.bbox_pred (4, fc_dim)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    res = cfg.pooler_resolution
    for k in range(stages):
        d = 256 * res * res
        for i in range(num_fc):
            sd[f"roi_heads.box_head.{k}.fc{i + 1}.weight"] = _xavier_fc(g, cfg.fc_dim, d)
            sd[f"roi_heads.box_head.{k}.fc{i + 1}.bias"] = torch.zeros(cfg.fc_dim)
            d = cfg.fc_dim
        sd[f"roi_heads.box_predictor.{k}.bbox_pred.weight"] = torch.randn(4, d, generator=g) * 0.001
        sd[f"roi_heads.box_predictor.{k}.bbox_pred.bias"] = torch.zeros(4)
    return sd
