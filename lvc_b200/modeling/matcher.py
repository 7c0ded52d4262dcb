"""Drop-in for detectron2/modeling/matcher.py:8-126 (``Matcher``) and detectron2/structures/boxes.py:315-347 (``pairwise_iou``) on
liblvcb200 -- the label-assignment step of RPN.label_and_sample_anchors (rpn.py:277-326) and ROIHeads.label_and_sample_proposals
(lvc/modeling/roi_heads/roi_heads.py:173-278).  ``Matcher.match_boxes(gt, boxes)`` is the fused form: IoU and matching in one pass,
the [G, P] matrix (G x 268 569 for the RPN anchors of one image) is never written."""
import ctypes
from typing import List

import torch

from .. import _lib
from ..ops import _workspace


def pairwise_iou(boxes1, boxes2):
    """boxes1 [G,4], boxes2 [P,4] (tensors or objects with ``.tensor``) -> IoU [G,P] fp32, bit-exact with the reference's arithmetic."""
    b1 = getattr(boxes1, "tensor", boxes1)
    b2 = getattr(boxes2, "tensor", boxes2)
    _lib.require_cuda(b1, b2)
    b1, b2 = b1.detach().to(torch.float32).contiguous(), b2.detach().to(torch.float32).contiguous()
    out = torch.zeros((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=b1.device)
    if out.numel():
        _lib.check(_lib.load().lvcb200_pairwise_iou(_lib.ptr(b1), b1.shape[0], _lib.ptr(b2), b2.shape[0], _lib.ptr(out), _lib.stream_ptr()),
                   "lvcb200_pairwise_iou")
    return out


class Matcher:
    def __init__(self, thresholds: List[float], labels: List[int], allow_low_quality_matches: bool = False):
        thresholds = list(thresholds)
        assert thresholds[0] > 0                                        # matcher.py:46-55 (the +-inf ends are implicit in the C ABI)
        assert all(lo <= hi for lo, hi in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in [-1, 0, 1] for l in labels)
        assert len(labels) == len(thresholds) + 1
        self.thresholds, self.labels = thresholds, list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches

    def _run(self, gt, boxes, quality, G, P, dev):
        lib = _lib.load()
        matches = torch.zeros(P, dtype=torch.int64, device=dev)
        labels = torch.full((P,), self.labels[0], dtype=torch.int8, device=dev)
        if P == 0:
            return matches, labels
        thr = (ctypes.c_float * len(self.thresholds))(*self.thresholds)
        lab = (ctypes.c_int8 * len(self.labels))(*self.labels)
        ws = _workspace("match", lib.lvcb200_match_boxes_workspace(G), dev)
        rc = lib.lvcb200_match_boxes(_lib.ptr(gt), G, _lib.ptr(boxes), _lib.ptr(quality), P, thr, len(self.thresholds), lab,
                                     int(self.allow_low_quality_matches), _lib.ptr(matches), _lib.ptr(labels), None, _lib.ptr(ws),
                                     ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_match_boxes")
        return matches, labels

    def __call__(self, match_quality_matrix):
        """match_quality_matrix [G, P] fp32 (>= 0) -> matches int64 [P], match_labels int8 [P]  (matcher.py:61-100)."""
        _lib.require_cuda(match_quality_matrix)
        assert match_quality_matrix.dim() == 2
        q = match_quality_matrix.detach().to(torch.float32).contiguous()
        G, P = q.shape
        return self._run(None, None, q if G > 0 else None, G, P, q.device) if G > 0 else self._run(None, q.new_zeros((max(P, 1), 4)), None, 0, P, q.device)

    def match_boxes(self, gt_boxes, boxes):
        """Fused pairwise_iou + matching."""
        g = getattr(gt_boxes, "tensor", gt_boxes)
        b = getattr(boxes, "tensor", boxes)
        _lib.require_cuda(g, b)
        g, b = g.detach().to(torch.float32).contiguous(), b.detach().to(torch.float32).contiguous()
        return self._run(g if g.shape[0] else None, b, None, g.shape[0], b.shape[0], b.device)


def rpn_losses(anchors, pred_objectness_logits, pred_anchor_deltas, gt_labels, gt_boxes, batch_size_per_image=256,
               box2box_weights=(1.0, 1.0, 1.0, 1.0), smooth_l1_beta=0.0, loss_weight=None):
    """RPN.losses (rpn.py:328-400), smooth-L1 flavour: anchors [A,4] (levels concatenated), logits [N,A], deltas [N,A,4], gt_labels [N,A]
    (int8: -1 ignore / 0 / 1), gt_boxes [N,A,4] -> {"loss_rpn_cls", "loss_rpn_loc"} (0-d CUDA tensors).
    FORWARD ONLY (loss values for evaluation / logging): inputs are detached and there is no backward kernel, so the result carries
    no autograd graph -- training with it needs the dense backward passes, which are out of scope."""
    a = getattr(anchors, "tensor", anchors)
    _lib.require_cuda(a, pred_objectness_logits, pred_anchor_deltas, gt_labels, gt_boxes)
    a = a.detach().to(torch.float32).contiguous()
    lg = pred_objectness_logits.detach().to(torch.float32).contiguous()
    dl = pred_anchor_deltas.detach().to(torch.float32).contiguous()
    lb = gt_labels.detach().to(torch.int8).contiguous()
    gb = gt_boxes.detach().to(torch.float32).contiguous()
    N, A = lg.shape
    out = torch.zeros(2, dtype=torch.float64, device=lg.device)
    w = (ctypes.c_float * 4)(*box2box_weights)
    _lib.check(_lib.load().lvcb200_rpn_losses(_lib.ptr(a), _lib.ptr(lg), _lib.ptr(dl), _lib.ptr(lb), _lib.ptr(gb), N, A, w, float(smooth_l1_beta),
                                              _lib.ptr(out), _lib.stream_ptr()), "lvcb200_rpn_losses")
    norm = float(batch_size_per_image * N)
    lw = loss_weight or {}
    return {"loss_rpn_cls": (out[0] / norm).float() * lw.get("loss_rpn_cls", 1.0), "loss_rpn_loc": (out[1] / norm).float() * lw.get("loss_rpn_loc", 1.0)}


def fast_rcnn_losses(pred_class_logits, pred_proposal_deltas, gt_classes, proposal_boxes, gt_boxes, box2box_weights=(10.0, 10.0, 5.0, 5.0),
                     smooth_l1_beta=0.0, loss_weight=None):
    """FastRCNNOutputs.losses (lvc/modeling/roi_heads/fast_rcnn.py:424-438), smooth-L1 flavour: logits [R,K+1], deltas [R,4K] or [R,4],
    gt_classes [R] (K = background), proposal / gt boxes [R,4] -> {"loss_cls", "loss_box_reg"} (0-d CUDA tensors, both divided by R).
    FORWARD ONLY, like ``rpn_losses``: detached inputs, no backward kernel."""
    pb = getattr(proposal_boxes, "tensor", proposal_boxes)
    gb = getattr(gt_boxes, "tensor", gt_boxes)
    _lib.require_cuda(pred_class_logits, pred_proposal_deltas, gt_classes, pb, gb)
    lg = pred_class_logits.detach().to(torch.float32).contiguous()
    dl = pred_proposal_deltas.detach().to(torch.float32).contiguous()
    gc = gt_classes.detach().to(torch.int64).contiguous()
    pb, gb = pb.detach().to(torch.float32).contiguous(), gb.detach().to(torch.float32).contiguous()
    R, K1 = lg.shape
    out = torch.zeros(2, dtype=torch.float64, device=lg.device)
    w = (ctypes.c_float * 4)(*box2box_weights)
    _lib.check(_lib.load().lvcb200_fast_rcnn_losses(_lib.ptr(lg), _lib.ptr(dl), dl.shape[1], _lib.ptr(gc), _lib.ptr(pb), _lib.ptr(gb), R, K1 - 1, w,
                                                    float(smooth_l1_beta), _lib.ptr(out), _lib.stream_ptr()), "lvcb200_fast_rcnn_losses")
    lw = loss_weight or {}
    n = float(max(R, 1))
    return {"loss_cls": (out[0] / n).float() * lw.get("loss_cls", 1.0), "loss_box_reg": (out[1] / n).float() * lw.get("loss_box_reg", 1.0)}
