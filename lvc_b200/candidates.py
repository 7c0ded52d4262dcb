"""Candidate filter between detection ("Label") and verification ("Verify"): the vectorised mirror of ``get_ret_anns``
(tools/create_coco_dataset_from_dets_all.py:129-193).  The reference walks pycocotools indices per novel class; here the
same decisions are taken on flat per-detection tensors (CPU or CUDA -- pure torch indexing, no arithmetic beyond compares),
so the filter can run right behind the detector without a JSON round trip.

    flags[i] = 1  pseudo-label candidate   (ann['ignore_qe'] = 0, ann['iscrowd'] = 0)
    flags[i] = 2  ignore region            (ann['ignore_qe'] = 1, ann['iscrowd'] = 1; --full only)
    flags[i] = 0  dropped
"""
from typing import Dict, Iterable, Set

import torch


def select_candidates(image_id: torch.Tensor, category: torch.Tensor, score: torch.Tensor, area: torch.Tensor,
                      image_area: torch.Tensor, train_imgs: Dict[int, Set[int]], novel_classes: Iterable[int], k_min, k_max,
                      ar: float = 0.0, full: bool = True, top: bool = False) -> torch.Tensor:
    """Arguments mirror the reference's CLI: ``--K-min/--K-max`` (score bounds, or ranks with ``top``), ``--ar`` (minimum
    box-area / image-area ratio), ``--full`` (mark the other same-class detections of the kept images as ignore regions).
    Score mode keeps ``K_min < score <= K_max`` (the reference's left ``searchsorted`` on ``-scores``, :169-174); valid
    detections are those of a novel class on images that do NOT hold that class's few-shot ground truth (:133-134), with
    ``0 < area < 1e10`` and ``ar < area / image_area < 1`` (:39-43, :136-137).
    One deliberate difference: with ``full`` and NO kept detection for a class the reference calls pycocotools'
    ``getAnnIds(imgIds=[])``, which treats the empty list as "all images" and therefore marks every valid detection of that class as
    an ignore region (:175-186); here (and in ``lvcb200_candidate_filter``) such a class yields nothing, which is what the
    surrounding code intends ("the other detections of the images that hold a pseudo-label")."""
    dev = score.device
    n = score.shape[0]
    flags = torch.zeros(n, dtype=torch.int8, device=dev)
    if n == 0:
        return flags
    area64, ratio = area.double(), area.double() / image_area.double()
    geom_ok = (area64 > 0.0) & (area64 < 1e10) & (ratio > ar) & (ratio < 1.0)
    sc = score.float().double()
    for cid in novel_classes:
        valid = (category == cid) & geom_ok
        excl = train_imgs.get(cid)
        if excl:
            valid &= ~torch.isin(image_id, torch.tensor(sorted(excl), dtype=image_id.dtype, device=dev))
        if top:
            idx = valid.nonzero().flatten()
            order = idx[torch.sort(-sc[idx], stable=True).indices]
            keep_idx = order[int(k_max):int(k_min)]
            keep = torch.zeros(n, dtype=torch.bool, device=dev)
            keep[keep_idx] = True
        else:
            keep = valid & (sc > float(k_min)) & (sc <= float(k_max))
        flags[keep] = 1
        if full:
            pres = torch.unique(image_id[keep])
            flags[valid & ~keep & torch.isin(image_id, pres)] = 2
    return flags


class CandidateFilter:
    """Device-side score-mode filter (``lvcb200_candidate_filter``): the same decisions as ``select_candidates(top=False)``,
    taken per batch on the detector's output block right behind the NMS -- nothing but the flagged detections needs to leave
    the GPU.  Classes are the detector's contiguous ids (the reference filters dataset ids after the id mapping of
    coco_evaluation.py:288-312; map ``novel_classes`` / ``train_imgs`` keys accordingly).  Attach to a model with
    ``model.candidate_filter = CandidateFilter(...)``: every result ``Instances`` then carries ``candidate_flags``
    (1 = pseudo-label candidate, 2 = ignore region, 0 = dropped)."""

    def __init__(self, novel_classes: Iterable[int], k_min: float, k_max: float, ar: float = 0.0, full: bool = True,
                 train_imgs: Dict[int, Set[int]] = None, num_classes: int = 80, device="cuda"):
        self.k_min, self.k_max, self.ar, self.full = float(k_min), float(k_max), float(ar), bool(full)
        self.num_classes = num_classes
        self.novel_ids = sorted(int(c) for c in novel_classes)
        self.train_imgs = {int(c): set(v) for c, v in (train_imgs or {}).items() if v}
        nv = torch.zeros(num_classes, dtype=torch.uint8)
        nv[self.novel_ids] = 1
        self.novel = nv.to(device)

    def excluded(self, image_ids):
        """[n, K] uint8 host matrix: image i holds the few-shot ground truth of class c (create_coco_dataset_from_dets_all.py:133-134)."""
        if not self.train_imgs or image_ids is None:
            return None
        ex = torch.zeros((len(image_ids), self.num_classes), dtype=torch.uint8)
        for c, imgs in self.train_imgs.items():
            for i, iid in enumerate(image_ids):
                if iid in imgs:
                    ex[i, c] = 1
        return ex

    def __call__(self, boxes, scores, classes, counts, out_sizes, image_ids=None):
        """boxes [n,topk,4] fp32, scores [n,topk], classes [n,topk] int64, counts [n] int32 (CUDA, lvcb200_detections' outputs);
        out_sizes: list of (height, width) the boxes are expressed in.  Returns (flags [n,topk] int8, n_keep [n] int32)."""
        from . import _lib
        _lib.require_cuda(boxes, scores, classes, counts)
        n, topk = scores.shape
        dev = scores.device
        area = torch.tensor([float(h) * float(w) for h, w in out_sizes], dtype=torch.float64).to(dev, non_blocking=True)
        ex = self.excluded(image_ids)
        ex_dev = ex.to(dev, non_blocking=True) if ex is not None else None
        flags = torch.empty((n, topk), dtype=torch.int8, device=dev)
        n_keep = torch.empty(n, dtype=torch.int32, device=dev)
        rc = _lib.load().lvcb200_candidate_filter(_lib.ptr(boxes.contiguous()), _lib.ptr(scores.contiguous()), _lib.ptr(classes.contiguous()),
                                                  _lib.ptr(counts), _lib.ptr(area), n, topk, self.num_classes, _lib.ptr(self.novel),
                                                  _lib.ptr(ex_dev), self.k_min, self.k_max, self.ar, int(self.full), _lib.ptr(flags),
                                                  _lib.ptr(n_keep), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_candidate_filter")
        return flags, n_keep
