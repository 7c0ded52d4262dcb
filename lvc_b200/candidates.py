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
    ``0 < area < 1e10`` and ``ar < area / image_area < 1`` (:39-43, :136-137)."""
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
