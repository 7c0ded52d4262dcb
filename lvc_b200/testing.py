"""Seeded synthetic-input generators shared by the parity tests, the golden generator and bench.py
(COCO-shaped boxes -- centres uniform in the image, log-uniform sides in [8, 600], clipped)."""
import numpy as np


def coco_like_boxes(rng, n, W=1333, H=800, min_side=8.0, max_side=600.0):
    cx = rng.uniform(0, W, n)
    cy = rng.uniform(0, H, n)
    w = np.exp(rng.uniform(np.log(min_side), np.log(max_side), n))
    h = np.exp(rng.uniform(np.log(min_side), np.log(max_side), n))
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = b[:, 0::2].clip(0, W)
    b[:, 1::2] = b[:, 1::2].clip(0, H)
    return b.astype(np.float32)


def distinct_scores(rng, n, lo=0.0, hi=1.0):
    """Tie-free fp32 scores (sort order on ties is implementation-defined in the reference)."""
    s = (rng.permutation(n).astype(np.float64) + 0.5) / n
    return (lo + (hi - lo) * s).astype(np.float32)


# ---------------------------------------------------------------------------------------------- end-to-end parity metrics
def box_iou_np(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None] - inter + 1e-12)


def e2e_parity_metrics(g, debug, boxes, scores, classes, counts):
    """Engine outputs against a reference-generated end-to-end fixture (tests/golden/e2e_*.npz, written by oracle/make_golden.py
    from the UNMODIFIED fp32 reference): per-level feature relative L2 on the fixture's sampled elements, the fraction of the
    reference's proposals the engine reproduces (a box within 0.05 px; and, looser, a box of IoU > 0.9), the fraction of its
    detections it reproduces (same class, IoU > 0.9), and the largest score / box difference among matched detections.
    ``debug`` is DetectorEngine.debug after a run; boxes / scores / classes / counts the run's outputs."""
    import torch
    out = {"features_rel_l2": {}}
    for l in (2, 3, 4, 5, 6):
        idx = torch.from_numpy(g[f"feat_p{l}_idx"].astype(np.int64))
        want = g[f"feat_p{l}_val"].astype(np.float64)
        got = debug["pyramid"][l].to_nchw().flatten()[idx.to(debug["pyramid"][l].t.device)].double().cpu().numpy()
        out["features_rel_l2"][f"p{l}"] = float(np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-30))
    n = len(g["sizes"])
    pm = pn = dm = dn = pi90 = 0
    max_ds, max_db = 0.0, 0.0
    for i in range(n):
        rp = g[f"prop_boxes{i}"]
        c = int(debug["prop_counts"][i])
        ours = debug["props"][i, :c].cpu().numpy()
        if len(rp) and len(ours):
            d = np.abs(rp[:, None, :] - ours[None, :, :]).max(-1)             # [ref, ours] max coordinate difference (px)
            pm += int((d.min(1) < 0.05).sum())
            pi90 += int((box_iou_np(rp, ours).max(1) > 0.9).sum())
        pn += len(rp)
        gb, gs, gc = g[f"det_boxes{i}"], g[f"det_scores{i}"], g[f"det_classes{i}"]
        k = int(counts[i])
        ob, os_, oc = boxes[i, :k].cpu().numpy(), scores[i, :k].cpu().numpy(), classes[i, :k].cpu().numpy()
        dn += len(gs)
        if len(gs) and k:
            iou = box_iou_np(gb, ob)
            iou[gc[:, None] != oc[None, :]] = -1.0
            j = iou.argmax(1)
            ok = iou.max(1) > 0.9
            dm += int(ok.sum())
            if ok.any():
                max_ds = max(max_ds, float(np.abs(os_[j][ok] - gs[ok]).max()))
                max_db = max(max_db, float(np.abs(ob[j][ok] - gb[ok]).max()))
    out.update(proposals_reproduced=pm / max(pn, 1), proposals_matched_iou90=pi90 / max(pn, 1), detections_reproduced=dm / max(dn, 1), n_ref_detections=dn,
               n_detections=int(sum(int(counts[i]) for i in range(n))), max_score_delta_matched=max_ds, max_box_delta_px_matched=max_db)
    return out
