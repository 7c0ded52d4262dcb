"""CPU oracle for the LVC pseudo-label mining hot path (python side).

TEST INFRASTRUCTURE ONLY -- the checker for lvc_b200's CUDA path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs may import this module.

Integer / index / byte decisions (NMS, level assignment, top-k, vote) are computed by the scalar C
restatement in ``lvc_oracle.c`` (loaded through ctypes); this file holds the glue that strings those
pieces together exactly the way the reference's Python does, plus a dense fp32 restatement of the
detector (conv / linear through torch's CPU kernels, which is the very library layer the reference
itself calls: detectron2/layers/wrappers.py:94-98 -> F.conv2d).

Parity pinning: see the header of lvc_oracle.c and tests/test_oracle_golden.py.
Reference citations are relative to /root/reference (prannaykaul/lvc @ 3b5e5fa).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblvc_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """gcc-compile lvc_oracle.c -> liblvc_oracle.so (in oracle/, git-ignored)."""
    src = os.path.join(_HERE, "lvc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden",
             "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_nms.restype = ctypes.c_int64
        _lib.orc_batched_nms.restype = ctypes.c_int64
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _p(a, t):
    return a.ctypes.data_as(t)


# ------------------------------------------------------------------------------------ RoIAlign
def roi_align(inp, rois, output_size, spatial_scale, sampling_ratio, aligned):
    """detectron2/layers/roi_align.py:63-108 -> torchvision.ops.roi_align; NCHW fp32."""
    inp = _f32(inp)
    rois = _f32(rois).reshape(-1, 5)
    N, C, H, W = inp.shape
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    out = np.zeros((rois.shape[0], C, ph, pw), np.float32)
    if rois.shape[0]:
        lib().orc_roi_align_forward(_p(inp, _f32p), N, C, H, W, _p(rois, _f32p), rois.shape[0], ph, pw,
                                    ctypes.c_float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                                    _p(out, _f32p))
    return out


def assign_boxes_to_levels(boxes, min_level=2, max_level=5, canonical_box_size=224, canonical_level=4):
    """detectron2/modeling/poolers.py:23-59."""
    boxes = _f32(boxes).reshape(-1, 4)
    out = np.zeros(boxes.shape[0], np.int64)
    lib().orc_assign_boxes_to_levels(_p(boxes, _f32p), boxes.shape[0], min_level, max_level,
                                     canonical_box_size, canonical_level, _p(out, _i64p))
    return out


def roi_pooler(features, box_lists, output_size=7, scales=(0.25, 0.125, 0.0625, 0.03125),
               sampling_ratio=0, canonical_box_size=224, canonical_level=4):
    """ROIPooler.forward, detectron2/modeling/poolers.py:191-246 (pooler type ROIAlignV2: aligned=True).

    features: list of NCHW arrays; box_lists: list (per image) of [Ri,4] arrays.
    Returns (pooled [M,C,s,s], level_assignments [M]).
    """
    min_level = int(-math.log2(scales[0]))
    max_level = int(-math.log2(scales[-1]))
    fmt = np.concatenate(
        [np.concatenate([np.full((len(b), 1), i, np.float32), _f32(b).reshape(-1, 4)], 1)
         for i, b in enumerate(box_lists)], 0)
    allb = np.concatenate([_f32(b).reshape(-1, 4) for b in box_lists], 0)
    if len(scales) == 1:
        return roi_align(features[0], fmt, output_size, scales[0], sampling_ratio, True), np.zeros(len(fmt), np.int64)
    lvls = assign_boxes_to_levels(allb, min_level, max_level, canonical_box_size, canonical_level)
    C = features[0].shape[1]
    out = np.zeros((len(fmt), C, output_size, output_size), np.float32)
    for level, scale in enumerate(scales):
        inds = np.nonzero(lvls == level)[0]
        out[inds] = roi_align(features[level], fmt[inds], output_size, scale, sampling_ratio, True)
    return out, lvls


# ------------------------------------------------------------------------------------ NMS
def nms(boxes, scores, thr):
    """torchvision.ops.nms as reached from detectron2/layers/nms.py:7,25."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores)
    keep = np.zeros(len(scores), np.int64)
    n = lib().orc_nms(_p(boxes, _f32p), _p(scores, _f32p), ctypes.c_int64(len(scores)), ctypes.c_float(thr),
                      _p(keep, _i64p))
    return keep[:n].copy()


TRICK, VANILLA = 0, 1


def nms_mode_for(n_boxes, device="cuda"):
    """Which branch the reference takes for ``n_boxes`` (detectron2/layers/nms.py:19-29 on top of
    torchvision.ops.boxes.batched_nms of torchvision 0.26: trick iff numel <= 4000 (cpu) / 100000 (cuda))."""
    if n_boxes >= 40000:
        return VANILLA
    limit = 4000 if device == "cpu" else 100_000
    return VANILLA if n_boxes * 4 > limit else TRICK


def batched_nms(boxes, scores, idxs, thr, mode=None, device="cuda"):
    """detectron2/layers/nms.py:10-29.  mode None = the branch the reference would take on ``device``."""
    boxes = _f32(boxes).reshape(-1, 4)
    scores = _f32(scores)
    idxs = _i64(idxs)
    if mode is None:
        mode = nms_mode_for(len(scores), device)
    keep = np.zeros(len(scores), np.int64)
    n = lib().orc_batched_nms(_p(boxes, _f32p), _p(scores, _f32p), _p(idxs, _i64p), ctypes.c_int64(len(scores)),
                              ctypes.c_float(thr), int(mode), _p(keep, _i64p))
    return keep[:n].copy()


# ------------------------------------------------------------------------------------ boxes / anchors
SCALE_CLAMP = math.log(1000.0 / 16)  # detectron2/modeling/box_regression.py:14


def apply_deltas(deltas, boxes, weights):
    """Box2BoxTransform.apply_deltas, detectron2/modeling/box_regression.py:73-110."""
    deltas = _f32(deltas)
    boxes = _f32(boxes).reshape(-1, 4)
    R = boxes.shape[0]
    deltas = deltas.reshape(R, -1)
    K = deltas.shape[1] // 4
    out = np.zeros_like(deltas)
    if R:
        lib().orc_apply_deltas(_p(deltas, _f32p), _p(boxes, _f32p), R, K, *[ctypes.c_float(w) for w in weights],
                               ctypes.c_float(SCALE_CLAMP), _p(out, _f32p))
    return out


def clip_boxes(boxes, image_size):
    """Boxes.clip, detectron2/structures/boxes.py:183-196.  image_size = (h, w)."""
    b = _f32(boxes).reshape(-1, 4).copy()
    lib().orc_clip_boxes(_p(b, _f32p), ctypes.c_int64(len(b)), int(image_size[0]), int(image_size[1]))
    return b


def nonempty(boxes, threshold=0.0):
    """Boxes.nonempty, detectron2/structures/boxes.py:198-212."""
    b = _f32(boxes).reshape(-1, 4)
    return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)


def cell_anchors(sizes, ratios):
    """generate_cell_anchors, detectron2/modeling/anchor_generator.py:173-208."""
    s = np.ascontiguousarray(sizes, np.float64)
    r = np.ascontiguousarray(ratios, np.float64)
    out = np.zeros((len(s) * len(r), 4), np.float32)
    lib().orc_cell_anchors(_p(s, _f64p), len(s), _p(r, _f64p), len(r), _p(out, _f32p))
    return out


def grid_anchors(cell, H, W, stride):
    """_grid_anchors, detectron2/modeling/anchor_generator.py:157-171 (offset 0)."""
    cell = _f32(cell)
    out = np.zeros((H * W * len(cell), 4), np.float32)
    lib().orc_grid_anchors(_p(cell, _f32p), len(cell), H, W, stride, _p(out, _f32p))
    return out


# ------------------------------------------------------------------------------------ RPN post-processing
def find_top_rpn_proposals(proposals, logits, image_sizes, nms_thresh=0.7, pre_nms_topk=1000,
                           post_nms_topk=1000, min_box_size=0.0, nms_mode=None, device="cuda"):
    """detectron2/modeling/proposal_generator/proposal_utils.py:13-118 (inference branch).

    proposals: list over levels of [N, HWA, 4]; logits: list of [N, HWA].
    Returns per image (boxes [K,4], logits [K]).  Sort ties: lower index first (stable).
    """
    N = len(image_sizes)
    tk_scores, tk_props, lvl_ids = [], [], []
    for level_id, (p, l) in enumerate(zip(proposals, logits)):
        k = min(pre_nms_topk, l.shape[1])
        idx = np.argsort(-_f32(l), axis=1, kind="stable")[:, :k]
        tk_scores.append(np.take_along_axis(_f32(l), idx, 1))
        tk_props.append(np.take_along_axis(_f32(p), idx[:, :, None], 1))
        lvl_ids.append(np.full(k, level_id, np.int64))
    tk_scores = np.concatenate(tk_scores, 1)
    tk_props = np.concatenate(tk_props, 1)
    lvl_ids = np.concatenate(lvl_ids)
    res = []
    for n in range(N):
        boxes, sc, lvl = tk_props[n], tk_scores[n], lvl_ids
        valid = np.isfinite(boxes).all(1) & np.isfinite(sc)
        boxes, sc, lvl = boxes[valid], sc[valid], lvl[valid]
        boxes = clip_boxes(boxes, image_sizes[n])
        keep = nonempty(boxes, min_box_size)
        boxes, sc, lvl = boxes[keep], sc[keep], lvl[keep]
        k = batched_nms(boxes, sc, lvl, nms_thresh, mode=nms_mode, device=device)[:post_nms_topk]
        res.append((boxes[k], sc[k]))
    return res


# ------------------------------------------------------------------------------------ box-head post-processing
def softmax_rows(x):
    x = _f32(x)
    out = np.zeros_like(x)
    lib().orc_softmax_rows(_p(x, _f32p), ctypes.c_int64(x.shape[0]), x.shape[1], _p(out, _f32p))
    return out


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image,
                                     nms_mode=None, device="cuda"):
    """lvc/modeling/roi_heads/fast_rcnn.py:95-137.  boxes [R, K*4] (or [R,4]), scores [R, K+1].

    Returns (pred_boxes [n,4], scores [n], pred_classes [n], kept_row_idx [n]).
    """
    scores = _f32(scores)[:, :-1]
    nreg = boxes.shape[1] // 4
    b = clip_boxes(_f32(boxes).reshape(-1, 4), image_shape).reshape(-1, nreg, 4)
    mask = scores > np.float32(score_thresh)
    r_idx, c_idx = np.nonzero(mask)
    bsel = b[r_idx, 0] if nreg == 1 else b[mask]
    ssel = scores[mask]
    keep = batched_nms(bsel, ssel, c_idx, nms_thresh, mode=nms_mode, device=device)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    return bsel[keep], ssel[keep], c_idx[keep].astype(np.int64), r_idx[keep].astype(np.int64)


def detector_postprocess(boxes, image_size, out_h, out_w):
    """detectron2/modeling/postprocessing.py:10-79 (boxes only).  Returns (scaled+clipped boxes, keep mask)."""
    sx, sy = out_w / image_size[1], out_h / image_size[0]
    b = _f32(boxes).reshape(-1, 4).copy()
    b[:, 0::2] *= np.float32(sx)
    b[:, 1::2] *= np.float32(sy)
    b = clip_boxes(b, (out_h, out_w))
    return b, nonempty(b)


# ------------------------------------------------------------------------------------ kNN label verification
def knn_verify(bank, bank_cls, queries, query_cls, topk=10, knn=10):
    """tools/run_nearest_neighbours.py:142-162 + :214-227.

    Returns dict(top_idx [Q,topk], top_sim, votes, nn_class [Q], keep [Q] uint8)."""
    bank = _f32(bank)
    queries = _f32(queries)
    bank_cls = _i64(bank_cls)
    query_cls = _i64(query_cls)
    Q, D = queries.shape
    S = bank.shape[0]
    top_idx = np.full((Q, topk), -1, np.int64)
    top_sim = np.zeros((Q, topk), np.float32)
    votes = np.zeros((Q, topk), np.int64)
    nn_class = np.zeros(Q, np.int64)
    keep = np.zeros(Q, np.uint8)
    lib().orc_knn_verify(_p(bank, _f32p), _p(bank_cls, _i64p), S, D, _p(queries, _f32p), _p(query_cls, _i64p),
                         ctypes.c_int64(Q), topk, knn, _p(top_idx, _i64p), _p(top_sim, _f32p), _p(votes, _i64p),
                         _p(nn_class, _i64p), _p(keep, _u8p))
    return dict(top_idx=top_idx, top_sim=top_sim, votes=votes, nn_class=nn_class, keep=keep)


def knn_verify_batched_torch(bank, bank_cls, queries, query_cls, topk=10, knn=10, chunk=8192):
    """Same result through torch CPU matmul (multi-threaded) -- used as the *fast* CPU baseline form
    (BASELINE.md section 2 'batched-GEMM form'); the scalar C version above is the checker."""
    import torch
    bank = torch.as_tensor(bank, dtype=torch.float32)
    q = torch.as_tensor(queries, dtype=torch.float32)
    mu = bank.mean(0, keepdim=True)
    bn = torch.nn.functional.normalize(bank - mu, dim=1, eps=1e-8)
    bcls = torch.as_tensor(bank_cls)
    outs = []
    for s in range(0, q.shape[0], chunk):
        qn = torch.nn.functional.normalize(q[s:s + chunk] - mu, dim=1, eps=1e-8)
        outs.append((qn @ bn.t()).topk(topk, dim=1)[1])
    idx = torch.cat(outs)
    votes = bcls[idx]
    nn_class = torch.mode(votes[:, :knn], dim=1)[0]
    keep = (nn_class == torch.as_tensor(query_cls)).to(torch.uint8)
    return dict(top_idx=idx.numpy(), votes=votes.numpy(), nn_class=nn_class.numpy(), keep=keep.numpy())


def knn_reference_form_torch(bank, bank_cls, queries, per_call=5, topk=10):
    """The reference's literal per-image broadcast form (run_nearest_neighbours.py:146-161), Q=per_call
    queries per call -- the honest 'reference CPU path' for timing."""
    import torch
    import torch.nn.functional as F
    bank = torch.as_tensor(bank, dtype=torch.float32)
    q = torch.as_tensor(queries, dtype=torch.float32)
    bcls = torch.as_tensor(bank_cls)
    crop_mean = bank.mean(dim=0, keepdim=True)
    out = []
    for s in range(0, q.shape[0], per_call):
        qf = q[s:s + per_call]
        sim = F.cosine_similarity(bank.sub(crop_mean).unsqueeze(0).float(), qf.sub(crop_mean).unsqueeze(1).float(), dim=-1)
        out.append(bcls[sim.topk(topk, dim=-1)[1]])
    return torch.cat(out).numpy()


def knn_cdist_torch(bank, bank_cls, queries, query_cls=None, topk=10, knn=10, chunk=4096):
    """The QUERY_EXPAND.COSINE_SIM = False branch (run_nearest_neighbours.py:154-159): sim = -torch.cdist(bank, query), no centring.
    Returns dict(top_idx, top_sim, votes[, keep])."""
    import torch
    bank = torch.as_tensor(bank, dtype=torch.float32)
    q = torch.as_tensor(queries, dtype=torch.float32)
    bcls = torch.as_tensor(bank_cls)
    sims, idxs = [], []
    for s in range(0, q.shape[0], chunk):
        sim = torch.cdist(bank.unsqueeze(0), q[s:s + chunk].unsqueeze(0)).squeeze(0).t().mul(-1.0)
        v, i = sim.topk(topk, dim=-1)
        sims.append(v)
        idxs.append(i)
    idx = torch.cat(idxs)
    votes = bcls[idx]
    out = dict(top_idx=idx.numpy(), top_sim=torch.cat(sims).numpy(), votes=votes.numpy())
    if query_cls is not None:
        out["keep"] = (torch.mode(votes[:, :knn], dim=1)[0] == torch.as_tensor(query_cls)).to(torch.uint8).numpy()
    return out


# ------------------------------------------------------------------------------------ candidate filter (a15)
def select_candidates(image_id, category, score, area, image_area, train_imgs, novel_classes, k_min, k_max, ar=0.0, full=True,
                      top=False):
    """get_ret_anns, tools/create_coco_dataset_from_dets_all.py:129-193, on flat arrays (one entry per detection, in
    annotation-id order).  train_imgs: dict class -> set(image ids holding that class's few-shot GT).
    Returns flags [n] int8: 0 = dropped, 1 = pseudo-label candidate (ignore_qe=0, iscrowd=0), 2 = ignore region
    (ignore_qe=1, iscrowd=1; only with ``full``).  Score mode keeps K_min < score <= K_max (left searchsorted on -scores,
    :169-174); top mode keeps ranks [K_max, K_min) of the per-class descending-score list (:143-149)."""
    n = len(score)
    flags = np.zeros(n, np.int8)
    image_id, category = np.asarray(image_id), np.asarray(category)
    score, area, image_area = _f32(score).astype(np.float64), np.asarray(area, np.float64), np.asarray(image_area, np.float64)
    ratio = area / image_area
    for cid in novel_classes:
        excl = train_imgs.get(cid, set())
        valid = [i for i in range(n) if category[i] == cid and image_id[i] not in excl and 0.0 < area[i] < 1e10
                 and ar < ratio[i] < 1.0]
        order = sorted(valid, key=lambda i: score[i], reverse=True)       # python's sort is stable, like the reference's
        if top:
            keep = order[int(k_max):int(k_min)]
        else:
            sc = np.array([score[i] for i in order])
            keep = order[int(np.searchsorted(-sc, -float(k_max))):int(np.searchsorted(-sc, -float(k_min)))]
        for i in keep:
            flags[i] = 1
        if full:
            pres = set(image_id[i] for i in keep)
            for i in valid:
                if image_id[i] in pres and flags[i] != 1:
                    flags[i] = 2
    return flags


# ------------------------------------------------------------------------------------ crops (next row f1)
def get_crops_qe(img, boxes, operation="context", size=224):
    """lvc/data/utils.py:485-519 restated with explicit loops: img [3,H,W], integer boxes [n,4] -> [n,3,size,size] float32."""
    img = np.asarray(img)
    _, H, W = img.shape

    def get_padding(h, w):
        max_d = max(h, w)
        hp, vp = (max_d - w) / 2, (max_d - h) / 2
        l = hp if hp % 1 == 0 else hp + 0.5
        t = vp if vp % 1 == 0 else vp + 0.5
        r = hp if hp % 1 == 0 else hp - 0.5
        b = vp if vp % 1 == 0 else vp - 0.5
        return int(l), int(r), int(t), int(b)

    out = np.zeros((len(boxes), 3, size, size), np.float32)
    for i, (x1, y1, x2, y2) in enumerate(np.asarray(boxes, np.int64).tolist()):
        if operation == "pad":
            l_p, r_p, t_p, b_p = get_padding(y2 - y1 + 1, x2 - x1 + 1)
            crop = img[:, y1:y2 + 1, x1:x2 + 1]
        else:
            l_p, r_p, t_p, b_p = get_padding(y2 - y1 + 1, x2 - x1 + 1)
            y1n, x1n = max(0, y1 - t_p), max(0, x1 - l_p)
            y2n, x2n = min(H, y2 + b_p), min(W, x2 + r_p)
            l_p, r_p, t_p, b_p = get_padding(y2n - y1n + 1, x2n - x1n + 1)
            crop = img[:, y1n:y2n + 1, x1n:x2n + 1]
        pad = np.zeros((3, crop.shape[1] + t_p + b_p, crop.shape[2] + l_p + r_p), np.float32)
        pad[:, t_p:t_p + crop.shape[1], l_p:l_p + crop.shape[2]] = crop
        Hp, Wp = pad.shape[1:]
        sy, sx = np.float32(Hp) / np.float32(size), np.float32(Wp) / np.float32(size)
        for oy in range(size):
            py = min(int(np.floor(np.float32(oy) * sy)), Hp - 1)
            for ox in range(size):
                px = min(int(np.floor(np.float32(ox) * sx)), Wp - 1)
                out[i, :, oy, ox] = pad[:, py, px]
    return out


# ------------------------------------------------------------------------------------ "next" row f4: training-side ops
def roi_align_backward(grad_out, rois, input_shape, spatial_scale, sampling_ratio, aligned):
    """autograd backward of detectron2.layers.roi_align (ROIAlign_cuda.cu:141-306): grad_out [R,C,ph,pw] -> grad_in [N,C,H,W]."""
    g, r = _f32(grad_out), _f32(rois)
    N, C, H, W = input_shape
    R, _, ph, pw = g.shape
    out = np.zeros((N, C, H, W), np.float32)
    lib().orc_roi_align_backward(_p(g, _f32p), N, C, H, W, _p(r, _f32p), R, ph, pw, ctypes.c_float(spatial_scale),
                                 int(sampling_ratio), int(bool(aligned)), _p(out, _f32p))
    return out


def pairwise_iou(boxes1, boxes2):
    """detectron2/structures/boxes.py:315-347 -> [G, P] fp32."""
    a, b = _f32(boxes1).reshape(-1, 4), _f32(boxes2).reshape(-1, 4)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().orc_pairwise_iou(_p(a, _f32p), ctypes.c_int64(a.shape[0]), _p(b, _f32p), ctypes.c_int64(b.shape[0]), _p(out, _f32p))
    return out


def match_boxes(gt, boxes, thresholds, labels, allow_low_quality):
    """Matcher(thresholds, labels, allow_low_quality)(pairwise_iou(gt, boxes)) (matcher.py:61-126) -> matches int64 [P], labels int8 [P]."""
    a, b = _f32(gt).reshape(-1, 4), _f32(boxes).reshape(-1, 4)
    P = b.shape[0]
    m, l, v = np.zeros(P, np.int64), np.zeros(P, np.int8), np.zeros(P, np.float32)
    th, lb = _f32(thresholds), np.ascontiguousarray(labels, np.int8)
    lib().orc_match_boxes(_p(a, _f32p), ctypes.c_int64(a.shape[0]), _p(b, _f32p), ctypes.c_int64(P), _p(th, _f32p), len(th),
                          lb.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), int(bool(allow_low_quality)), _p(m, _i64p),
                          l.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), _p(v, _f32p))
    return m, l, v


def rpn_losses(anchors, logits, deltas, labels, gt_boxes, weights=(1.0, 1.0, 1.0, 1.0), beta=0.0):
    """The two un-normalised sums of RPN.losses (rpn.py:328-400): (objectness BCE, localisation smooth-L1)."""
    a, lg, dl, gb = _f32(anchors), _f32(logits), _f32(deltas), _f32(gt_boxes)
    lb = np.ascontiguousarray(labels, np.int8)
    N, A = lg.shape
    out = np.zeros(2, np.float64)
    lib().orc_rpn_losses(_p(a, _f32p), _p(lg, _f32p), _p(dl, _f32p), lb.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), _p(gb, _f32p),
                         ctypes.c_int64(N), ctypes.c_int64(A), _p(_f32(weights), _f32p), ctypes.c_float(beta), _p(out, _f64p))
    return out


def fast_rcnn_losses(logits, deltas, gt_classes, proposals, gt_boxes, weights=(10.0, 10.0, 5.0, 5.0), beta=0.0):
    """The two un-normalised sums of FastRCNNOutputs.losses (lvc fast_rcnn.py:267-358, 424-438): (cross entropy, box smooth-L1)."""
    lg, dl, pr, gb = _f32(logits), _f32(deltas), _f32(proposals), _f32(gt_boxes)
    gc = _i64(gt_classes)
    R, K1 = lg.shape
    out = np.zeros(2, np.float64)
    lib().orc_fast_rcnn_losses(_p(lg, _f32p), _p(dl, _f32p), dl.shape[1], _p(gc, _i64p), _p(pr, _f32p), _p(gb, _f32p), ctypes.c_int64(R), K1 - 1,
                               _p(_f32(weights), _f32p), ctypes.c_float(beta), _p(out, _f64p))
    return out


def subsample_labels(labels, keys, num_samples, positive_fraction, bg_label):
    """detectron2/modeling/sampling.py:9-54 with the two ``torch.randperm`` draws replaced by the stable argsort of per-element
    random keys (``keys`` uint32 [N]): positive[argsort(keys[positive])[:num_pos]], negative[argsort(keys[negative])[:num_neg]].
    Returns (pos_idx, neg_idx) int64."""
    labels = np.asarray(labels).astype(np.int64)
    keys = np.asarray(keys, np.uint32)
    positive = [i for i in range(len(labels)) if labels[i] != -1 and labels[i] != bg_label]
    negative = [i for i in range(len(labels)) if labels[i] == bg_label]
    num_pos = min(len(positive), int(num_samples * positive_fraction))
    num_neg = min(len(negative), num_samples - num_pos)
    pos = sorted(positive, key=lambda i: (int(keys[i]), i))[:num_pos]
    neg = sorted(negative, key=lambda i: (int(keys[i]), i))[:num_neg]
    return np.asarray(pos, np.int64), np.asarray(neg, np.int64)
