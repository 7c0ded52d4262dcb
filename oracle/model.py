"""Dense fp32 CPU restatement of the detector forward (TEST INFRASTRUCTURE ONLY; see oracle.py header).

Functional, state-dict driven (reference key names).  Dense layers run through
torch's CPU kernels (the same L0 library the reference calls); every index / selection decision goes
through the C oracle in ``oracle.py``.

Reference rows: a1 preprocess (lvc/modeling/meta_arch/rcnn.py:324-333;
detectron2/structures/image_list.py:57-119), a2 ResNet (detectron2/modeling/backbone/resnet.py:564-592,
:195-211, :708-731), a3 FPN (fpn.py:109-144, :165-177), a4 RPN head (proposal_generator/rpn.py:120-139),
a5-a8 anchors / decode / top-k / NMS, a9-a10 ROIPooler, a11 box head (lvc/modeling/roi_heads/box_head.py:82-91),
a12 predictors (fast_rcnn.py:583-598, :811-841), a13 inference (:95-137, :440-493), a14 postprocess,
a17 box corrector (cascade_rcnn.py:167-203, roi_heads_cascade.py:134-138,197-211).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as O


_EMULATE_BF16 = False   # set by detector_forward(emulate_bf16=True): restate the engine's precision policy


def _q(x):
    """Round to bf16 and back (what an activation stored in a bf16 plane goes through)."""
    return x.bfloat16().float() if _EMULATE_BF16 else x


def _conv_bn(sd, prefix, x, stride=1, padding=0, relu=False, eps=1e-5, residual=None, quant_out=True):
    if _EMULATE_BF16:
        # engine policy: FrozenBN folded into the weights in fp32, weights cast to bf16, fp32 accumulate,
        # bias / residual / ReLU in fp32, one rounding to bf16 on store
        w = sd[prefix + ".weight"].float()
        if prefix + ".norm.weight" in sd:
            scale = sd[prefix + ".norm.weight"] * (sd[prefix + ".norm.running_var"] + eps).rsqrt()
            bias = sd[prefix + ".norm.bias"] - sd[prefix + ".norm.running_mean"] * scale
            w = w * scale.view(-1, 1, 1, 1)
        else:
            bias = sd.get(prefix + ".bias")
        y = F.conv2d(x, w.bfloat16().float(), bias, stride=stride, padding=padding)
        if residual is not None:
            y = y + residual
        if relu:
            y = F.relu_(y)
        return _q(y) if quant_out else y
    x = F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)
    if prefix + ".norm.weight" in sd:
        x = F.batch_norm(x, sd[prefix + ".norm.running_mean"], sd[prefix + ".norm.running_var"],
                         sd[prefix + ".norm.weight"], sd[prefix + ".norm.bias"], training=False, eps=eps)
    if residual is not None:
        x = x + residual
    return F.relu_(x) if relu else x


def preprocess(cfg, images):
    """images: list of [3,H,W] float tensors -> (batched NCHW fp32 padded to /32, image_sizes)."""
    mean = torch.tensor(cfg.pixel_mean).view(-1, 1, 1)
    std = torch.tensor(cfg.pixel_std).view(-1, 1, 1)
    ims = [(im.float() - mean) / std for im in images]
    sizes = [tuple(im.shape[-2:]) for im in ims]
    d = cfg.size_divisibility
    H = (max(s[0] for s in sizes) + d - 1) // d * d
    W = (max(s[1] for s in sizes) + d - 1) // d * d
    out = torch.zeros(len(ims), 3, H, W)
    for i, im in enumerate(ims):
        out[i, :, : im.shape[1], : im.shape[2]] = im
    return out, sizes


def resnet(cfg, sd, x, collect=None):
    bu = "backbone.bottom_up."
    x = _conv_bn(sd, bu + "stem.conv1", _q(x), stride=2, padding=3, relu=True)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if collect is not None:
        collect["stem"] = x
    feats = {}
    for si, nblocks in enumerate(cfg.blocks_per_stage):
        stage = si + 2
        for b in range(nblocks):
            p = f"{bu}res{stage}.{b}."
            stride = 2 if (b == 0 and stage > 2) else 1
            sc = _conv_bn(sd, p + "shortcut", x, stride=stride) if (p + "shortcut.weight") in sd else x
            out = _conv_bn(sd, p + "conv1", x, stride=stride, relu=True)  # STRIDE_IN_1X1
            out = _conv_bn(sd, p + "conv2", out, padding=1, relu=True)
            x = _conv_bn(sd, p + "conv3", out, residual=sc, relu=True)
        feats[f"res{stage}"] = x
    return feats


def fpn(cfg, sd, feats):
    res = {}
    prev = None
    for lvl in (5, 4, 3, 2):
        # engine policy: the top-down add happens in fp32 in the lateral conv's epilogue, ONE rounding to bf16 (quant_out=False here)
        lat = _conv_bn(sd, f"backbone.fpn_lateral{lvl}", feats[f"res{lvl}"], quant_out=prev is None)
        if prev is not None:
            lat = _q(lat + F.interpolate(prev, scale_factor=2, mode="nearest"))
        prev = lat
        res[f"p{lvl}"] = _conv_bn(sd, f"backbone.fpn_output{lvl}", prev, padding=1)
    res["p6"] = F.max_pool2d(res["p5"], kernel_size=1, stride=2, padding=0)
    return res


def rpn_head(sd, feats):
    rp = "proposal_generator.rpn_head."
    logits, deltas = [], []
    for name in ("p2", "p3", "p4", "p5", "p6"):
        t = _conv_bn(sd, rp + "conv", feats[name], padding=1, relu=True)
        lg = _conv_bn(sd, rp + "objectness_logits", t, quant_out=False)
        dl = _conv_bn(sd, rp + "anchor_deltas", t, quant_out=False)
        N, A, H, W = lg.shape
        logits.append(lg.permute(0, 2, 3, 1).flatten(1))
        deltas.append(dl.view(N, A, 4, H, W).permute(0, 3, 4, 1, 2).flatten(1, -2))
    return logits, deltas


def rpn_proposals(cfg, logits, deltas, feat_shapes, image_sizes, nms_mode=None, device="cuda"):
    strides = (4, 8, 16, 32, 64)
    props = []
    for lvl, (lg, dl) in enumerate(zip(logits, deltas)):
        H, W = feat_shapes[lvl]
        cell = O.cell_anchors([cfg.anchor_sizes[lvl]], cfg.anchor_ratios)
        anchors = O.grid_anchors(cell, H, W, strides[lvl])
        N = dl.shape[0]
        d = dl.numpy().reshape(-1, 4)
        a = np.broadcast_to(anchors[None], (N,) + anchors.shape).reshape(-1, 4)
        props.append(O.apply_deltas(d, a, cfg.rpn_bbox_weights).reshape(N, -1, 4))
    return O.find_top_rpn_proposals(props, [l.numpy() for l in logits], image_sizes, cfg.rpn_nms_thresh,
                                    cfg.rpn_pre_nms_topk, cfg.rpn_post_nms_topk, cfg.rpn_min_box_size,
                                    nms_mode=nms_mode, device=device)


def box_head(cfg, sd, pooled, prefix="roi_heads.box_head.", num_fc=None):
    x = _q(torch.as_tensor(pooled).flatten(1))
    for i in range(num_fc or cfg.num_fc):
        w = sd[f"{prefix}fc{i + 1}.weight"]
        x = _q(F.relu(F.linear(x, _q(w), sd[f"{prefix}fc{i + 1}.bias"])))
    return x


def box_predictor(cfg, sd, x):
    wp = "roi_heads.box_predictor."
    if cfg.output_layer == "CosineSimOutputLayers":
        xn = x / (torch.norm(x, p=2, dim=1, keepdim=True) + 1e-5)
        w = sd[wp + "cls_score.weight"]
        w = w / (torch.norm(w, p=2, dim=1, keepdim=True) + 1e-5)
        if _EMULATE_BF16:   # engine: GEMM on the bf16 x and w_hat, then the per-row scale
            scores = F.linear(x, _q(w)) * (cfg.cosine_scale / (torch.norm(x, p=2, dim=1, keepdim=True) + 1e-5))
        else:
            scores = cfg.cosine_scale * F.linear(xn, w)
    else:
        scores = F.linear(x, _q(sd[wp + "cls_score.weight"]), sd[wp + "cls_score.bias"])
    deltas = F.linear(x, _q(sd[wp + "bbox_pred.weight"]), sd[wp + "bbox_pred.bias"])
    return scores, deltas


def _roi_pooler_library(features, box_lists, output_size, sampling_ratio):
    """ROIPooler.forward (detectron2/modeling/poolers.py:191-246) through torchvision.ops.roi_align -- the multi-threaded CPU
    kernel the reference itself calls (roi_align.py:3,15).  Timing leg only; the scalar C restatement stays the checker."""
    from torchvision.ops import roi_align
    fmt = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), torch.as_tensor(b, dtype=torch.float32).reshape(-1, 4)], 1)
                     for i, b in enumerate(box_lists)])
    lvls = torch.from_numpy(O.assign_boxes_to_levels(fmt[:, 1:].numpy(), 2, 5, 224, 4))
    out = torch.zeros((len(fmt), features[0].shape[1], output_size, output_size))
    for level, f in enumerate(features):
        inds = torch.nonzero(lvls == level).flatten()
        out[inds] = roi_align(torch.as_tensor(f), fmt[inds], output_size, 1.0 / (4 << level), sampling_ratio, True)
    return out.numpy(), lvls.numpy()


def _batched_nms_library(boxes, scores, idxs, thr, mode=None, device="cuda"):
    """detectron2.layers.batched_nms (nms.py:10-29) through torchvision, as the reference calls it."""
    from torchvision.ops import boxes as box_ops
    b, s, i = torch.as_tensor(boxes, dtype=torch.float32).reshape(-1, 4), torch.as_tensor(scores, dtype=torch.float32), torch.as_tensor(idxs)
    if len(b) < 40000:
        return box_ops.batched_nms(b, s, i, thr).numpy()
    keep = torch.zeros_like(s, dtype=torch.bool)
    for c in torch.unique(i).tolist():
        m = (i == c).nonzero().view(-1)
        keep[m[box_ops.nms(b[m], s[m], thr)]] = True
    k = keep.nonzero().view(-1)
    return k[s[k].argsort(descending=True)].numpy()


def detector_forward(cfg, sd, images, out_sizes=None, nms_mode=None, device="cuda", collect=None, emulate_bf16=False,
                     library_ops=False, timings=None):
    """Full candidate-sourcing forward.  images: list of [3,H,W] tensors (BGR, 0..255).

    Returns per image dict(pred_boxes, scores, pred_classes).  ``collect`` (dict) receives intermediates.
    emulate_bf16=True restates the engine's precision policy (bf16 weights with folded FrozenBN, bf16 activation
    storage, fp32 accumulation) so that the engine can be checked layer by layer at bf16-rounding tolerance.
    library_ops=True (CPU-baseline timing leg): RoIAlign and NMS run through torchvision's multi-threaded CPU kernels, which is
    what the reference calls, instead of the scalar C restatement; ``timings`` (dict) accumulates seconds per stage.
    """
    global _EMULATE_BF16
    _EMULATE_BF16 = bool(emulate_bf16)
    saved = O.batched_nms
    if library_ops:
        O.batched_nms = _batched_nms_library
    try:
        return _detector_forward(cfg, sd, images, out_sizes, nms_mode, device, collect, library_ops, timings)
    finally:
        _EMULATE_BF16 = False
        O.batched_nms = saved


def _detector_forward(cfg, sd, images, out_sizes, nms_mode, device, collect, library_ops=False, timings=None):
    import time
    t_last = [time.perf_counter()]

    def lap(stage):
        if timings is not None:
            now = time.perf_counter()
            timings[stage] = timings.get(stage, 0.0) + now - t_last[0]
            t_last[0] = now

    with torch.no_grad():
        x, sizes = preprocess(cfg, images)
        lap("preprocess")
        res_feats = resnet(cfg, sd, x, collect)
        lap("resnet")
        if collect is not None:
            collect["features_res"] = res_feats
        feats = fpn(cfg, sd, res_feats)
        lap("fpn")
        logits, deltas = rpn_head(sd, feats)
        lap("rpn_head")
        names = ("p2", "p3", "p4", "p5", "p6")
        shapes = [tuple(feats[n].shape[-2:]) for n in names]
        props = rpn_proposals(cfg, logits, deltas, shapes, sizes, nms_mode, device)
        lap("rpn_proposals")
        if library_ops:
            pooled, lvls = _roi_pooler_library([feats[n] for n in names[:4]], [p[0] for p in props], cfg.pooler_resolution,
                                               cfg.pooler_sampling_ratio)
        else:
            pooled, lvls = O.roi_pooler([feats[n].numpy() for n in names[:4]], [p[0] for p in props],
                                        cfg.pooler_resolution, sampling_ratio=cfg.pooler_sampling_ratio)
        lap("roi_pooler")
        xh = box_head(cfg, sd, pooled)
        scores, dl = box_predictor(cfg, sd, xh)
        lap("box_head")
        probs = O.softmax_rows(scores.numpy())
        allp = np.concatenate([p[0] for p in props], 0)
        boxes = O.apply_deltas(dl.numpy(), allp, cfg.roi_bbox_weights)
        if collect is not None:
            collect.update(features={k: v for k, v in feats.items()}, rpn_logits=logits, rpn_deltas=deltas,
                           proposals=props, pooled=pooled, levels=lvls, head=xh, cls_logits=scores, box_deltas=dl,
                           probs=probs, boxes=boxes)
        results, off = [], 0
        for i, (pb, _) in enumerate(props):
            n = len(pb)
            b, s, c, r = O.fast_rcnn_inference_single_image(
                boxes[off:off + n], probs[off:off + n], sizes[i], cfg.score_thresh_test, cfg.nms_thresh_test,
                cfg.detections_per_image, nms_mode=nms_mode, device=device)
            off += n
            oh, ow = out_sizes[i] if out_sizes else sizes[i]
            b2, keep = O.detector_postprocess(b, sizes[i], oh, ow)
            results.append(dict(pred_boxes=b2[keep], scores=s[keep], pred_classes=c[keep], rows=r[keep]))
        lap("detections")
        return results


def box_corrector_forward(cfg, sd, feats, box_lists, classes, image_sizes, num_fc=3, stages=3):
    """CascadeROIHeads._forward_box_qe (cascade_rcnn.py:167-203) with BoxOnlyLayersCascade heads.

    feats: dict p2..p5 NCHW tensors; box_lists: list of [Ri,4] arrays (verified candidate boxes).
    Returns list of corrected boxes [Ri,4] in input order (one-hot score 'NMS' at thr 1.0 keeps all).
    """
    names = ("p2", "p3", "p4", "p5")
    f = [feats[n].numpy() for n in names]
    cur = [np.asarray(b, np.float32) for b in box_lists]
    with torch.no_grad():
        for k in range(stages):
            if k > 0:
                cur = [O.clip_boxes(b, s) for b, s in zip(cur, image_sizes)]  # _create_proposals_from_boxes :348-369
            pooled, _ = O.roi_pooler(f, cur, cfg.pooler_resolution, sampling_ratio=cfg.pooler_sampling_ratio)
            x = box_head(cfg, sd, pooled, prefix=f"roi_heads.box_head.{k}.", num_fc=num_fc)
            d = F.linear(x, sd[f"roi_heads.box_predictor.{k}.bbox_pred.weight"],
                         sd[f"roi_heads.box_predictor.{k}.bbox_pred.bias"]).numpy()
            allb = np.concatenate(cur, 0)
            nb = O.apply_deltas(d, allb, cfg.cascade_bbox_weights[k])
            out, off = [], 0
            for b in cur:
                out.append(nb[off:off + len(b)])
                off += len(b)
            cur = out
    # fast_rcnn_inference(score 0.1, nms 1.0, topk 1e10) on one-hot scores then restore order == clip only
    return [O.clip_boxes(b, s) for b, s in zip(cur, image_sizes)]
