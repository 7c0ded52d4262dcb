"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (via oracle/ref_shim.py).

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
Every fixture stores the seeded inputs (or the seed) and the reference's outputs; the tests compare the
C / numpy oracle and -- on the GPU box -- the CUDA path against them.  Kept small (a few MB total).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from lvc_b200.config import DetectorConfig  # noqa: E402
from lvc_b200.weights import synthetic_state_dict, synthetic_corrector_head  # noqa: E402
from lvc_b200.testing import coco_like_boxes, distinct_scores  # noqa: E402


def save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


def gold_batched_nms():
    """detectron2.layers.batched_nms (nms.py:10-29) in all three regimes of the CPU reference run,
    plus forced trick / vanilla torchvision branches (the CUDA reference run takes trick up to 25000 boxes)."""
    from detectron2.layers import batched_nms, nms
    from torchvision.ops import boxes as box_ops
    rng = np.random.default_rng(11)
    out = {}
    for tag, n, ncls, thr in (("rpn900", 900, 5, 0.7), ("rpn4819", 4819, 5, 0.7), ("head3000", 3000, 80, 0.5),
                              ("head900", 900, 80, 0.5), ("big41000", 41000, 80, 0.5)):
        b = coco_like_boxes(rng, n)
        s = distinct_scores(rng, n)
        idx = rng.integers(0, ncls, n).astype(np.int64)
        tb, ts, ti = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(idx)
        out[tag + "_boxes"], out[tag + "_scores"], out[tag + "_idxs"] = b, s, idx
        out[tag + "_thr"] = np.float32(thr)
        out[tag + "_keep_d2cpu"] = batched_nms(tb, ts, ti, thr)
        if n < 40000:
            out[tag + "_keep_trick"] = box_ops._batched_nms_coordinate_trick(tb, ts, ti, thr)
            out[tag + "_keep_vanilla"] = box_ops._batched_nms_vanilla(tb, ts, ti, thr)
        out[tag + "_keep_plain"] = nms(tb, ts, thr)
    save("batched_nms", **out)


def gold_roi_pooler():
    """ROIPooler.forward (poolers.py:191-246) with the Base-RCNN-FPN settings (7x7, ratio 0, ROIAlignV2)."""
    from detectron2.modeling.poolers import ROIPooler
    from detectron2.structures import Boxes
    rng = np.random.default_rng(12)
    C, N = 8, 2
    H, W = 96, 128  # image 96*4 x 128*4 = 384 x 512
    feats = [torch.from_numpy(rng.standard_normal((N, C, H >> i, W >> i)).astype(np.float32)) for i in range(4)]
    boxes = [coco_like_boxes(rng, 150, W=512, H=384), coco_like_boxes(rng, 90, W=512, H=384)]
    # engineered edge cases on image 1: zero-size, inverted, fully outside, exact level boundaries
    edge = np.array([[10, 10, 10, 10], [300, 300, 100, 100], [-50, -50, -20, -20], [0, 0, 112, 112], [0, 0, 224, 224],
                     [0, 0, 448, 448], [0, 0, 111.99, 111.99], [5, 5, 511.9, 383.9], [500, 380, 700, 500]], np.float32)
    boxes[1] = np.concatenate([boxes[1], edge], 0)
    pooler = ROIPooler(output_size=7, scales=(1 / 4, 1 / 8, 1 / 16, 1 / 32), sampling_ratio=0, pooler_type="ROIAlignV2")
    from detectron2.modeling.poolers import assign_boxes_to_levels
    bl = [Boxes(torch.from_numpy(b)) for b in boxes]
    out = pooler(feats, bl)
    lv = assign_boxes_to_levels(bl, 2, 5, 224, 4)
    save("roi_pooler", seed=12, C=C, N=N, H=H, W=W, boxes0=boxes[0], boxes1=boxes[1], levels=lv,
         **{f"feat{i}": feats[i] for i in range(4)}, pooled=out)


def gold_rpn_postproc():
    """RPN decode + find_top_rpn_proposals (rpn.py:455-508; proposal_utils.py:13-118), CPU reference branch."""
    from detectron2.modeling.anchor_generator import DefaultAnchorGenerator
    from detectron2.modeling.box_regression import Box2BoxTransform
    from detectron2.modeling.proposal_generator.proposal_utils import find_top_rpn_proposals
    from detectron2.layers import ShapeSpec
    rng = np.random.default_rng(13)
    N = 2
    image_sizes = [(320, 416), (300, 400)]
    shapes = [(80, 104), (40, 52), (20, 26), (10, 13), (5, 7)]
    ag = DefaultAnchorGenerator(sizes=[[32], [64], [128], [256], [512]], aspect_ratios=[[0.5, 1.0, 2.0]],
                                strides=[4, 8, 16, 32, 64], offset=0.0)
    feats = [torch.zeros(N, 1, h, w) for h, w in shapes]
    anchors = ag(feats)
    tr = Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0))
    logits, deltas, props = [], [], []
    total = sum(h * w * 3 for h, w in shapes) * N
    allv = (rng.permutation(total).astype(np.float64) / total * 8 - 4).astype(np.float32)  # distinct across levels
    assert len(np.unique(allv)) == total
    cur = 0
    for (h, w), a in zip(shapes, anchors):
        n = h * w * 3
        lg = allv[cur:cur + N * n].reshape(N, n)
        cur += N * n
        dl = (rng.standard_normal((N, n, 4)) * 0.5).astype(np.float32)
        dl[0, :3] = np.array([[0, 0, 9, 9], [np.inf, 0, 0, 0], [0, 0, -30, -30]], np.float32)[: min(3, n)]
        logits.append(torch.from_numpy(lg))
        deltas.append(torch.from_numpy(dl))
        at = a.tensor.unsqueeze(0).expand(N, -1, -1).reshape(-1, 4)
        props.append(tr.apply_deltas(torch.from_numpy(dl).reshape(-1, 4), at).view(N, -1, 4))
    res = find_top_rpn_proposals(props, logits, image_sizes, 0.7, 1000, 1000, 0.0, False)
    out = dict(image_sizes=np.array(image_sizes), shapes=np.array(shapes))
    for i in range(5):
        out[f"logits{i}"], out[f"deltas{i}"] = logits[i], deltas[i]
        out[f"anchors{i}"] = anchors[i].tensor
    for n in range(N):
        out[f"prop_boxes{n}"] = res[n].proposal_boxes.tensor
        out[f"prop_logits{n}"] = res[n].objectness_logits
    save("rpn_postproc", **out)


def gold_fast_rcnn_inference():
    """FastRCNNOutputs.predict_boxes/predict_probs + fast_rcnn_inference (fast_rcnn.py:51-137,440-493)."""
    from lvc.modeling.roi_heads.fast_rcnn import fast_rcnn_inference
    from detectron2.modeling.box_regression import Box2BoxTransform
    import torch.nn.functional as F
    rng = np.random.default_rng(14)
    R, K = 300, 80
    image_shape = (800, 1333)
    props = coco_like_boxes(rng, R)
    logits = (rng.standard_normal((R, K + 1)) * 2.0).astype(np.float32)
    deltas = (rng.standard_normal((R, K * 4)) * 0.7).astype(np.float32)
    tr = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
    boxes = tr.apply_deltas(torch.from_numpy(deltas).view(R * K, 4),
                            torch.from_numpy(props).unsqueeze(1).expand(R, K, 4).reshape(-1, 4)).view(R, K * 4)
    probs = F.softmax(torch.from_numpy(logits), dim=-1)
    out = dict(props=props, logits=logits, deltas=deltas, boxes=boxes.clone(), probs=probs.clone(), image_shape=np.array(image_shape))
    for tag, thr in (("t05", 0.05), ("t00", 0.0)):
        inst, kept = fast_rcnn_inference([boxes.clone()], [probs.clone()], [image_shape], thr, 0.5, 100)  # Boxes.clip mutates its input
        out[tag + "_boxes"] = inst[0].pred_boxes.tensor
        out[tag + "_scores"] = inst[0].scores
        out[tag + "_classes"] = inst[0].pred_classes
        out[tag + "_rows"] = kept[0]
    save("fast_rcnn_inference", **out)


def gold_knn():
    """run_nearest_neighbours + get_nn_class_confirmatory (tools/run_nearest_neighbours.py:142-162,214-227)."""
    import importlib
    tool = importlib.import_module("tools.run_nearest_neighbours")
    from detectron2.structures import Instances
    rng = np.random.default_rng(15)
    S, D, ncls = 120, 256, 20
    cls = np.repeat(np.arange(ncls), S // ncls).astype(np.int64)
    means = (rng.standard_normal((ncls, D)) * 0.25).astype(np.float32)
    bank = (rng.standard_normal((S, D)).astype(np.float32) + means[cls])
    qf = []
    qcls_all, feats_all = [], []
    for i in range(40):
        q = int(rng.integers(1, 8))
        qc = rng.integers(0, ncls, q).astype(np.int64)
        f = rng.standard_normal((q, D)).astype(np.float32) + means[(qc + (rng.random(q) < 0.3)) % ncls]
        inst = Instances((10, 10))
        inst.set("crop_feats", torch.from_numpy(f))
        inst.set("gt_classes", torch.from_numpy(qc))
        qf.append({"instances": inst})
        qcls_all.append(qc)
        feats_all.append(f)
    # engineered: a zero query and a query equal to the bank mean (cosine -> 0 everywhere)
    tool.tqdm = lambda x, **k: x
    res = tool.run_nearest_neighbours(torch.from_numpy(cls), torch.from_numpy(bank), qf, cosine=True)
    out = dict(bank=bank, bank_cls=cls, queries=np.concatenate(feats_all), query_cls=np.concatenate(qcls_all),
               counts=np.array([len(x) for x in qcls_all]))
    out["votes"] = torch.cat([d["instances"].top10_shots for d in res])
    for k in (10, 5, 1):
        tool.get_nn_class_confirmatory(res, k)
        out[f"keep_k{k}"] = torch.cat([d["instances"].keep for d in res])
    # QUERY_EXPAND.COSINE_SIM = False: the -cdist branch of the same function (:154-159)
    res = tool.run_nearest_neighbours(torch.from_numpy(cls), torch.from_numpy(bank), qf, cosine=False)
    out["votes_cdist"] = torch.cat([d["instances"].top10_shots for d in res])
    tool.get_nn_class_confirmatory(res, 10)
    out["keep_cdist_k10"] = torch.cat([d["instances"].keep for d in res])
    save("knn", **out)


def _images(seed, sizes):
    ims = []
    for i, (h, w) in enumerate(sizes):
        g = torch.Generator().manual_seed(seed + i)
        ims.append(torch.rand(3, h, w, generator=g) * 255)
    return ims


def gold_e2e(tag, config_rel, depth, sizes, out_sizes, opts=(), seed=100, nfeat=2048, npooled=4096):
    """GeneralizedRCNN.inference (rcnn.py:177-322) end to end with lvc_b200.weights synthetic weights loaded
    strict=True into the reference model."""
    cfg, model = ref_shim.build_reference_model(config_rel, ["MODEL.RESNETS.DEPTH", depth] + list(opts), calibrate=False)
    dcfg = DetectorConfig.from_reference_cfg(cfg)
    sd = synthetic_state_dict(dcfg, seed=0)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    assert all("cell_anchors" in k for k in missing.missing_keys), missing.missing_keys
    ims = _images(seed, sizes)
    inputs = [{"image": im, "height": oh, "width": ow} for im, (oh, ow) in zip(ims, out_sizes)]
    cap = {}

    def hook(name):
        def f(mod, inp, out):
            cap[name] = out
        return f
    model.backbone.register_forward_hook(hook("features"))
    model.proposal_generator.register_forward_hook(hook("proposals"))
    model.roi_heads.box_pooler.register_forward_hook(hook("pooled"))
    model.roi_heads.box_head.register_forward_hook(hook("head"))
    model.roi_heads.box_predictor.register_forward_hook(hook("pred"))
    with torch.no_grad():
        res = model(inputs)
    out = dict(sizes=np.array(sizes), out_sizes=np.array(out_sizes), seed=seed, depth=depth,
               output_layer=dcfg.output_layer, score_thresh=dcfg.score_thresh_test)
    rng = np.random.default_rng(5)
    for k, v in cap["features"].items():
        flat = v.flatten()
        idx = rng.integers(0, flat.numel(), nfeat)
        out[f"feat_{k}_idx"], out[f"feat_{k}_val"] = idx, flat[idx]
        out[f"feat_{k}_absmean"] = v.abs().mean()
    props = cap["proposals"][0]
    for n, p in enumerate(props):
        out[f"prop_boxes{n}"] = p.proposal_boxes.tensor
        out[f"prop_logits{n}"] = p.objectness_logits
    pooled = cap["pooled"]
    idx = rng.integers(0, pooled.numel(), npooled)
    out["pooled_idx"], out["pooled_val"], out["pooled_shape"] = idx, pooled.flatten()[idx], np.array(pooled.shape)
    out["head_sample"] = cap["head"][:64, :64]
    out["cls_logits_sample"] = cap["pred"][0][:128]
    out["box_deltas_sample"] = cap["pred"][1][:128, :16]
    for n, r in enumerate(res):
        inst = r["instances"]
        out[f"det_boxes{n}"] = inst.pred_boxes.tensor
        out[f"det_scores{n}"] = inst.scores
        out[f"det_classes{n}"] = inst.pred_classes
    print(tag, "detections", [len(r["instances"]) for r in res], "proposals", [len(p) for p in props])
    save("e2e_" + tag, **out)


def gold_corrector():
    """Box corrector: GeneralizedRCNN + CascadeROIHeads/BoxOnlyLayersCascade (cascade_rcnn.py:167-203) driven
    through CascadeROIHeads._forward_box_qe on fixed FPN features."""
    from detectron2.structures import Boxes, Instances, ImageList
    cfg, model = ref_shim.build_reference_model(
        "COCO-detection/cascade_ubbr_R_50_FPN_ft_all_30shot_aug_ftmore.yaml",
        ["QUERY_EXPAND.ENABLED", True], calibrate=False)
    dcfg = DetectorConfig(depth=50, num_fc=3)
    sd = synthetic_corrector_head(dcfg, seed=3)
    # give bbox_pred a larger scale so that corrections are visible
    for k in list(sd):
        if "bbox_pred.weight" in k:
            sd[k] = sd[k] * 30
    miss = model.roi_heads.load_state_dict({k[len("roi_heads."):]: v for k, v in sd.items()}, strict=True)
    rng = np.random.default_rng(16)
    N, C = 1, 256
    H, W = 24, 32
    image_sizes = [(96, 128)]
    # features rounded to fp16 so the fixture stores them exactly in half the bytes
    feats = {f"p{l}": torch.from_numpy((rng.standard_normal((N, C, H >> i, W >> i)) * 0.5).astype(np.float16).astype(np.float32))
             for i, l in enumerate((2, 3, 4, 5))}
    boxes = [coco_like_boxes(rng, 48, W=128, H=96, min_side=4, max_side=120)]
    classes = [rng.integers(0, 80, len(b)).astype(np.int64) for b in boxes]
    targets = []
    for b, c, s in zip(boxes, classes, image_sizes):
        inst = Instances(s)
        inst.gt_boxes = Boxes(torch.from_numpy(b))
        inst.gt_classes = torch.from_numpy(c)
        targets.append(inst)
    with torch.no_grad():
        res, _ = model.roi_heads._forward_box_qe(feats, None, targets)
    out = dict(image_sizes=np.array(image_sizes), seed_head=3, scale=30)
    for k, v in feats.items():
        out["feat_" + k] = v.numpy().astype(np.float16)
    for n in range(N):
        out[f"boxes{n}"], out[f"classes{n}"] = boxes[n], classes[n]
        out[f"out_boxes{n}"] = res[n].pred_boxes.tensor
        out[f"out_classes{n}"] = res[n].pred_classes
    save("box_corrector", **out)


def gold_candidates():
    """get_ret_anns (tools/create_coco_dataset_from_dets_all.py:129-193) in score mode and top-K mode, with and without --full,
    on a synthetic detection set.  pycocotools is absent: the shim supplies a minimal COCO index with pycocotools' semantics."""
    import importlib
    import types
    sys.argv = ["x", "--dt-path", "none", "--K-min", "0", "--K-max", "1"]
    tool = importlib.import_module("tools.create_coco_dataset_from_dets_all")
    rng = np.random.default_rng(17)
    n_img, n = 40, 1500
    imgs = [{"id": 1000 + i, "height": int(rng.integers(300, 800)), "width": int(rng.integers(300, 1000))} for i in range(n_img)]
    novel = [3, 7, 11]
    cats = [{"id": c, "name": str(c)} for c in range(15)]
    anns = []
    for k in range(n):
        im = imgs[int(rng.integers(0, n_img))]
        w, h = float(rng.uniform(1, im["width"] * 1.05)), float(rng.uniform(1, im["height"] * 1.05))
        if k % 97 == 0:
            w = 0.0                                    # zero-area detection
        anns.append({"id": k + 1, "image_id": im["id"], "category_id": int(rng.integers(0, 15)), "bbox": [0.0, 0.0, w, h],
                     "area": w * h, "score": float(np.float32(rng.uniform(0, 1))), "iscrowd": 0})
    train_imgs = {c: set(int(v) for v in rng.choice([i["id"] for i in imgs], 6, replace=False)) for c in novel}
    out = dict(image_id=np.array([a["image_id"] for a in anns]), category=np.array([a["category_id"] for a in anns]),
               score=np.array([a["score"] for a in anns], np.float32), area=np.array([a["area"] for a in anns]),
               image_area=np.array([next(i for i in imgs if i["id"] == a["image_id"]) for a in anns] and
                                   [float(next(i for i in imgs if i["id"] == a["image_id"])["height"]) *
                                    float(next(i for i in imgs if i["id"] == a["image_id"])["width"]) for a in anns]),
               novel=np.array(novel), train_keys=np.array(novel),
               **{f"train_{c}": np.array(sorted(train_imgs[c])) for c in novel})
    for tag, kw in (("score_full", dict(top=False, full=True, K_min=0.8, K_max=1.0, ar=0.0)),
                    ("score_nofull", dict(top=False, full=False, K_min=0.5, K_max=0.9, ar=0.05)),
                    ("top_full", dict(top=True, full=True, K_min=25, K_max=3, ar=0.0))):
        import copy
        coco = tool.COCO_PK()
        coco.dataset = {"images": copy.deepcopy(imgs), "annotations": copy.deepcopy(anns), "categories": cats}
        coco.createIndex()
        args = types.SimpleNamespace(**kw)
        ret = tool.get_ret_anns(coco, train_imgs, args, novel)
        flags = np.zeros(n, np.int8)
        for a in ret:
            flags[a["id"] - 1] = 2 if a.get("ignore_qe", 0) == 1 else 1
        out["flags_" + tag] = flags
        print(tag, "kept", int((flags == 1).sum()), "ignore", int((flags == 2).sum()))
    save("candidates", **out)


def gold_crops():
    """get_crops_qe (lvc/data/utils.py:485-519), both operations, on a small image with edge-touching boxes (crop size 32)."""
    import lvc.data.utils as U
    import torch.nn.functional as F
    from detectron2.structures import Boxes, Instances
    rng = np.random.default_rng(18)
    H, W = 60, 90
    img = torch.from_numpy(rng.integers(0, 256, (1, 3, H, W)).astype(np.uint8))
    boxes = np.array([[10, 12, 30, 40], [0, 0, 89, 59], [80, 50, 89, 59], [5, 5, 6, 30], [40, 20, 70, 21], [0, 30, 20, 59], [33, 7, 34, 8]],
                     np.float32)
    insts = []
    for b in boxes:
        i = Instances((H, W))
        i.gt_boxes = Boxes(torch.from_numpy(b[None]))
        insts.append(i)
    out = dict(img=img[0], boxes=boxes.astype(np.int64))
    orig = F.interpolate
    for op in ("pad", "context"):
        U.F.interpolate = lambda x, size, mode: orig(x.float(), (32, 32), mode=mode)      # the reference hard-codes 224; shrink the fixture
        out["crops_" + op] = U.get_crops_qe(img, insts, operation=op)
    U.F.interpolate = orig
    save("crops", **out)


def gold_training_ops():
    """"Next" row f4: backward of detectron2.layers.ROIAlign through autograd (roi_align.py:63-108 -> torchvision roi_align),
    pairwise_iou (boxes.py:315-347) and Matcher (matcher.py:8-126) with the RPN and ROI-heads settings of Base-RCNN-FPN
    (IOU_THRESHOLDS [0.3, 0.7] / labels [0, -1, 1] / low-quality matches; [0.5] / [0, 1])."""
    from detectron2.layers import ROIAlign
    from detectron2.modeling.matcher import Matcher
    from detectron2.structures import Boxes, pairwise_iou
    rng = np.random.default_rng(41)
    out = {}
    for tag, (N, C, H, W, R, scale, ratio) in {"p3": (2, 6, 25, 34, 40, 0.125, 0), "p2r2": (1, 4, 40, 56, 30, 0.25, 2)}.items():
        x = torch.from_numpy(rng.standard_normal((N, C, H, W)).astype(np.float32)).requires_grad_(True)
        b = coco_like_boxes(rng, R, W=int(W / scale), H=int(H / scale), min_side=4.0, max_side=H / scale)
        b[0] = [10.0, 10.0, 10.0, 30.0]                                   # zero-width roi
        b[1] = [-30.0, -20.0, 60.0, 50.0]                                 # hangs over the top-left corner
        rois = torch.from_numpy(np.concatenate([rng.integers(0, N, (R, 1)).astype(np.float32), b], 1))
        y = ROIAlign(7, scale, ratio, aligned=True)(x, rois)
        g = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
        y.backward(g)
        out.update({f"{tag}_shape": np.array([N, C, H, W]), f"{tag}_rois": rois, f"{tag}_scale": np.float32(scale), f"{tag}_ratio": ratio,
                    f"{tag}_grad_out": g, f"{tag}_grad_in": x.grad.detach()})
    gt = coco_like_boxes(rng, 12)
    gt[5] = gt[4]                                                         # duplicate gt: argmax tie -> first index
    props = coco_like_boxes(rng, 3000)
    props[:12] = gt                                                       # exact hits (IoU 1)
    props[12:24] = gt + rng.uniform(-6, 6, (12, 4)).astype(np.float32)    # near hits
    props[30] = props[31]                                                 # duplicate proposals: low-quality ties
    iou = pairwise_iou(Boxes(torch.from_numpy(gt)), Boxes(torch.from_numpy(props)))
    out.update(gt=gt, props=props, iou=iou)
    for tag, thr, lab, low in (("rpn", [0.3, 0.7], [0, -1, 1], True), ("roi", [0.5], [0, 1], False)):
        m, l = Matcher(thr, lab, allow_low_quality_matches=low)(iou)
        out.update({f"{tag}_thr": np.array(thr, np.float32), f"{tag}_lab": np.array(lab, np.int8), f"{tag}_low": int(low),
                    f"{tag}_matches": m, f"{tag}_labels": l})
    m0, l0 = Matcher([0.5], [0, 1])(pairwise_iou(Boxes(torch.zeros(0, 4)), Boxes(torch.from_numpy(props[:5]))))
    out.update(empty_matches=m0, empty_labels=l0)
    # RPN.losses (rpn.py:328-400) through the reference's own method on a stand-in `self` (the two loss terms before the
    # loss weights; normaliser = batch_size_per_image * num_images), for beta = 0 (L1, the Base-RCNN-FPN default) and beta = 1/9
    import types
    from detectron2.modeling.box_regression import Box2BoxTransform
    from detectron2.modeling.proposal_generator.rpn import RPN
    from detectron2.utils.events import EventStorage
    Nn, per_level = 2, [900, 300, 75]
    A = sum(per_level)
    anc = coco_like_boxes(rng, A)
    gtb = anc[None] + rng.uniform(-12, 12, (Nn, A, 4)).astype(np.float32)
    gtb[..., 2:] = np.maximum(gtb[..., 2:], gtb[..., :2] + 2.0)
    labels = rng.choice(np.array([-1, 0, 1], np.int8), size=(Nn, A), p=[0.5, 0.4, 0.1])
    logits = (rng.standard_normal((Nn, A)) * 3).astype(np.float32)
    dl = (rng.standard_normal((Nn, A, 4)) * 0.5).astype(np.float32)
    out.update(loss_anchors=anc, loss_gt_boxes=gtb, loss_labels=labels, loss_logits=logits, loss_deltas=dl)
    offs = np.cumsum([0] + per_level)
    for tag, beta in (("l1", 0.0), ("sl1", 1.0 / 9)):
        me = types.SimpleNamespace(box_reg_loss_type="smooth_l1", box2box_transform=Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0)),
                                   smooth_l1_beta=beta, batch_size_per_image=256, loss_weight={})
        with EventStorage(0):
            ls = RPN.losses(me, [Boxes(torch.from_numpy(anc[offs[i]:offs[i + 1]])) for i in range(3)],
                            [torch.from_numpy(logits[:, offs[i]:offs[i + 1]]) for i in range(3)],
                            [torch.from_numpy(labels[n]) for n in range(Nn)],
                            [torch.from_numpy(dl[:, offs[i]:offs[i + 1]]) for i in range(3)],
                            [torch.from_numpy(gtb[n]) for n in range(Nn)])
        out[f"loss_{tag}"] = np.array([float(ls["loss_rpn_cls"]), float(ls["loss_rpn_loc"])], np.float64) * (256 * Nn)
        out[f"loss_{tag}_beta"] = np.float32(beta)
    # FastRCNNOutputs.losses (lvc/modeling/roi_heads/fast_rcnn.py:424-438) through the reference's own class
    from detectron2.structures import Instances
    from lvc.modeling.roi_heads.fast_rcnn import FastRCNNOutputs
    R, K = 1024, 80
    props = coco_like_boxes(rng, R)
    gtb2 = props + rng.uniform(-15, 15, (R, 4)).astype(np.float32)
    gtb2[:, 2:] = np.maximum(gtb2[:, 2:], gtb2[:, :2] + 2.0)
    gcls = np.where(rng.random(R) < 0.25, rng.integers(0, K, R), K).astype(np.int64)
    clog = (rng.standard_normal((R, K + 1)) * 2).astype(np.float32)
    pdel = (rng.standard_normal((R, 4 * K)) * 0.5).astype(np.float32)
    out.update(frcnn_props=props, frcnn_gt_boxes=gtb2, frcnn_gt_classes=gcls, frcnn_logits=clog, frcnn_deltas=pdel)
    halves = [slice(0, 600), slice(600, R)]
    insts = []
    for sl in halves:
        ins = Instances((800, 1333))
        ins.proposal_boxes, ins.gt_boxes, ins.gt_classes = Boxes(torch.from_numpy(props[sl])), Boxes(torch.from_numpy(gtb2[sl])), torch.from_numpy(gcls[sl])
        insts.append(ins)
    for tag, beta in (("l1", 0.0), ("sl1", 0.5)):
        with EventStorage(0):
            ls = FastRCNNOutputs(Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)), torch.from_numpy(clog), torch.from_numpy(pdel), insts, beta).losses()
        out[f"frcnn_{tag}"] = np.array([float(ls["loss_cls"]), float(ls["loss_box_reg"])], np.float64) * R
        out[f"frcnn_{tag}_beta"] = np.float32(beta)
    save("training_ops", **out)


def gold_sampling():
    """"Next" row f4: subsample_labels (detectron2/modeling/sampling.py:9-54) and RPN._subsample_labels (rpn.py:249-266), executed from
    the reference with ``torch.randperm(n)`` standing on per-element random keys: the i-th call returns the stable argsort of the keys
    of the class it permutes (first the positives, then the negatives -- the order sampling.py:49-50 draws them in)."""
    import detectron2.modeling.sampling as S
    from detectron2.modeling.proposal_generator.rpn import RPN
    rng = np.random.default_rng(77)
    out = {}
    real = torch.randperm
    cases = {"rpn": (20000, 256, 0.5, 0, [0.93, 0.06, 0.01]), "roi": (2100, 512, 0.25, 80, None), "few": (300, 256, 0.5, 0, [0.2, 0.75, 0.05]),
             "ties": (4000, 64, 0.5, 0, [0.5, 0.3, 0.2])}
    for tag, (n, ns, frac, bg, p) in cases.items():
        if tag == "roi":
            labels = rng.integers(0, 80, n).astype(np.int64)
            labels[rng.random(n) < 0.85] = 80
            labels[rng.random(n) < 0.02] = -1
        else:
            labels = rng.choice(np.array([-1, 0, 1], np.int64), size=n, p=p)
        keys = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        if tag == "ties":
            keys = (keys % 7).astype(np.uint32)                 # many equal keys: the stable order decides
        lt, kt = torch.from_numpy(labels), torch.from_numpy(keys.astype(np.int64))
        subsets = [torch.nonzero((lt != -1) & (lt != bg)).flatten(), torch.nonzero(lt == bg).flatten()]
        calls = []

        def fake(m, device=None, _s=subsets, _c=calls):
            sub = _s[len(_c)]
            assert m == sub.numel()
            _c.append(m)
            return torch.argsort(kt[sub], stable=True)
        torch.randperm = fake
        try:
            pos, neg = S.subsample_labels(lt, ns, frac, bg)
            if tag in ("rpn", "few"):
                del calls[:]
                me = types.SimpleNamespace(batch_size_per_image=ns, positive_fraction=frac)
                out[f"{tag}_rpn_label"] = RPN._subsample_labels(me, lt.to(torch.int8).clone()).numpy()
        finally:
            torch.randperm = real
        out.update({f"{tag}_labels": labels, f"{tag}_keys": keys, f"{tag}_args": np.array([ns, frac, bg], np.float64),
                    f"{tag}_pos": pos.numpy(), f"{tag}_neg": neg.numpy()})
    save("sampling", **out)


def main():
    which = sys.argv[1:] or ["nms", "pooler", "rpn", "frcnn", "knn", "e2e", "corrector", "cand", "crops", "train"]
    if "train" in which:
        gold_training_ops()
    if "sampling" in which:
        gold_sampling()
    if "cand" in which:
        gold_candidates()
    if "crops" in which:
        gold_crops()
    if "nms" in which:
        gold_batched_nms()
    if "pooler" in which:
        gold_roi_pooler()
    if "rpn" in which:
        gold_rpn_postproc()
    if "frcnn" in which:
        gold_fast_rcnn_inference()
    if "knn" in which:
        gold_knn()
    if "corrector" in which:
        gold_corrector()
    if "e2e" in which:
        # config #1 family: Base-RCNN-FPN R50 (FastRCNNOutputLayers), two ragged images
        gold_e2e("r50_base", "Base-RCNN-FPN.yaml", 50, [(320, 416), (300, 400)], [(480, 624), (300, 400)])
        # candidate-sourcing config (CosineSimOutputLayers), R101, one image
        gold_e2e("r101_cosine", "COCO-detection/faster_rcnn_R_50_FPN_ft_all_30shot_aug_ftmore_dropout.yaml", 101,
                 [(256, 320)], [(256, 320)])
    if "e2e_b8" in which:
        # BASELINE config #2 itself: R101-FPN candidate-sourcing model, batch 8 of 3x800x1333 (the bench's images: seeds 0..7)
        gold_e2e("r101_b8", "COCO-detection/faster_rcnn_R_50_FPN_ft_all_30shot_aug_ftmore_dropout.yaml", 101,
                 [(800, 1333)] * 8, [(800, 1333)] * 8, seed=0, nfeat=32768, npooled=32768)


if __name__ == "__main__":
    main()
