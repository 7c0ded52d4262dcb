"""DetectorEngine end to end on the B200 (-m gpu): against the bf16-emulating oracle (tight: same precision policy) and
against the reference's fp32 golden outputs (loose: bf16 activations through 50-101 layers).
Tolerances (stated, see DESIGN.md 'Precision'): features 3e-2 relative L2 vs the bf16-emulating oracle; proposal /
detection agreement measured as matched fractions because near-tied scores legitimately reorder under bf16 rounding."""
import numpy as np
import pytest
import torch

from lvc_b200.config import DetectorConfig
from lvc_b200.modeling import DetectorEngine, GeneralizedRCNN
from lvc_b200.weights import synthetic_state_dict
from oracle import model as OM

pytestmark = pytest.mark.gpu


def _images(seed, sizes):
    return [torch.rand(3, h, w, generator=torch.Generator().manual_seed(seed + i)) * 255 for i, (h, w) in enumerate(sizes)]


def _iou(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None] - inter + 1e-12)


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def _rpn_head_arrays(eng, n):
    """Engine's fused fp32 [rows,16] RPN head buffers -> reference layout logits [N,HWA], deltas [N,HWA,4]."""
    logits, deltas = [], []
    for lv in eng.debug["rpn_levels"]:
        H, W, A = lv["H"], lv["W"], lv["A"]
        h = lv["logits"].view(n, H + 2, W + 2, 16)[:, 1:H + 1, 1:W + 1]
        logits.append(h[..., :A].reshape(n, -1).cpu().numpy())
        deltas.append(h[..., A:5 * A].reshape(n, -1, 4).cpu().numpy())
    return logits, deltas


@pytest.mark.parametrize("depth,layer,sizes", [(50, "FastRCNNOutputLayers", [(320, 416), (300, 400)]),
                                               (101, "CosineSimOutputLayers", [(256, 320)])])
def test_engine_stagewise_vs_oracle(depth, layer, sizes):
    """Stage-wise parity: every stage of the engine is checked against the oracle evaluated ON THE ENGINE'S OWN INPUTS to
    that stage, so that selection steps (top-k, NMS) are compared bit-exactly instead of through accumulated bf16 noise."""
    from oracle import oracle as O
    cfg = DetectorConfig(depth=depth, output_layer=layer)
    sd = synthetic_state_dict(cfg, 0)
    ims = _images(100, sizes)
    n = len(ims)
    eng = DetectorEngine(cfg, sd)
    eng.debug = {}
    boxes, scores, classes, rows, counts = eng.run([im.cuda() for im in ims])
    torch.cuda.synchronize()
    dbg = eng.debug
    # (1) dense backbone + FPN vs the bf16-emulating oracle (same precision policy): bf16-rounding-level agreement
    col = {}
    OM.detector_forward(cfg, sd, ims, device="cuda", collect=col, emulate_bf16=True)
    assert _rel(dbg["stem_pool"].to_nchw().cpu(), col["stem"]) < 1e-2, "stem (7x7/2 conv as s2d shift-GEMM) + maxpool"
    for l in (2, 3, 4, 5):
        assert _rel(dbg["feats"][l].to_nchw().cpu(), col["features_res"][f"res{l}"]) < 3e-2, f"res{l}"
    for l in (2, 3, 4, 5, 6):
        assert _rel(dbg["pyramid"][l].to_nchw().cpu(), col["features"][f"p{l}"]) < 3e-2, f"p{l}"
    # (2) RPN head on the engine's pyramid
    pyr = {f"p{l}": dbg["pyramid"][l].to_nchw().cpu() for l in (2, 3, 4, 5, 6)}
    OM._EMULATE_BF16 = True
    try:
        ref_logits, ref_deltas = OM.rpn_head(sd, pyr)
    finally:
        OM._EMULATE_BF16 = False
    e_logits, e_deltas = _rpn_head_arrays(eng, n)
    for i in range(5):
        assert _rel(torch.from_numpy(e_logits[i]), ref_logits[i]) < 1e-2, f"rpn logits level {i}"
        assert _rel(torch.from_numpy(e_deltas[i]), ref_deltas[i]) < 1e-2, f"rpn deltas level {i}"
    # (3) proposals from the engine's own logits / deltas: selection and order bit-exact, boxes 1e-3
    shapes = [(lv["H"], lv["W"]) for lv in dbg["rpn_levels"]]
    want = OM.rpn_proposals(cfg, [torch.from_numpy(a) for a in e_logits], [torch.from_numpy(a) for a in e_deltas], shapes, sizes, device="cuda")
    for i in range(n):
        c = int(dbg["prop_counts"][i])
        assert c == len(want[i][1])
        assert np.array_equal(dbg["prop_logits"][i, :c].cpu().numpy(), want[i][1])
        np.testing.assert_allclose(dbg["props"][i, :c].cpu().numpy(), want[i][0], rtol=1e-3, atol=1e-3)
    # (4) pooler on the engine's pyramid + proposals (bf16 output: one rounding step)
    P = dbg["props"].shape[1]
    props = [dbg["props"][i].cpu().numpy() for i in range(n)]           # all P slots, padded rows are zero boxes
    ref_pooled, _ = O.roi_pooler([pyr[f"p{l}"].numpy() for l in (2, 3, 4, 5)], props, cfg.pooler_resolution)
    got = dbg["pooled"].float().permute(0, 3, 1, 2).cpu().numpy()
    np.testing.assert_allclose(got, ref_pooled, rtol=2 ** -7, atol=1e-3 * np.abs(ref_pooled).max())
    # (5) box head + predictor on the engine's pooled features
    OM._EMULATE_BF16 = True
    try:
        xh = OM.box_head(cfg, sd, torch.from_numpy(got))
        ref_scores, ref_deltas2 = OM.box_predictor(cfg, sd, dbg["head"].float().cpu())
    finally:
        OM._EMULATE_BF16 = False
    assert _rel(dbg["head"].float().cpu(), xh) < 1e-2
    K = cfg.num_classes
    pred = dbg["pred"].cpu()
    e_scores = pred[:, :K + 1]
    if layer == "CosineSimOutputLayers":
        e_scores = e_scores * (cfg.cosine_scale / (dbg["head"].float().cpu().norm(dim=1, keepdim=True) + 1e-5))
    assert _rel(e_scores, ref_scores) < 1e-2 and _rel(pred[:, eng.cls_cols:], ref_deltas2) < 1e-2
    # (6) detections from the engine's own logits / deltas / proposals: classes, rows, order bit-exact
    probs = O.softmax_rows(e_scores.numpy())
    dec = O.apply_deltas(pred[:, eng.cls_cols:].numpy(), np.concatenate(props), cfg.roi_bbox_weights)
    for i in range(n):
        c = int(dbg["prop_counts"][i])
        sl = slice(i * P, i * P + c)
        wb, ws, wc, wr = O.fast_rcnn_inference_single_image(dec[sl], probs[sl], sizes[i], cfg.score_thresh_test, cfg.nms_thresh_test,
                                                            cfg.detections_per_image, device="cuda")
        wb, keep = O.detector_postprocess(wb, sizes[i], *sizes[i])
        k = int(counts[i])
        assert k == int(keep.sum())
        # scores within 1e-6 of each other may swap (GPU expf vs libm differ in the last ulp): compare as sets there
        if not np.array_equal(classes[i, :k].cpu().numpy(), wc[keep]):
            assert sorted(zip(classes[i, :k].tolist(), rows[i, :k].tolist())) == sorted(zip(wc[keep].tolist(), wr[keep].tolist()))
        np.testing.assert_allclose(scores[i, :k].cpu().numpy(), ws[keep], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(np.sort(boxes[i, :k].cpu().numpy(), 0), np.sort(wb[keep], 0), rtol=1e-3, atol=1e-2)


def _run_debug(cfg, sd, ims, precision, outs=None):
    eng = DetectorEngine(cfg, sd, precision=precision)
    eng.debug = {}
    res = eng.run([im.cuda() for im in ims], outs)
    torch.cuda.synchronize()
    return eng, res


@pytest.mark.parametrize("depth,layer,sizes", [(50, "FastRCNNOutputLayers", [(320, 416), (300, 400)]),
                                               (101, "CosineSimOutputLayers", [(256, 320)])])
def test_strict_engine_vs_fp32_oracle(depth, layer, sizes):
    """STRICT mode against the plain fp32 oracle (no bf16 emulation): BASELINE.json's contract -- features / logits / scores within
    1e-3 relative, index work identical wherever the fp32 inputs are not within rounding noise of a tie."""
    cfg = DetectorConfig(depth=depth, output_layer=layer)
    sd = synthetic_state_dict(cfg, 0, randomize_bn=True)
    ims = _images(100, sizes)
    n = len(ims)
    eng, (boxes, scores, classes, rows, counts) = _run_debug(cfg, sd, ims, "strict")
    dbg = eng.debug
    col = {}
    ref = OM.detector_forward(cfg, sd, ims, device="cuda", collect=col)
    rels = {"stem": _rel(dbg["stem_pool"].to_nchw().cpu(), col["stem"])}
    for l in (2, 3, 4, 5):
        rels[f"res{l}"] = _rel(dbg["feats"][l].to_nchw().cpu(), col["features_res"][f"res{l}"])
    for l in (2, 3, 4, 5, 6):
        rels[f"p{l}"] = _rel(dbg["pyramid"][l].to_nchw().cpu(), col["features"][f"p{l}"])
    e_logits, e_deltas = _rpn_head_arrays(eng, n)
    for i in range(5):
        rels[f"rpn_logits{i}"] = _rel(torch.from_numpy(e_logits[i]), col["rpn_logits"][i])
        rels[f"rpn_deltas{i}"] = _rel(torch.from_numpy(e_deltas[i]), col["rpn_deltas"][i])
    print("strict rel-L2 vs fp32 oracle:", {k: f"{v:.2e}" for k, v in rels.items()})
    assert max(rels.values()) < 1e-3, rels
    # proposals: the reference's boxes are all there (within 0.01 px); the ORDER may differ where fp32 logits tie within rounding noise
    # (random-init RPN weights of std 0.01 put hundreds of logits within 1e-6 of each other)
    same = tot = 0
    for i in range(n):
        c = int(dbg["prop_counts"][i])
        want = col["proposals"][i][0]
        assert abs(c - len(want)) <= 2
        d = np.abs(want[:, None, :] - dbg["props"][i, :c].cpu().numpy()[None, :, :]).max(-1)
        same += int((d.min(1) < 1e-2).sum())
        tot += len(want)
    print(f"strict proposals reproduced: {same}/{tot}")
    assert same >= 0.95 * tot
    # detections
    ok = tot_d = 0
    for i in range(n):
        k = int(counts[i])
        gb, gs, gc = ref[i]["pred_boxes"], ref[i]["scores"], ref[i]["pred_classes"]
        tot_d += len(gs)
        if k == 0 or len(gs) == 0:
            continue
        iou = _iou(gb, boxes[i, :k].cpu().numpy())
        j = iou.argmax(1)
        good = (iou.max(1) > 0.95) & (classes[i, :k].cpu().numpy()[j] == gc) & (np.abs(scores[i, :k].cpu().numpy()[j] - gs) < 1e-3 * np.maximum(gs, 1e-3) + 1e-5)
        ok += int(good.sum())
    print(f"strict detections reproduced within 1e-3: {ok}/{tot_d}")
    assert ok >= 0.97 * tot_d


@pytest.mark.parametrize("precision", ["strict", "bf16"])
@pytest.mark.parametrize("name", ["e2e_r101_cosine", "e2e_r101_b8"])
def test_engine_vs_reference_golden_e2e(golden, name, precision):
    """Both engine modes against the UNMODIFIED reference's fp32 outputs (oracle/make_golden.py): the candidate-sourcing config on one
    small image, and BASELINE config #2 itself (R101-FPN, batch 8 x 3x800x1333, the bench's images).  The strict mode must meet
    north_star's 1e-3; the bf16 throughput mode is measured and held to the tolerance DESIGN.md states for it."""
    from lvc_b200.testing import e2e_parity_metrics
    g = golden(name)
    cfg = DetectorConfig(depth=int(g["depth"]), output_layer=str(g["output_layer"]), score_thresh_test=float(g["score_thresh"]))
    sizes = [tuple(int(v) for v in s) for s in g["sizes"]]
    outs = [tuple(int(v) for v in s) for s in g["out_sizes"]]
    ims = _images(int(g["seed"]), sizes)
    eng, (boxes, scores, classes, rows, counts) = _run_debug(cfg, synthetic_state_dict(cfg, 0), ims, precision, outs)
    m = e2e_parity_metrics(g, eng.debug, boxes, scores, classes, counts)
    print(f"{name} [{precision}]:", m)
    worst = max(m["features_rel_l2"].values())
    if precision == "strict":
        assert worst < 1e-3 and m["proposals_reproduced"] > 0.98 and m["detections_reproduced"] > 0.97
        assert m["max_score_delta_matched"] < 1e-3
    else:
        # bf16 storage through 101 layers: ~1 % feature error; with random-init heads (logit spread ~1e-2) that is enough to pick
        # different near-tied anchors, so boxes are compared by IoU and only a majority of the reference's detections reappears
        assert worst < 2e-2 and m["detections_reproduced"] > 0.4


def test_strict_cuda_graph_and_api():
    """precision='strict' through the public model API and a CUDA-graph replay: identical to the eager strict run."""
    cfg = DetectorConfig(depth=50)
    sd = synthetic_state_dict(cfg, 0)
    ims = [im.cuda() for im in _images(7, [(192, 256), (192, 256)])]
    eager = DetectorEngine(cfg, sd, precision="strict").run(ims)
    eng = DetectorEngine(cfg, sd, use_cuda_graph=True, precision="strict")
    for _ in range(3):
        out = eng.run(ims)
    torch.cuda.synchronize()
    for a, b in zip(eager, out):
        assert torch.equal(a, b)
    model = GeneralizedRCNN(cfg, sd, precision="strict", use_cuda_graph=False)
    res = model([{"image": im.cpu()} for im in ims])
    assert len(res) == 2 and len(res[0]["instances"]) == int(eager[4][0])


def test_shape_lru_and_back_to_back_runs():
    """(a) shape-keyed state is LRU-bounded -- cycling through more input shapes than max_shapes keeps working and
    reproduces the first shape's results after it was evicted; (b) back-to-back run() calls with different images and no host
    sync in between keep their own metadata (pinned staging ring)."""
    cfg = DetectorConfig(depth=50)
    sd = synthetic_state_dict(cfg, 0)
    eng = DetectorEngine(cfg, sd, use_cuda_graph=True, max_shapes=2)
    shapes = [(128, 160), (160, 192), (192, 224)]
    ims = {s: [im.cuda() for im in _images(40 + i, [s])] for i, s in enumerate(shapes)}
    first = [t.clone() for t in eng.run(ims[shapes[0]])]
    for _ in range(2):
        for s in shapes:
            for _ in range(3):
                eng.run(ims[s])
    assert len(eng._states) == 2
    again = eng.run(ims[shapes[0]])
    torch.cuda.synchronize()
    for a, b in zip(first, again):
        assert torch.equal(a, b)
    eng2 = DetectorEngine(cfg, sd)
    a_im, b_im = _images(60, [(128, 160)])[0].cuda(), _images(61, [(100, 150)])[0].cuda()
    want_a = [t.clone() for t in eng2.run([a_im])]
    want_b = [t.clone() for t in eng2.run([b_im])]
    torch.cuda.synchronize()
    got = []
    for _ in range(6):                      # no synchronisation between calls
        got.append([t.clone() for t in eng2.run([a_im])])
        got.append([t.clone() for t in eng2.run([b_im])])
    torch.cuda.synchronize()
    for k, r in enumerate(got):
        for x, y in zip(r, want_a if k % 2 == 0 else want_b):
            assert torch.equal(x, y)


def test_model_api_vs_golden_fp32(golden):
    """Public API (list[dict] in, list[dict{'instances'}] out) against the reference's fp32 outputs on config #1's family."""
    g = golden("e2e_r50_base")
    cfg = DetectorConfig(depth=50)
    sizes = [tuple(int(v) for v in s) for s in g["sizes"]]
    outs = [tuple(int(v) for v in s) for s in g["out_sizes"]]
    model = GeneralizedRCNN(cfg, synthetic_state_dict(cfg, 0), use_cuda_graph=False)
    ims = _images(int(g["seed"]), sizes)
    res = model([{"image": im, "height": o[0], "width": o[1]} for im, o in zip(ims, outs)])
    for n, r in enumerate(res):
        inst = r["instances"]
        assert inst.image_size == outs[n]
        gb, gs, gc = g[f"det_boxes{n}"], g[f"det_scores{n}"], g[f"det_classes{n}"]
        assert abs(len(inst) - len(gs)) <= 0.2 * len(gs) + 3
        mm = _iou(gb, inst.pred_boxes.tensor.numpy())
        j = mm.argmax(1)
        ok = (mm.max(1) > 0.8) & (inst.pred_classes.numpy()[j] == gc) & (np.abs(inst.scores.numpy()[j] - gs) < 0.05)
        assert ok.mean() > 0.6, f"only {ok.mean():.2f} of the reference detections reproduced under bf16"


def test_cuda_graph_replay_is_deterministic():
    cfg = DetectorConfig(depth=50)
    sd = synthetic_state_dict(cfg, 0)
    ims = [im.cuda() for im in _images(7, [(192, 256), (192, 256)])]
    eager = DetectorEngine(cfg, sd).run(ims)
    eng = DetectorEngine(cfg, sd, use_cuda_graph=True)
    for _ in range(3):
        out = eng.run(ims)
    torch.cuda.synchronize()
    for a, b in zip(eager, out):
        assert torch.equal(a, b)


def test_uint8_images_match_float_images():
    """DatasetMapper yields uint8 BGR images; the stem kernel reads them directly (4x less H2D traffic)."""
    cfg = DetectorConfig(depth=50)
    sd = synthetic_state_dict(cfg, 0)
    u8 = [(torch.rand(3, 200, 264, generator=torch.Generator().manual_seed(3)) * 255).to(torch.uint8).cuda()]
    eng = DetectorEngine(cfg, sd)
    a = eng.run(u8)
    b = eng.run([u8[0].float()])
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_box_corrector_vs_golden_and_oracle(golden):
    """a17: 3-stage box corrector against the reference's CascadeROIHeads._forward_box_qe output (fp32 golden; bf16 engine ->
    boxes within 0.5 px) and against the bf16-emulating oracle fed the same rounded features (tight)."""
    from lvc_b200 import ops
    from lvc_b200.modeling import BoxCorrectorHead
    from lvc_b200.weights import synthetic_corrector_head
    g = golden("box_corrector")
    cfg = DetectorConfig(depth=50, num_fc=3)
    sd = synthetic_corrector_head(cfg, seed=int(g["seed_head"]))
    for k in list(sd):
        if "bbox_pred.weight" in k:
            sd[k] = sd[k] * float(g["scale"])
    feats = {l: torch.from_numpy(g[f"feat_p{l}"].astype(np.float32)) for l in (2, 3, 4, 5)}
    planes = [ops.Plane.from_nchw(feats[l].cuda()) for l in (2, 3, 4, 5)]
    sizes = [tuple(int(v) for v in s) for s in g["image_sizes"]]
    head = BoxCorrectorHead(cfg, sd)
    out = head(planes, [torch.from_numpy(g["boxes0"]).cuda()], sizes)
    got = out[0].cpu().numpy()
    assert np.abs(got - g["boxes0"]).max() > 0.5                       # the head moved the boxes
    assert np.abs(got - g["out_boxes0"]).max() < 0.5                   # vs the reference's fp32 result (bf16 features / weights)
    OM._EMULATE_BF16 = True
    try:
        ref = OM.box_corrector_forward(cfg, sd, {f"p{l}": feats[l].bfloat16().float() for l in (2, 3, 4, 5)}, [g["boxes0"]],
                                       [g["classes0"]], sizes)
    finally:
        OM._EMULATE_BF16 = False
    assert np.abs(got - ref[0]).max() < 0.1


def test_inference_stream_equals_blocking_calls():
    cfg = DetectorConfig(depth=50)
    sd = synthetic_state_dict(cfg, 0)
    model = GeneralizedRCNN(cfg, sd, use_cuda_graph=True)
    batches = [[{"image": (im * 1).to(torch.uint8), "height": 200, "width": 260} for im in _images(20 + 2 * b, [(192, 256), (192, 256)])]
               for b in range(4)]
    want = [model(b) for b in batches]
    got = list(model.inference_stream(batches))
    assert len(got) == len(want)
    for w, g in zip(want, got):
        for a, b in zip(w, g):
            assert torch.equal(a["instances"].pred_boxes.tensor, b["instances"].pred_boxes.tensor)
            assert torch.equal(a["instances"].scores, b["instances"].scores)
            assert torch.equal(a["instances"].pred_classes, b["instances"].pred_classes)


def test_edge_cases_tiny_image_and_no_detections():
    """Smallest legal input (one 3x40x56 image -> padded 64x64, p6 = 1x1, fewer anchors than PRE_NMS_TOPK on every level) and a
    score threshold nothing passes: the pipeline must return well-formed empty results, like the reference."""
    cfg = DetectorConfig(depth=50, score_thresh_test=0.999)
    sd = synthetic_state_dict(cfg, 0)
    model = GeneralizedRCNN(cfg, sd, use_cuda_graph=False)
    im = _images(5, [(40, 56)])[0]
    res = model([{"image": im, "height": 80, "width": 112}])
    inst = res[0]["instances"]
    assert len(inst) == 0 and inst.image_size == (80, 112) and inst.pred_boxes.tensor.shape == (0, 4)
    cfg2 = DetectorConfig(depth=50, score_thresh_test=0.0)
    eng = DetectorEngine(cfg2, sd)
    eng.debug = {}
    boxes, scores, classes, rows, counts = eng.run([im.cuda()])
    torch.cuda.synchronize()
    c = int(eng.debug["prop_counts"][0])
    assert 0 < c <= 1000 and int(counts[0]) > 0
    col = {}
    ref = OM.detector_forward(cfg2, sd, [im], device="cuda", collect=col, emulate_bf16=True)
    assert abs(c - len(col["proposals"][0][0])) <= 0.1 * c + 3


def test_proposal_network_and_regonly_models():
    """The other two meta-architectures of the mining path through their public API (rcnn.py:336-488).
    ProposalNetwork: the proposals are those the detector's own RPN stage produces (bit-equal), rescaled like detector_postprocess.
    GeneralizedRCNNRegOnly: gt_boxes regressed through the three cascade heads == BoxCorrectorHead on the same features, and within
    0.2 px of the oracle's box corrector evaluated on those features."""
    from lvc_b200.modeling import BoxCorrectorHead, GeneralizedRCNNRegOnly, ProposalNetwork
    from lvc_b200.structures import Boxes, Instances
    from lvc_b200.weights import synthetic_corrector_head
    cfg = DetectorConfig(depth=50)
    sd = synthetic_state_dict(cfg, 0)
    ims = _images(31, [(192, 256), (160, 224)])
    # --- ProposalNetwork
    pn = ProposalNetwork(cfg, {k: v for k, v in sd.items() if not k.startswith("roi_heads.")}, use_cuda_graph=False)
    out = pn([{"image": im, "height": 2 * im.shape[1], "width": 2 * im.shape[2]} for im in ims])
    eng = DetectorEngine(cfg, sd)
    eng.debug = {}
    eng.run([im.cuda() for im in ims])
    torch.cuda.synchronize()
    for i, o in enumerate(out):
        c = int(eng.debug["prop_counts"][i])
        want = eng.debug["props"][i, :c].clone()
        want[:, 0::2] *= 2.0
        want[:, 1::2] *= 2.0
        p = o["proposals"]
        assert p.image_size == (2 * ims[i].shape[1], 2 * ims[i].shape[2]) and len(p) == c
        assert torch.equal(p.proposal_boxes.tensor, want) and torch.equal(p.objectness_logits, eng.debug["prop_logits"][i, :c])
    raw, sizes = pn([{"image": im} for im in ims], no_post=True)
    assert len(raw) == 2 and sizes == [tuple(im.shape[-2:]) for im in ims]
    # --- GeneralizedRCNNRegOnly
    ccfg = DetectorConfig(depth=50, num_fc=3)
    hsd = synthetic_corrector_head(ccfg, 3)
    for k in list(hsd):
        if "bbox_pred.weight" in k:
            hsd[k] = hsd[k] * 30                      # visible corrections
    full = {k: v for k, v in sd.items() if not k.startswith("roi_heads.")}
    full.update(hsd)
    model = GeneralizedRCNNRegOnly(ccfg, full)
    rng = np.random.default_rng(5)
    from lvc_b200.testing import coco_like_boxes
    gts, inputs = [], []
    for im in ims:
        b = coco_like_boxes(rng, 20, W=im.shape[2], H=im.shape[1], min_side=8, max_side=150)
        inst = Instances(tuple(im.shape[-2:]))
        inst.gt_boxes = Boxes(torch.from_numpy(b))
        inst.gt_classes = torch.from_numpy(rng.integers(0, 80, 20))
        gts.append(b)
        inputs.append({"image": im, "instances": inst, "height": im.shape[1], "width": im.shape[2]})
    res = model(inputs)
    assert len(res) == 2 and "image" not in res[0]
    pyramid, _ = model.engine.run_features([im.cuda() for im in ims])
    head = BoxCorrectorHead(ccfg, hsd)
    want = head([pyramid[l] for l in (2, 3, 4, 5)], [torch.from_numpy(g).cuda() for g in gts], [tuple(im.shape[-2:]) for im in ims])
    # the oracle's box corrector on the ENGINE's features (bf16-emulating arithmetic): isolates the three cascade stages; the backbone
    # itself is covered by the stage-wise tests above (with bbox_pred scaled x30 a 1 % feature difference would move boxes by pixels)
    feats = {f"p{l}": pyramid[l].to_nchw().cpu() for l in (2, 3, 4, 5)}
    OM._EMULATE_BF16 = True
    try:
        ref = OM.box_corrector_forward(ccfg, hsd, feats, gts, [np.zeros(20, np.int64)] * 2, [tuple(im.shape[-2:]) for im in ims])
    finally:
        OM._EMULATE_BF16 = False
    for i, r in enumerate(res):
        got = r["instances"].pred_boxes.tensor
        keep = torch.ones(len(want[i]), dtype=torch.bool, device=want[i].device)      # the reference filters on the (non-empty) gt boxes
        assert torch.equal(got.cuda(), want[i][keep]) and torch.equal(r["instances"].pred_classes, inputs[i]["instances"].gt_classes)
        assert np.abs(got.numpy() - gts[i][keep.cpu().numpy()]).max() > 0.5          # the heads moved the boxes
        move = np.abs(got.numpy() - gts[i][keep.cpu().numpy()]).max()
        diff = np.abs(got.numpy() - ref[i][keep.cpu().numpy()])
        print(f"RegOnly image {i}: boxes moved by up to {move:.1f} px; vs the bf16-emulating oracle on the same features: max {diff.max():.3f} px, median {np.median(diff):.3f} px")
        assert diff.max() < max(1.0, 0.1 * move) and np.median(diff) < 0.25     # three cascade stages with bbox_pred scaled x30 on real backbone features
