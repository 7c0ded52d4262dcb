"""Parity tests proper (run on the B200 with -m gpu): every CUDA op, called through the C ABI (ctypes), against the
CPU oracle on the same seeded inputs and against the reference-generated golden fixtures.
Bars: bit-exact for indices / levels / kept sets / orders; fp32 features and boxes within 1e-3 relative (stated per test)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from lvc_b200 import ops
from lvc_b200.layers import batched_nms, nms, roi_align
from lvc_b200.testing import coco_like_boxes, distinct_scores

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


# ------------------------------------------------------------------------------------------ RoIAlign
def test_roi_align_kats():
    x = cu(np.arange(25, dtype=np.float32).reshape(1, 1, 5, 5))

    def run(roi, sr=0, aligned=True):
        return roi_align(x, cu(np.array([roi], np.float32)), 2, 1.0, sr, aligned).flatten().tolist()
    assert run([0, 1, 1, 3, 3]) == [6, 7, 11, 12]
    assert run([0, 1, 1, 3, 3], aligned=False) == [9, 10, 14, 15]
    assert run([0, 1, 1, 3, 3], sr=2) == [6, 7, 11, 12]
    assert run([0, -10, -10, -5, -5]) == [0, 0, 0, 0]
    assert run([0, 3, 3, 9, 9]) == [22, 0, 0, 0]
    assert run([0, 2, 2, 2, 2]) == [0, 0, 0, 0]
    assert run([0, 3, 3, 1, 1]) == [0, 0, 0, 0]
    assert roi_align(x, torch.zeros((0, 5), device=DEV), 2, 1.0, 0, True).shape == (0, 1, 2, 2)


@pytest.mark.parametrize("sr,aligned", [(0, True), (2, True), (0, False)])
def test_roi_align_nchw_vs_oracle(sr, aligned):
    rng = np.random.default_rng(1)
    feat = rng.standard_normal((2, 16, 50, 84)).astype(np.float32)
    b = coco_like_boxes(rng, 300)
    rois = np.concatenate([rng.integers(0, 2, (300, 1)).astype(np.float32), b], 1)
    want = O.roi_align(feat, rois, 7, 1 / 16, sr, aligned)
    got = roi_align(cu(feat), cu(rois), 7, 1 / 16, sr, aligned).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-5)   # fp32 features: 1e-3 rel (north_star); observed ~1e-6


def test_level_assignment_bit_exact():
    rng = np.random.default_rng(2)
    b = coco_like_boxes(rng, 20000)
    sides = np.array([111.99, 112, 112.01, 223.99, 224, 224.01, 447.99, 448, 448.01, 1, 2000], np.float32)
    kat = np.stack([np.zeros_like(sides), np.zeros_like(sides), sides, sides], 1)
    allb = np.concatenate([b, kat])
    got = ops.assign_boxes_to_levels(cu(allb)).cpu().numpy()
    assert np.array_equal(got, O.assign_boxes_to_levels(allb))
    assert (got[-11:] + 2).tolist() == [2, 3, 3, 3, 4, 4, 4, 5, 5, 2, 5]


def _planes_from_nchw(feats, dtype):
    return [ops.Plane.from_nchw(cu(f), dtype=dtype) for f in feats]


def test_roi_pool_fpn_golden(golden):
    """Fused pooler == reference ROIPooler.forward output (tests/golden/roi_pooler.npz), incl. edge-case boxes."""
    g = golden("roi_pooler")
    feats = [g[f"feat{i}"] for i in range(4)]
    boxes = [g["boxes0"], g["boxes1"]]
    rois = np.concatenate([np.concatenate([np.full((len(b), 1), i, np.float32), b], 1) for i, b in enumerate(boxes)])
    planes = _planes_from_nchw(feats, torch.float32)
    out, lv = ops.roi_pool_fpn(planes, (1 / 4, 1 / 8, 1 / 16, 1 / 32), cu(rois), return_levels=True)
    assert np.array_equal(lv.cpu().numpy(), g["levels"])                       # box-to-level assignment bit-exact
    np.testing.assert_allclose(out.cpu().numpy(), g["pooled"], rtol=1e-3, atol=1e-5)
    # engine layout (NHWC out) carries the same numbers
    out2 = ops.roi_pool_fpn(planes, (1 / 4, 1 / 8, 1 / 16, 1 / 32), cu(rois), out_layout=ops.OUT_NHWC)
    np.testing.assert_array_equal(out2.permute(0, 3, 1, 2).cpu().numpy(), out.cpu().numpy())
    # bf16 planes: inputs rounded to bf16, fp32 accumulate -> compare against the oracle on the rounded features
    fb = [torch.from_numpy(f).bfloat16().float().numpy() for f in feats]
    want, _ = O.roi_pooler(fb, boxes)
    planes16 = _planes_from_nchw(feats, torch.bfloat16)
    out3 = ops.roi_pool_fpn(planes16, (1 / 4, 1 / 8, 1 / 16, 1 / 32), cu(rois))
    np.testing.assert_allclose(out3.cpu().numpy(), want, rtol=1e-3, atol=1e-5)


# ------------------------------------------------------------------------------------------ NMS
def test_nms_kats():
    b = cu(np.array([[0, 0, 10, 10], [0, 0, 10, 5]], np.float32))
    s = cu(np.array([0.9, 0.8], np.float32))
    assert nms(b, s, 0.5).tolist() == [0, 1]
    assert nms(b, s, 0.49).tolist() == [0]
    far = cu(np.array([[0, 0, 1, 1], [5, 5, 6, 6], [9, 9, 10, 10]], np.float32))
    k = nms(far, cu(np.array([.1, .9, .5], np.float32)), 0.5)
    assert k.dtype == torch.int64 and k.tolist() == [1, 2, 0]
    e = nms(torch.zeros((0, 4), device=DEV), torch.zeros(0, device=DEV), 0.5)
    assert e.dtype == torch.int64 and e.numel() == 0
    d = cu(np.array([[3, 3, 3, 3], [3, 3, 3, 3], [0, 0, 4, 4]], np.float32))
    assert sorted(nms(d, cu(np.array([.3, .2, .1], np.float32)), 0.5).tolist()) == [0, 1, 2]
    same = cu(np.zeros((3, 4), np.float32) + np.array([[1, 1, 5, 5]], np.float32))
    assert nms(same, cu(np.array([.1, .9, .5], np.float32)), 0.5).tolist() == [1]


@pytest.mark.parametrize("tag", ["rpn900", "rpn4819", "head3000", "head900", "big41000"])
def test_batched_nms_golden(golden, tag):
    """Kept indices AND order identical to the reference's outputs in every regime (coordinate trick, vanilla, plain)."""
    g = golden("batched_nms")
    b, s, i, thr = cu(g[tag + "_boxes"]), cu(g[tag + "_scores"]), cu(g[tag + "_idxs"]), float(g[tag + "_thr"])
    assert np.array_equal(nms(b, s, thr).cpu().numpy(), g[tag + "_keep_plain"])
    if tag + "_keep_trick" in g:
        assert np.array_equal(batched_nms(b, s, i, thr, mode=0).cpu().numpy(), g[tag + "_keep_trick"])
        assert np.array_equal(batched_nms(b, s, i, thr, mode=1).cpu().numpy(), g[tag + "_keep_vanilla"])
        # default = what the reference does on a CUDA device: trick for n <= 25000
        assert np.array_equal(batched_nms(b, s, i, thr).cpu().numpy(), g[tag + "_keep_trick"])
    else:
        assert np.array_equal(batched_nms(b, s, i, thr).cpu().numpy(), g[tag + "_keep_d2cpu"])   # >= 40000: per-class loop


def test_batched_nms_random_vs_oracle():
    rng = np.random.default_rng(7)
    for n, ncls, thr in ((1, 3, 0.5), (63, 2, 0.3), (64, 1, 0.7), (65, 80, 0.5), (2500, 80, 0.5), (5000, 5, 0.7), (26000, 80, 0.5)):
        b = coco_like_boxes(rng, n)
        s = distinct_scores(rng, n)
        idx = rng.integers(0, ncls, n).astype(np.int64)
        for mode in (0, 1):
            got = batched_nms(cu(b), cu(s), cu(idx), thr, mode=mode).cpu().numpy()
            assert np.array_equal(got, O.batched_nms(b, s, idx, thr, mode=mode)), (n, mode)
        assert np.array_equal(batched_nms(cu(b), cu(s), cu(idx), thr).cpu().numpy(), O.batched_nms(b, s, idx, thr, device="cuda"))
    # trick mode with negative coordinates (classes may overlap after the offset, exactly like the reference)
    b = coco_like_boxes(rng, 800) - 300.0
    s = distinct_scores(rng, 800)
    idx = rng.integers(0, 4, 800).astype(np.int64)
    assert np.array_equal(batched_nms(cu(b), cu(s), cu(idx), 0.5, mode=0).cpu().numpy(), O.batched_nms(b, s, idx, 0.5, mode=0))
    # ties in score: set equality only (order on ties is implementation-defined in the reference)
    s2 = np.round(s * 20) / 20
    got = batched_nms(cu(b + 300), cu(s2), cu(idx), 0.5, mode=1).cpu().numpy()
    assert set(got.tolist()) == set(O.batched_nms(b + 300, s2, idx, 0.5, mode=1).tolist())


# ------------------------------------------------------------------------------------------ RPN post-processing
def _rpn_inputs(g):
    lv = []
    for i in range(5):
        h, w = (int(v) for v in g["shapes"][i])
        lv.append(ops.rpn_level_dense(cu(g[f"logits{i}"]), cu(g[f"deltas{i}"]), h, w, 3))
    return lv


def test_rpn_proposals_golden(golden):
    g = golden("rpn_postproc")
    sizes = cu(g["image_sizes"].astype(np.int32))
    # the golden run is the reference on CPU: 3495 candidates > 1000 -> torchvision takes the per-level (vanilla) branch
    props, logits, counts = ops.rpn_proposals(_rpn_inputs(g), sizes, (32, 64, 128, 256, 512), (0.5, 1.0, 2.0), nms_mode=1)
    for n in range(2):
        c = int(counts[n])
        assert c == len(g[f"prop_logits{n}"])
        assert np.array_equal(logits[n, :c].cpu().numpy(), g[f"prop_logits{n}"])          # same anchors, same order
        np.testing.assert_allclose(props[n, :c].cpu().numpy(), g[f"prop_boxes{n}"], rtol=1e-3, atol=1e-3)
        assert float(props[n, c:].abs().sum()) == 0.0


def test_rpn_proposals_trick_vs_oracle(golden):
    """The branch the reference takes on a CUDA device (<= 25000 candidates: coordinate trick), against the oracle."""
    g = golden("rpn_postproc")
    sizes = [tuple(int(v) for v in s) for s in g["image_sizes"]]
    props = []
    for i in range(5):
        a = g[f"anchors{i}"]
        d = g[f"deltas{i}"]
        props.append(O.apply_deltas(d.reshape(-1, 4), np.broadcast_to(a[None], (2,) + a.shape).reshape(-1, 4), (1, 1, 1, 1)).reshape(2, -1, 4))
    want = O.find_top_rpn_proposals(props, [g[f"logits{i}"] for i in range(5)], sizes, device="cuda")
    got_p, got_l, counts = ops.rpn_proposals(_rpn_inputs(g), cu(g["image_sizes"].astype(np.int32)), (32, 64, 128, 256, 512), (0.5, 1.0, 2.0))
    for n in range(2):
        c = int(counts[n])
        assert np.array_equal(got_l[n, :c].cpu().numpy(), want[n][1])
        np.testing.assert_allclose(got_p[n, :c].cpu().numpy(), want[n][0], rtol=1e-3, atol=1e-3)


def test_rpn_topk_ties_and_small_levels():
    """All-equal logits (a zero-initialised head): selection must be the k lowest indices, deterministically."""
    H, W, A = 9, 11, 3
    logits = torch.zeros((1, H * W * A), device=DEV)
    deltas = torch.zeros((1, H * W * A, 4), device=DEV)
    sizes = cu(np.array([[64, 64]], np.int32))
    p, l, c = ops.rpn_proposals([ops.rpn_level_dense(logits, deltas, H, W, A)], sizes, (32,), (0.5, 1.0, 2.0), strides=(8,),
                                pre_nms_topk=50, post_nms_topk=50, nms_thresh=1.0)
    cell = O.cell_anchors([32], (0.5, 1.0, 2.0))
    anchors = O.clip_boxes(O.grid_anchors(cell, H, W, 8)[:50], (64, 64))
    keep = O.nonempty(anchors)
    assert int(c[0]) == int(keep.sum())
    np.testing.assert_allclose(p[0, : int(c[0])].cpu().numpy(), anchors[keep], rtol=0, atol=1e-5)


def test_rpn_topk_degenerate_fallback_and_big_level():
    """A large all-equal logit map overflows the candidate list (22-bit prefix cannot separate) -> single-CTA exact select;
    and a large random level goes through the grid-wide scan path.  Both against the oracle."""
    H, W, A = 60, 70, 3
    n = H * W * A
    rng = np.random.default_rng(5)
    sizes = [(480, 560)]
    for kind in ("equal", "random", "few_distinct"):
        if kind == "equal":
            lg = np.zeros((1, n), np.float32)
        elif kind == "random":
            lg = (rng.permutation(n).astype(np.float32) / n * 10 - 5).reshape(1, n)
        else:
            lg = rng.integers(0, 3, (1, n)).astype(np.float32)          # massive ties at the k-th value
        dl = (rng.standard_normal((1, n, 4)) * 0.3).astype(np.float32)
        p, l, c = ops.rpn_proposals([ops.rpn_level_dense(cu(lg), cu(dl), H, W, A)], cu(np.array(sizes, np.int32)), (64,), (0.5, 1.0, 2.0),
                                    strides=(8,), pre_nms_topk=1000, post_nms_topk=1000, nms_thresh=0.7, nms_mode=1)
        cell = O.cell_anchors([64], (0.5, 1.0, 2.0))
        anchors = O.grid_anchors(cell, H, W, 8)
        props = O.apply_deltas(dl.reshape(-1, 4), anchors, (1, 1, 1, 1)).reshape(1, n, 4)
        want = O.find_top_rpn_proposals([props], [lg], sizes, nms_mode=O.VANILLA)
        c = int(c[0])
        assert c == len(want[0][1]), kind
        if kind == "random":
            assert np.array_equal(l[0, :c].cpu().numpy(), want[0][1])
            np.testing.assert_allclose(p[0, :c].cpu().numpy(), want[0][0], rtol=1e-3, atol=1e-3)
        else:   # tied scores: same selected set (stable index order), NMS order among equal scores follows index order too
            np.testing.assert_allclose(p[0, :c].cpu().numpy(), want[0][0], rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------------------------------ box-head post-processing
@pytest.mark.parametrize("thr", [0.05, 0.0])
@pytest.mark.parametrize("mode", [0, 1])
def test_detections_vs_oracle(golden, thr, mode):
    g = golden("fast_rcnn_inference")
    R, K = g["logits"].shape[0], g["logits"].shape[1] - 1
    ih, iw = (int(v) for v in g["image_shape"])
    out_hw = (600, 1000)
    b, s, c, r, n = ops.detections(cu(g["logits"]), cu(g["deltas"]), cu(g["props"]), torch.zeros(R, dtype=torch.int32, device=DEV),
                                   cu(np.array([[ih, iw]], np.int32)), cu(np.array([out_hw], np.int32)), K, max_rois_per_image=R,
                                   score_thresh=thr, nms_mode=mode)
    probs = O.softmax_rows(g["logits"])
    boxes = O.apply_deltas(g["deltas"], g["props"], (10, 10, 5, 5))
    wb, wsc, wc, wr = O.fast_rcnn_inference_single_image(boxes, probs, (ih, iw), thr, 0.5, 100, nms_mode=mode)
    wb2, keep = O.detector_postprocess(wb, (ih, iw), *out_hw)
    n = int(n[0])
    assert n == int(keep.sum())
    assert np.array_equal(c[0, :n].cpu().numpy(), wc[keep]) and np.array_equal(r[0, :n].cpu().numpy(), wr[keep])
    np.testing.assert_allclose(s[0, :n].cpu().numpy(), wsc[keep], rtol=1e-3, atol=1e-7)
    np.testing.assert_allclose(b[0, :n].cpu().numpy(), wb2[keep], rtol=1e-3, atol=1e-3)


def test_detections_golden_multi_image(golden):
    """Two images in one call (second = first with rows reversed); each must reproduce the reference output."""
    g = golden("fast_rcnn_inference")
    R, K = g["logits"].shape[0], g["logits"].shape[1] - 1
    ih, iw = (int(v) for v in g["image_shape"])
    lg = np.concatenate([g["logits"], g["logits"][::-1]])
    dl = np.concatenate([g["deltas"], g["deltas"][::-1]])
    pr = np.concatenate([g["props"], g["props"][::-1]])
    img = np.concatenate([np.zeros(R, np.int32), np.ones(R, np.int32)])
    sz = np.array([[ih, iw], [ih, iw]], np.int32)
    # reference CPU run: 1341 (t05) candidates > 1000 -> vanilla branch
    b, s, c, r, n = ops.detections(cu(lg), cu(dl), cu(pr), cu(img), cu(sz), cu(sz), K, max_rois_per_image=R, score_thresh=0.05, nms_mode=1)
    n0, n1 = int(n[0]), int(n[1])
    assert n0 == len(g["t05_classes"]) == n1
    assert np.array_equal(c[0, :n0].cpu().numpy(), g["t05_classes"]) and np.array_equal(r[0, :n0].cpu().numpy(), g["t05_rows"])
    np.testing.assert_allclose(s[0, :n0].cpu().numpy(), g["t05_scores"], rtol=1e-3)
    np.testing.assert_allclose(b[0, :n0].cpu().numpy(), g["t05_boxes"], rtol=1e-3, atol=1e-3)
    assert np.array_equal(c[1, :n1].cpu().numpy(), g["t05_classes"]) and np.array_equal(r[1, :n1].cpu().numpy(), R - 1 - g["t05_rows"])


# ------------------------------------------------------------------------------------------ kNN
def test_knn_golden(golden):
    g = golden("knn")
    bank = ops.KnnBank(cu(g["bank"]), cu(g["bank_cls"]))
    for k in (10, 5, 1):
        r = bank.verify(cu(g["queries"]), cu(g["query_cls"]), topk=10, knn=k)
        assert np.array_equal(np.sort(r["votes"].cpu().numpy(), 1), np.sort(g["votes"], 1))   # class multiset of the top-10
        assert np.array_equal(r["keep"].cpu().numpy(), g[f"keep_k{k}"].astype(np.uint8))


def test_knn_cdist_branch(golden):
    """QUERY_EXPAND.COSINE_SIM = False: neighbours by -cdist (run_nearest_neighbours.py:154-159) against the reference's votes and,
    tie-aware, against torch.cdist on a larger problem."""
    g = golden("knn")
    bank = ops.KnnBank(cu(g["bank"]), cu(g["bank_cls"]), cosine=False)
    r = bank.verify(cu(g["queries"]), cu(g["query_cls"]), topk=10, knn=10)
    assert np.array_equal(np.sort(r["votes"].cpu().numpy(), 1), np.sort(g["votes_cdist"], 1))
    assert np.array_equal(r["keep"].cpu().numpy(), g["keep_cdist_k10"].astype(np.uint8))
    rng = np.random.default_rng(22)
    S, D, Q = 300, 384, 2000
    cls = np.repeat(np.arange(20), S // 20).astype(np.int64)
    b = rng.standard_normal((S, D)).astype(np.float32) + 1.0
    q = rng.standard_normal((Q, D)).astype(np.float32) + 1.0
    qc = rng.integers(0, 20, Q).astype(np.int64)
    want = O.knn_cdist_torch(b, cls, q, qc)
    got = ops.KnnBank(cu(b), cu(cls), cosine=False).verify(cu(q), cu(qc), return_sim=True)
    gi, gs = got["top_idx"].cpu().numpy(), got["top_sim"].cpu().numpy()
    np.testing.assert_allclose(gs, want["top_sim"], rtol=1e-4, atol=1e-4)
    bad = gi != want["top_idx"]
    assert bad.mean() < 2e-3 and np.all(np.abs(gs[bad] - want["top_sim"][bad]) < 1e-3)
    rows_bad = bad.any(1)
    assert np.array_equal(got["keep"].cpu().numpy()[~rows_bad], want["keep"][~rows_bad])


def test_knn_vs_oracle_tie_aware():
    """Index parity against the scalar oracle; positions may differ only where the oracle's similarities tie to 1e-5."""
    rng = np.random.default_rng(21)
    S, D, Q, ncls = 600, 1024, 4000, 20
    cls = np.repeat(np.arange(ncls), S // ncls).astype(np.int64)
    means = (rng.standard_normal((ncls, D)) * 0.08).astype(np.float32)
    bank = rng.standard_normal((S, D)).astype(np.float32) + means[cls] + 3.0     # large common offset: centring matters
    qc = rng.integers(0, ncls, Q).astype(np.int64)
    q = rng.standard_normal((Q, D)).astype(np.float32) + means[qc] + 3.0
    want = O.knn_verify(bank, cls, q, qc)
    kb = ops.KnnBank(cu(bank), cu(cls))
    for path in ("simt", "tc3", "tc1"):     # exact fp32 FMA kernel; tensor-core v2 (bf16 pairs, epilogue top-k); v1 (TF32 + exact re-rank)
        got = kb.verify(cu(q), cu(qc), return_sim=True, path=path)
        gi, gs = got["top_idx"].cpu().numpy(), got["top_sim"].cpu().numpy()
        np.testing.assert_allclose(gs, want["top_sim"], rtol=1e-3, atol=2e-6)
        bad = gi != want["top_idx"]
        assert bad.mean() < 1e-3, path
        assert np.all(np.abs(gs[bad] - want["top_sim"][bad]) < 1e-5), path      # only near-ties may swap
        rows_bad = bad.any(1)
        assert np.array_equal(got["keep"].cpu().numpy()[~rows_bad], want["keep"][~rows_bad]), path
    # candidate-set overflow (all bank rows identical -> every score ties) must fall back to the exact kernel, not fail
    same = np.repeat(bank[:1], 128, 0) + np.arange(128, dtype=np.float32)[:, None] * 0     # 128 identical rows
    r_si = ops.KnnBank(cu(same), cu(cls[:128])).verify(cu(q[:64]), cu(qc[:64]), path="simt")
    for path in ("tc3", "tc1"):
        r_tc = ops.KnnBank(cu(same), cu(cls[:128])).verify(cu(q[:64]), cu(qc[:64]), path=path)
        assert torch.equal(r_tc["keep"], r_si["keep"]) and torch.equal(r_tc["votes"], r_si["votes"]), path
    # DINO-shaped descriptors (384-d, 80 x 30 bank), a ragged query count and a bank that is not a multiple of the chunk size
    S2, D2, Q2 = 2400 - 7, 384, 1000 + 13
    cls2 = (np.arange(S2) % 80).astype(np.int64)
    bank2 = rng.standard_normal((S2, D2)).astype(np.float32) + 0.5
    q2 = rng.standard_normal((Q2, D2)).astype(np.float32) + 0.5
    qc2 = rng.integers(0, 80, Q2).astype(np.int64)
    w2 = O.knn_verify(bank2, cls2, q2, qc2)
    for path in ("tc3", "tc1"):
        g2 = ops.KnnBank(cu(bank2), cu(cls2)).verify(cu(q2), cu(qc2), return_sim=True, path=path)
        gi, gs = g2["top_idx"].cpu().numpy(), g2["top_sim"].cpu().numpy()
        np.testing.assert_allclose(gs, w2["top_sim"], rtol=1e-3, atol=2e-6)
        bad = gi != w2["top_idx"]
        assert bad.mean() < 2e-3 and np.all(np.abs(gs[bad] - w2["top_sim"][bad]) < 1e-5), path
    # ragged / tiny inputs
    r1 = ops.KnnBank(cu(bank[:37]), cu(cls[:37])).verify(cu(q[:5]), cu(qc[:5]), topk=10, knn=5)
    w1 = O.knn_verify(bank[:37], cls[:37], q[:5], qc[:5], topk=10, knn=5)
    assert np.array_equal(r1["top_idx"].cpu().numpy(), w1["top_idx"]) and np.array_equal(r1["keep"].cpu().numpy(), w1["keep"])


# ---------------------------------------------------------------- "next" row f4: training-side ops
def test_roi_align_backward_golden_and_oracle(golden):
    """autograd backward of lvc_b200.layers.roi_align == the reference's (torchvision) backward; fp32 scatter-add: <= 1e-5 of scale."""
    from lvc_b200.layers import ROIAlign
    g = golden("training_ops")
    for tag in ("p3", "p2r2"):
        N, C, H, W = (int(v) for v in g[f"{tag}_shape"])
        x = torch.zeros((N, C, H, W), device=DEV, requires_grad=True)
        rois = torch.from_numpy(g[f"{tag}_rois"]).to(DEV)
        y = ROIAlign(7, float(g[f"{tag}_scale"]), int(g[f"{tag}_ratio"]), aligned=True)(x, rois)
        y.backward(torch.from_numpy(g[f"{tag}_grad_out"]).to(DEV))
        want = g[f"{tag}_grad_in"]
        assert np.abs(x.grad.cpu().numpy() - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), tag
    # larger random case against the C oracle
    rng = np.random.default_rng(3)
    N, C, H, W, R = 2, 16, 50, 84, 300
    rois = np.concatenate([rng.integers(0, N, (R, 1)).astype(np.float32), coco_like_boxes(rng, R)], 1)
    go = rng.standard_normal((R, C, 7, 7)).astype(np.float32)
    x = torch.zeros((N, C, H, W), device=DEV, requires_grad=True)
    ROIAlign(7, 1 / 16, 0, aligned=True)(x, torch.from_numpy(rois).to(DEV)).backward(torch.from_numpy(go).to(DEV))
    want = O.roi_align_backward(go, rois, (N, C, H, W), 1 / 16, 0, True)
    assert np.abs(x.grad.cpu().numpy() - want).max() <= 2e-5 * np.abs(want).max()


def test_pairwise_iou_and_matcher(golden):
    """bit-exact IoU matrix, matches (first maximum) and labels, on the reference's fixture and on an RPN-sized random case;
    matrix form (Matcher.__call__) and fused form (match_boxes) agree."""
    from lvc_b200.modeling import Matcher, pairwise_iou
    g = golden("training_ops")
    gt, props = torch.from_numpy(g["gt"]).to(DEV), torch.from_numpy(g["props"]).to(DEV)
    iou = pairwise_iou(gt, props)
    assert np.array_equal(iou.cpu().numpy(), g["iou"])
    for tag in ("rpn", "roi"):
        mt = Matcher(g[f"{tag}_thr"].tolist(), g[f"{tag}_lab"].tolist(), bool(g[f"{tag}_low"]))
        for m, l in (mt(iou), mt.match_boxes(gt, props)):
            assert np.array_equal(m.cpu().numpy(), g[f"{tag}_matches"]) and np.array_equal(l.cpu().numpy(), g[f"{tag}_labels"]), tag
    m0, l0 = Matcher([0.5], [0, 1])(torch.zeros((0, 5), device=DEV))
    assert np.array_equal(m0.cpu().numpy(), g["empty_matches"]) and np.array_equal(l0.cpu().numpy(), g["empty_labels"])
    m0, l0 = Matcher([0.5], [0, 1]).match_boxes(torch.zeros((0, 4), device=DEV), props[:5])
    assert np.array_equal(m0.cpu().numpy(), g["empty_matches"]) and np.array_equal(l0.cpu().numpy(), g["empty_labels"])
    rng = np.random.default_rng(9)
    gtb, anc = coco_like_boxes(rng, 37), coco_like_boxes(rng, 268569)
    wm, wl, _ = O.match_boxes(gtb, anc, [0.3, 0.7], [0, -1, 1], True)
    m, l = Matcher([0.3, 0.7], [0, -1, 1], True).match_boxes(torch.from_numpy(gtb).to(DEV), torch.from_numpy(anc).to(DEV))
    assert np.array_equal(m.cpu().numpy(), wm) and np.array_equal(l.cpu().numpy(), wl)


def test_rpn_losses(golden):
    """lvcb200_rpn_losses vs the reference's RPN.losses (fixture) and vs the oracle at the full anchor count; 1e-5 relative (fp32 terms)."""
    from lvc_b200.modeling import rpn_losses
    g = golden("training_ops")
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    for tag in ("l1", "sl1"):
        ls = rpn_losses(t("loss_anchors"), t("loss_logits"), t("loss_deltas"), t("loss_labels"), t("loss_gt_boxes"), 256, smooth_l1_beta=float(g[f"loss_{tag}_beta"]))
        want = g[f"loss_{tag}"] / (256 * 2)
        assert abs(float(ls["loss_rpn_cls"]) - want[0]) <= 1e-5 * want[0] and abs(float(ls["loss_rpn_loc"]) - want[1]) <= 1e-5 * want[1], tag
    rng = np.random.default_rng(5)
    N, A = 2, 268569
    anc = coco_like_boxes(rng, A)
    gtb = anc[None] + rng.uniform(-10, 10, (N, A, 4)).astype(np.float32)
    gtb[..., 2:] = np.maximum(gtb[..., 2:], gtb[..., :2] + 2.0)
    lab = rng.choice(np.array([-1, 0, 1], np.int8), size=(N, A), p=[0.9, 0.08, 0.02])
    lg = (rng.standard_normal((N, A)) * 2).astype(np.float32)
    dl = (rng.standard_normal((N, A, 4)) * 0.3).astype(np.float32)
    want = O.rpn_losses(anc, lg, dl, lab, gtb, beta=0.0) / (256 * N)
    ls = rpn_losses(*[torch.from_numpy(v).to(DEV) for v in (anc, lg, dl, lab, gtb)], 256)
    assert abs(float(ls["loss_rpn_cls"]) - want[0]) <= 1e-5 * want[0] and abs(float(ls["loss_rpn_loc"]) - want[1]) <= 1e-5 * want[1]


def test_fast_rcnn_losses(golden):
    """lvcb200_fast_rcnn_losses vs the reference's FastRCNNOutputs.losses (fixture): 1e-5 relative (fp32 terms, fp64 sums)."""
    from lvc_b200.modeling import fast_rcnn_losses
    g = golden("training_ops")
    t = lambda k: torch.from_numpy(g[k]).to(DEV)
    for tag in ("l1", "sl1"):
        ls = fast_rcnn_losses(t("frcnn_logits"), t("frcnn_deltas"), t("frcnn_gt_classes"), t("frcnn_props"), t("frcnn_gt_boxes"),
                              smooth_l1_beta=float(g[f"frcnn_{tag}_beta"]))
        want = g[f"frcnn_{tag}"] / 1024
        assert abs(float(ls["loss_cls"]) - want[0]) <= 1e-5 * want[0] and abs(float(ls["loss_box_reg"]) - want[1]) <= 1e-5 * want[1], tag
    # class-agnostic deltas: against the oracle
    dl4 = t("frcnn_deltas")[:, :4].contiguous()
    want = O.fast_rcnn_losses(g["frcnn_logits"], g["frcnn_deltas"][:, :4], g["frcnn_gt_classes"], g["frcnn_props"], g["frcnn_gt_boxes"]) / 1024
    ls = fast_rcnn_losses(t("frcnn_logits"), dl4, t("frcnn_gt_classes"), t("frcnn_props"), t("frcnn_gt_boxes"))
    assert abs(float(ls["loss_box_reg"]) - want[1]) <= 1e-5 * want[1]


# ---------------------------------------------------------------------------------------------- a15 / f2: device-side candidate filter
@pytest.mark.parametrize("tag,kw", [("score_full", dict(full=True, k_min=0.8, k_max=1.0, ar=0.0)),
                                    ("score_nofull", dict(full=False, k_min=0.5, k_max=0.9, ar=0.05))])
def test_candidate_filter_device_vs_reference(golden, tag, kw):
    """lvcb200_candidate_filter on per-image detection blocks against the reference's get_ret_anns output
    (tests/golden/candidates.npz, tools/create_coco_dataset_from_dets_all.py:129-193) and the host mirror: flags bit-exact."""
    from lvc_b200.candidates import CandidateFilter, select_candidates
    g = golden("candidates")
    ids, cat, sc, area, iarea = g["image_id"], g["category"], g["score"], g["area"], g["image_area"]
    novel = [int(c) for c in g["novel"]]
    train = {c: set(int(v) for v in g[f"train_{c}"]) for c in novel}
    w = np.sqrt(area).astype(np.float32)
    h = np.where(w > 0, area / np.maximum(w, 1e-30), 1.0).astype(np.float32)
    uniq = np.unique(ids)
    topk = max(int((ids == u).sum()) for u in uniq)
    n = len(uniq)
    boxes = np.zeros((n, topk, 4), np.float32); scores = np.zeros((n, topk), np.float32); classes = np.zeros((n, topk), np.int64)
    counts = np.zeros(n, np.int32); where = np.full((n, topk), -1, np.int64); sizes = []
    for i, u in enumerate(uniq):
        sel = np.nonzero(ids == u)[0]
        counts[i] = len(sel)
        boxes[i, :len(sel), 2], boxes[i, :len(sel), 3] = w[sel], h[sel]
        scores[i, :len(sel)], classes[i, :len(sel)], where[i, :len(sel)] = sc[sel], cat[sel], sel
        sizes.append((1.0, float(iarea[sel[0]])))                       # height * width = the image record's area
    filt = CandidateFilter(novel, kw["k_min"], kw["k_max"], kw["ar"], kw["full"], train, num_classes=15)
    flags, n_keep = filt(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(classes).cuda(),
                         torch.from_numpy(counts).cuda(), sizes, [int(u) for u in uniq])
    flags = flags.cpu().numpy()
    got = np.zeros(len(ids), np.int8)
    got[where[where >= 0]] = flags[where >= 0]
    assert np.array_equal(got, g["flags_" + tag])
    assert int((flags[where < 0] != 0).sum()) == 0                      # padding slots are dropped
    assert np.array_equal(n_keep.cpu().numpy(), (flags == 1).sum(1))
    a2 = torch.from_numpy(w.astype(np.float64) * h.astype(np.float64))
    mirror = select_candidates(torch.from_numpy(ids), torch.from_numpy(cat), torch.from_numpy(sc), a2, torch.from_numpy(iarea), train, novel,
                               kw["k_min"], kw["k_max"], kw["ar"], kw["full"])
    assert np.array_equal(got, mirror.numpy())


def test_model_candidate_flags_through_public_api():
    """model.candidate_filter: flags travel in the batch's one packed D2H and match the host mirror on the returned detections."""
    from lvc_b200.candidates import CandidateFilter, select_candidates
    from lvc_b200.config import DetectorConfig
    from lvc_b200.evaluation import CandidateCollector, inference_on_dataset
    from lvc_b200.modeling import GeneralizedRCNN
    from lvc_b200.weights import synthetic_state_dict
    cfg = DetectorConfig(depth=50, score_thresh_test=0.0)
    model = GeneralizedRCNN(cfg, synthetic_state_dict(cfg, 0), use_cuda_graph=True)
    novel = list(range(80))
    ims = [torch.rand(3, 160, 200, generator=torch.Generator().manual_seed(70 + i)) * 255 for i in range(6)]
    loader = [[{"image": ims[2 * b + j], "image_id": 500 + 2 * b + j, "height": 320, "width": 400} for j in range(2)] for b in range(3)]
    plain = [model(b) for b in loader]
    s_all = torch.cat([r["instances"].scores for b in plain for r in b])
    k_min = float(s_all.median())                                        # a threshold that splits the detections
    model.candidate_filter = CandidateFilter(novel, k_min, 1.0, ar=0.0, full=True, train_imgs={0: {501}})
    res = inference_on_dataset(model, iter(loader), CandidateCollector())   # an ITERATOR: the driver must not need len() / a list
    flat = [r["instances"] for b in (model(b) for b in loader) for r in b]
    img = torch.cat([torch.full((len(i),), 500 + k) for k, i in enumerate(flat)])
    boxes = torch.cat([i.pred_boxes.tensor for i in flat])
    area = ((boxes[:, 2] - boxes[:, 0]).double() * (boxes[:, 3] - boxes[:, 1]).double())
    want = select_candidates(img, torch.cat([i.pred_classes for i in flat]), torch.cat([i.scores for i in flat]), area,
                             torch.full((len(img),), 320.0 * 400.0, dtype=torch.float64), {0: {501}}, novel, k_min, 1.0, 0.0, True)
    got = torch.cat([i.candidate_flags for i in flat])
    assert torch.equal(got, want) and int((got == 1).sum()) > 0
    assert res["num_images"] == 6 and res["num_detections"] == len(img)
    assert res["num_candidates"] == int((want == 1).sum()) and len(res["annotations"]) == int((want != 0).sum())
    ids = [a["id"] for a in res["annotations"]]
    assert ids == sorted(ids) and ids == (want != 0).nonzero().flatten().add(1).tolist()   # loadRes' running index over all detections


# ---------------------------------------------------------------------------------------------- f4: subsample_labels
def test_subsample_labels_vs_reference_and_oracle(golden):
    """lvcb200_subsample_labels against the reference's subsample_labels / RPN._subsample_labels outputs (tests/golden/sampling.npz:
    randperm standing on the fixture's keys) -- index vectors identical, in order -- and against the oracle on a batch of full-size RPN
    label vectors (268 569 anchors) incl. heavy key ties; counts follow min(n_pos, int(ns * frac)), min(n_neg, ns - n_pos)."""
    from lvc_b200.modeling import subsample_labels, subsample_labels_batched, subsample_rpn_labels
    g = golden("sampling")
    for tag in ("rpn", "roi", "few", "ties"):
        ns, frac, bg = g[f"{tag}_args"]
        lab = torch.from_numpy(g[f"{tag}_labels"]).to(DEV)
        keys = torch.from_numpy(g[f"{tag}_keys"].astype(np.int64)).to(DEV)
        pos, neg = subsample_labels(lab, int(ns), float(frac), int(bg), keys=keys)
        assert np.array_equal(pos.cpu().numpy(), g[f"{tag}_pos"]) and np.array_equal(neg.cpu().numpy(), g[f"{tag}_neg"]), tag
        if f"{tag}_rpn_label" in g.files:
            out = subsample_rpn_labels(lab.to(torch.int8), int(ns), float(frac), keys=keys)
            assert np.array_equal(out.cpu().numpy(), g[f"{tag}_rpn_label"]), tag
    rng = np.random.default_rng(9)
    V, A = 3, 268569
    lab = rng.choice(np.array([-1, 0, 1], np.int8), size=(V, A), p=[0.3, 0.699, 0.001])
    keys = rng.integers(0, 2 ** 32, (V, A), dtype=np.uint64).astype(np.uint32)
    keys[1] = keys[1] % 50                                  # thousands of elements share the threshold key
    keys[2] = 12345                                         # all keys equal: pure index order
    pos, neg, counts = subsample_labels_batched(torch.from_numpy(lab).to(DEV), 256, 0.5, 0, keys=torch.from_numpy(keys.astype(np.int64)).to(DEV))
    for v in range(V):
        wp, wn = O.subsample_labels(lab[v], keys[v], 256, 0.5, 0)
        c = counts[v].tolist()
        assert c == [len(wp), len(wn)] and sum(c) == 256
        assert np.array_equal(pos[v, :c[0]].cpu().numpy(), wp) and np.array_equal(neg[v, :c[1]].cpu().numpy(), wn), v
        assert bool((pos[v, c[0]:] == -1).all()) and bool((neg[v, c[1]:] == -1).all())
    # keys drawn from torch's generator: a valid sample (right counts, right classes, no duplicates), different from call to call
    l1 = torch.from_numpy(lab[0].astype(np.int64)).to(DEV)
    a = subsample_labels(l1, 512, 0.25, 0)
    b = subsample_labels(l1, 512, 0.25, 0)
    assert len(a[0]) == 128 and len(a[1]) == 384 and bool((l1[a[0]] == 1).all()) and bool((l1[a[1]] == 0).all())
    assert len(torch.unique(torch.cat(a))) == 512 and not torch.equal(a[1], b[1])
    # empty vector, and no positives / no negatives at all
    e = subsample_labels(torch.zeros(0, dtype=torch.int64, device=DEV), 256, 0.5, 0)
    assert len(e[0]) == 0 and len(e[1]) == 0
    p_only = subsample_labels(torch.ones(100, dtype=torch.int64, device=DEV), 64, 0.5, 0)
    assert len(p_only[0]) == 32 and len(p_only[1]) == 0
