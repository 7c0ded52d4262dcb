"""PseudoLabelMiner (Label -> Verify -> Correct on the device, lvc_b200/mining.py) against the stage-by-stage composition the
reference runs through JSON files: every hand-over is checked against the oracle / the reference's conversion chain (-m gpu)."""
import numpy as np
import pytest
import torch

from lvc_b200 import ops
from lvc_b200.candidates import CandidateFilter
from lvc_b200.config import DetectorConfig
from lvc_b200.mining import PseudoLabelMiner
from lvc_b200.modeling import DinoViT, GeneralizedRCNN, GeneralizedRCNNRegOnly, synthetic_vit_state_dict
from lvc_b200.structures import Boxes, Instances
from lvc_b200.weights import synthetic_corrector_head, synthetic_state_dict
from oracle import oracle as O
from oracle.vit import vit_forward
from test_mining_host import _chain_with_library_calls

pytestmark = pytest.mark.gpu
MEAN, STD = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)


@pytest.mark.parametrize("with_corrector", [False, True])
def test_miner_equals_stagewise_composition(with_corrector):
    cfg = DetectorConfig(depth=50, score_thresh_test=0.0)
    sd = synthetic_state_dict(cfg, 0)
    g = torch.Generator().manual_seed(11)
    ims = [(torch.rand(3, h, w, generator=g) * 255).to(torch.uint8) for h, w in [(256, 320), (224, 352), (256, 320)]]
    inputs = [{"image": im, "height": int(1.5 * im.shape[1]), "width": int(1.5 * im.shape[2]), "image_id": 10 + i} for i, im in enumerate(ims)]
    det = GeneralizedRCNN(cfg, sd, use_cuda_graph=False)
    base = det(inputs)
    all_scores = torch.cat([r["instances"].scores for r in base])
    all_cls = torch.cat([r["instances"].pred_classes for r in base])
    novel = [int(c) for c in torch.bincount(all_cls, minlength=80).argsort(descending=True)[:6]]
    sc_novel = all_scores[torch.isin(all_cls, torch.tensor(novel))]
    k_min = float(sc_novel.sort().values[int(0.6 * len(sc_novel))])
    filt = CandidateFilter(novel, k_min, 1.0, ar=0.0, full=True, train_imgs={novel[0]: {11}})
    vsd = synthetic_vit_state_dict(depth=2, seed=2)
    vit = DinoViT(vsd)
    bank_desc = vit(torch.randn(96, 3, 224, 224, generator=g).cuda())
    bank_cls = torch.tensor(novel)[torch.randint(0, 3, (96,), generator=g)].cuda()
    bank = ops.KnnBank(bank_desc, bank_cls)
    corrector = None
    if with_corrector:
        ccfg = DetectorConfig(depth=50, num_fc=3)
        full = {k: v for k, v in sd.items() if not k.startswith("roi_heads.")}
        hsd = synthetic_corrector_head(ccfg, 3)
        for k in list(hsd):
            if "bbox_pred.weight" in k:
                hsd[k] = hsd[k] * 30
        full.update(hsd)
        corrector = GeneralizedRCNNRegOnly(ccfg, full)
    miner = PseudoLabelMiner(det, vit, bank, filt, knn=10, corrector=corrector, pixel_mean=MEAN, pixel_std=STD)
    out = miner(inputs)
    assert miner.stats["candidates"] >= 10, miner.stats

    # stage 1: the detector's own result
    for r, b in zip(out, base):
        assert torch.equal(r["instances"].pred_boxes.tensor, b["instances"].pred_boxes.tensor)
        assert torch.equal(r["instances"].scores, b["instances"].scores)
    # stage 2: get_ret_anns on the flat detection list (annotation-id order = image order, then detection order)
    img_id = np.concatenate([[x["image_id"]] * len(r["instances"]) for x, r in zip(inputs, out)])
    cat = torch.cat([r["instances"].pred_classes for r in out]).numpy()
    score = torch.cat([r["instances"].scores for r in out]).numpy()
    bx = torch.cat([r["instances"].pred_boxes.tensor for r in out]).numpy()
    w32, h32 = bx[:, 2] - bx[:, 0], bx[:, 3] - bx[:, 1]
    area = w32.astype(np.float64) * h32.astype(np.float64)
    img_area = np.concatenate([[float(x["height"]) * x["width"]] * len(r["instances"]) for x, r in zip(inputs, out)])
    want_flags = O.select_candidates(img_id, cat, score, area, img_area, {novel[0]: {11}}, novel, k_min, 1.0, ar=0.0, full=True)
    got_flags = torch.cat([r["instances"].candidate_flags for r in out]).numpy()
    assert np.array_equal(got_flags, want_flags)
    # stage 3: DatasetMapperQE boxes, crops, descriptors, votes
    off = 0
    feats_all, cls_all = [], []
    for x, r in zip(inputs, out):
        inst, ci = r["instances"], r["candidates"]
        sel = torch.nonzero(inst.candidate_flags == 1).flatten()
        fw, ww = _chain_with_library_calls(inst.pred_boxes.tensor[sel], (x["height"], x["width"]), tuple(x["image"].shape[-2:]))
        assert torch.equal(ci.det_index, sel) and torch.equal(ci.gt_boxes.tensor, fw)
        assert torch.equal(ci.gt_classes, inst.pred_classes[sel])
        if len(sel):
            crops = O.get_crops_qe(x["image"].numpy(), ww.numpy())
            crops = (torch.from_numpy(crops) - torch.tensor(MEAN).view(1, 3, 1, 1)) / torch.tensor(STD).view(1, 3, 1, 1)
            want = vit_forward(vsd, crops)
            rel = float((ci.crop_feats - want).norm() / want.norm())
            assert rel < 3e-2, rel
        feats_all.append(ci.crop_feats)
        cls_all.append(ci.gt_classes)
    feats_all, cls_all = torch.cat(feats_all), torch.cat(cls_all)
    ref = O.knn_verify(bank_desc.cpu().numpy(), bank_cls.cpu().numpy(), feats_all.numpy(), cls_all.numpy(), topk=10, knn=10)
    got_keep = torch.cat([r["candidates"].keep for r in out]).numpy()
    got_votes = torch.cat([r["candidates"].top10_shots for r in out]).numpy()
    assert np.array_equal(np.sort(got_votes, 1), np.sort(ref["votes"], 1)) and np.array_equal(got_keep, ref["keep"].astype(np.int64))
    assert miner.stats["verified"] == int(got_keep.sum())
    # stage 4: the verified detections (corrected by the cascade heads when a corrector is given)
    for x, r in zip(inputs, out):
        inst, ci, pl = r["instances"], r["candidates"], r["pseudo_labels"]
        kb = ci.keep.bool()
        assert torch.equal(pl.pred_classes, ci.gt_classes[kb]) and torch.equal(pl.scores, ci.scores[kb])
        if not with_corrector:
            assert torch.equal(pl.pred_boxes.tensor, inst.pred_boxes.tensor[ci.det_index[kb]])
    if with_corrector:
        # like the reference's corrector data set, the sub-batch holds only the images that kept a pseudo-label (a ragged batch is padded
        # to its own largest image, so the corrector's features depend on which images share the batch)
        reg_in, with_kept = [], []
        for x, r in zip(inputs, out):
            ci = r["candidates"]
            if not bool(ci.keep.any()):
                assert len(r["pseudo_labels"]) == 0
                continue
            gi = Instances(tuple(x["image"].shape[-2:]))
            gi.gt_boxes = Boxes(ci.gt_boxes.tensor[ci.keep.bool()].clone())
            gi.gt_classes = ci.gt_classes[ci.keep.bool()]
            reg_in.append({"image": x["image"], "height": x["height"], "width": x["width"], "instances": gi})
            with_kept.append(r)
        reg = corrector(reg_in) if reg_in else []
        moved = 0.0
        for q, r in zip(reg, with_kept):
            a, b = q["instances"].pred_boxes.tensor, r["pseudo_labels"].pred_boxes.tensor
            assert a.shape == b.shape and float((a - b).abs().max()) < 1e-3
            ci = r["candidates"]
            det_b = r["instances"].pred_boxes.tensor[ci.det_index[ci.keep.bool()]]
            moved = max(moved, float((b - det_b).abs().max()))
        assert miner.stats["verified"] == 0 or moved > 0.5       # the corrector did regress the boxes


def test_miner_no_candidates():
    cfg = DetectorConfig(depth=50)
    det = GeneralizedRCNN(cfg, synthetic_state_dict(cfg, 0), use_cuda_graph=False)
    vit = DinoViT(synthetic_vit_state_dict(depth=1, seed=0))
    bank = ops.KnnBank(torch.randn(64, 384).cuda(), torch.randint(0, 5, (64,)).cuda())
    miner = PseudoLabelMiner(det, vit, bank, CandidateFilter([1, 2], 0.999, 1.0))
    out = miner([{"image": torch.zeros(3, 128, 160, dtype=torch.uint8)}])
    assert miner.stats["candidates"] == 0 and len(out[0]["pseudo_labels"]) == 0 and len(out[0]["candidates"]) == 0


def test_miner_stream_equals_blocking_calls():
    """miner.stream(batches) (three batches in flight, copy stream, deferred assembly) returns what miner(batch) returns, in order."""
    cfg = DetectorConfig(depth=50, score_thresh_test=0.0)
    sd = synthetic_state_dict(cfg, 0)
    det = GeneralizedRCNN(cfg, sd, use_cuda_graph=True)
    vit = DinoViT(synthetic_vit_state_dict(depth=2, seed=2))
    g = torch.Generator().manual_seed(5)
    bank = ops.KnnBank(torch.randn(80, 384, generator=g).cuda(), torch.randint(0, 4, (80,), generator=g).cuda())
    ccfg = DetectorConfig(depth=50, num_fc=3)
    full = {k: v for k, v in sd.items() if not k.startswith("roi_heads.")}
    full.update(synthetic_corrector_head(ccfg, 3))
    miner = PseudoLabelMiner(det, vit, bank, CandidateFilter(range(0, 80, 3), 0.03, 1.0), corrector=GeneralizedRCNNRegOnly(ccfg, full))
    shapes = [(192, 256)] * 3 + [(160, 224)] * 2 + [(192, 256)] * 4
    batches = [[{"image": (torch.rand(3, h, w, generator=g) * 255).to(torch.uint8), "height": h, "width": w, "image_id": 2 * b + j} for j in range(2)]
               for b, (h, w) in enumerate(shapes)]
    want = [miner(b) for b in batches]
    n_cand = 0
    got = list(miner.stream(iter(batches)))
    assert len(got) == len(want)
    for rb, wb in zip(got, want):
        for r, w in zip(rb, wb):
            assert torch.equal(r["instances"].pred_boxes.tensor, w["instances"].pred_boxes.tensor)
            assert torch.equal(r["candidates"].keep, w["candidates"].keep) and torch.equal(r["candidates"].top10_shots, w["candidates"].top10_shots)
            assert torch.equal(r["pseudo_labels"].pred_boxes.tensor, w["pseudo_labels"].pred_boxes.tensor)
            n_cand += len(r["candidates"])
    assert n_cand > 0
    assert list(miner.stream([])) == []
    # the same run through the reference-style driver: inference_on_dataset(miner, loader, PseudoLabelCollector())
    from lvc_b200.evaluation import PseudoLabelCollector, inference_on_dataset
    res = inference_on_dataset(miner, iter(batches), PseudoLabelCollector())
    n_pl = sum(len(w["pseudo_labels"].pred_boxes) for wb in want for w in wb)
    assert res["num_images"] == 2 * len(batches) and res["num_pseudo_labels"] == n_pl and res["num_candidates"] == n_cand
    flat = [w for wb in want for w in wb]
    k = 0
    for w, x in zip(flat, [x for b in batches for x in b]):
        for j in range(len(w["pseudo_labels"].pred_boxes)):
            a = res["annotations"][k]
            bx = w["pseudo_labels"].pred_boxes.tensor[j]
            assert a["image_id"] == x["image_id"] and a["category_id"] == int(w["pseudo_labels"].pred_classes[j])
            assert abs(a["bbox"][0] - float(bx[0])) < 1e-4 and abs(a["bbox"][2] - float(bx[2] - bx[0])) < 1e-3
            k += 1
