"""Label -> Verify (-> Correct) in one pass over a batch, without the JSON files the reference puts between its stages.

The reference mines pseudo-labels with three tools run one after the other, each re-reading the images:

    tools/train_net.py --eval-only                     detections of the base detector -> coco_instances_results.json
    tools/create_coco_dataset_from_dets_all.py          score / class / area filter (get_ret_anns :129-238) -> candidate dataset
    tools/run_nearest_neighbours.py --eval-only         DatasetMapperQE crops (dataset_mapper.py:407-409) -> DINO ViT -> kNN vote -> keep
    tools/train_net_reg_qe.py --eval-only               GeneralizedRCNNRegOnly on the kept boxes -> corrected boxes

``PseudoLabelMiner`` chains the same operators on the device for one batch: the detector's output block stays in HBM, the
candidate filter runs on it, the crops are cut from the device-resident detector input (the image ``DatasetMapperQE`` would
re-load and resize the same way), the descriptors go straight into the kNN verifier, and the corrector reads the same images.
One packed D2H (detections + flags) is the only host round trip before the final result; it is needed because the number of
candidates sizes the ViT batch.

The crop windows are integer pixel boxes.  The reference obtains them from the detections through a chain of conversions
(``instances_to_coco_json`` coco_evaluation.py:566-603 -> JSON -> ``BoxMode.convert`` boxes.py:50-123 -> ``ResizeTransform`` ->
clip, detection_utils.py:279-283 -> ``Boxes`` fp32 -> ``.long()`` lvc/data/utils.py:491); ``reference_crop_boxes`` walks the
same chain so that the windows are the reference's, not merely close to them.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .candidates import CandidateFilter
from .crops import crop_geometry
from .structures import Boxes, Instances


def reference_crop_boxes(boxes_out: np.ndarray, out_hw, in_hw):
    """Detections ``boxes_out`` [m, 4] fp32 XYXY in the OUTPUT frame (height, width = ``out_hw``, what ``detector_postprocess``
    returns) -> (fp32 boxes in the detector-input frame ``in_hw`` as ``DatasetMapperQE`` builds them, their ``.long()`` crop
    windows, non-empty mask of ``filter_empty_instances``).  Every rounding step of the reference's chain is kept:

    * XYXY -> XYWH on the fp32 array (coco_evaluation.py:583-585), ``tolist()`` -> JSON doubles that hold fp32 values;
    * XYWH -> XYXY on ``torch.tensor(list)`` = fp32 again (boxes.py:63, 111-113), ``tolist()``;
    * float64 scaling by ``new_w / w``, ``new_h / h`` (fvcore ResizeTransform.apply_coords via apply_box), clip to the image
      (detection_utils.py:281-282);
    * ``Boxes`` casts to fp32 (annotations_to_instances), ``.long()`` truncates (lvc/data/utils.py:491)."""
    b = np.asarray(boxes_out, np.float32).reshape(-1, 4)
    oh, ow = float(out_hw[0]), float(out_hw[1])
    ih, iw = int(in_hw[0]), int(in_hw[1])
    xywh = b.copy()
    xywh[:, 2] = xywh[:, 2] - xywh[:, 0]
    xywh[:, 3] = xywh[:, 3] - xywh[:, 1]
    xyxy = xywh.copy()                                 # fp32: torch.tensor(python floats) is float32
    xyxy[:, 2] = xyxy[:, 2] + xyxy[:, 0]
    xyxy[:, 3] = xyxy[:, 3] + xyxy[:, 1]
    d = xyxy.astype(np.float64)
    d[:, 0::2] *= iw * 1.0 / ow
    d[:, 1::2] *= ih * 1.0 / oh
    d = np.minimum(d.clip(min=0), np.array([iw, ih, iw, ih], np.float64))
    f = d.astype(np.float32)
    nonempty = ((f[:, 2] - f[:, 0]) > 1e-5) & ((f[:, 3] - f[:, 1]) > 1e-5)      # filter_empty_instances, detection_utils.py:421-448
    return f, f.astype(np.int64), nonempty


class PseudoLabelMiner:
    """``miner(batched_inputs)``: ``batched_inputs`` as for ``GeneralizedRCNN`` (``image`` uint8 / float [3,H,W] in the detector's
    channel order, ``height`` / ``width`` of the output frame, optional ``image_id``).  Returns one dict per image:

    * ``"instances"``: the detector's ``Instances`` (pred_boxes, scores, pred_classes, candidate_flags) -- what stage 1 writes;
    * ``"candidates"``: ``Instances`` of the flag-1 detections in the detector-input frame: ``gt_boxes`` (fp32, the boxes
      ``DatasetMapperQE`` would build), ``gt_classes``, ``scores``, ``crop_feats``, ``top10_shots``, ``keep`` -- what stage 3 holds;
    * ``"pseudo_labels"``: ``Instances`` in the output frame of the verified candidates: ``pred_boxes`` (regressed by the corrector
      when one is given, else the detector's), ``pred_classes``, ``scores``.

    ``descriptor``: crops [n,3,224,224] fp32 -> [n, D] (lvc_b200.modeling.DinoViT); ``bank``: ops.KnnBank over the support set;
    ``pixel_mean`` / ``pixel_std``: ``cfg.MODEL.PIXEL_MEAN / PIXEL_STD`` of the verification config (preprocess_crops,
    run_nearest_neighbours.py:102-105), in the channel order of ``image`` unless ``swap_channels`` (then the crops are taken from
    the channel-reversed image: a BGR detector feeding an RGB descriptor)."""

    def __init__(self, detector, descriptor, bank: "ops.KnnBank", candidate_filter: CandidateFilter, knn: int = 10,
                 corrector=None, pixel_mean: Sequence[float] = (123.675, 116.28, 103.53), pixel_std: Sequence[float] = (58.395, 57.12, 57.375),
                 swap_channels: bool = False, operation: str = "context", crop_size: int = 224, max_descriptor_batch: int = 512):
        _lib.load()
        self.detector, self.descriptor, self.bank, self.filter = detector, descriptor, bank, candidate_filter
        self.corrector, self.knn = corrector, int(knn)
        dev = detector._device
        self.mean = torch.as_tensor(pixel_mean, dtype=torch.float32, device=dev)
        self.inv_std = 1.0 / torch.as_tensor(pixel_std, dtype=torch.float32, device=dev)
        self.swap, self.operation, self.size, self.max_batch = bool(swap_channels), operation, int(crop_size), int(max_descriptor_batch)
        self.stats = {}

    # --------------------------------------------------------------------------------------------------------------- stages
    def _crops(self, image: torch.Tensor, windows: np.ndarray) -> torch.Tensor:
        img = image.flip(0) if self.swap else image
        img = img.contiguous() if img.dtype == torch.uint8 else img.float().contiguous()
        _, H, W = img.shape
        geom = torch.from_numpy(crop_geometry(windows, H, W, self.operation)).to(img.device, non_blocking=True)
        out = torch.empty((len(windows), 3, self.size, self.size), dtype=torch.float32, device=img.device)
        rc = _lib.load().lvcb200_crops_qe(_lib.ptr(img), _lib.U8 if img.dtype == torch.uint8 else _lib.F32, H, W, _lib.ptr(geom),
                                          len(windows), self.size, _lib.ptr(self.mean), _lib.ptr(self.inv_std), _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_crops_qe")
        return out

    def _describe(self, crops: torch.Tensor) -> torch.Tensor:
        if crops.shape[0] <= self.max_batch:
            return self.descriptor(crops)
        return torch.cat([self.descriptor(crops[i:i + self.max_batch]) for i in range(0, crops.shape[0], self.max_batch)])

    @torch.no_grad()
    def __call__(self, batched_inputs: List[dict]) -> List[dict]:
        det = self.detector
        dev = det._device
        images = det.to_device(batched_inputs)
        sizes = [tuple(im.shape[-2:]) for im in images]
        outs = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)]
        boxes, scores, classes, rows, counts = det.engine.run(images, outs)
        ids = [x.get("image_id") for x in batched_inputs]
        flags, _ = self.filter(boxes, scores, classes, counts, outs, ids if all(i is not None for i in ids) else None)
        n, k = scores.shape
        # the one host round trip in front of the verifier: detections + flags, packed
        host = torch.cat([boxes.reshape(n, 4 * k), scores, classes.float(), flags.float(), counts.float()[:, None]], dim=1).cpu()

        results, cand = [], []          # cand: per image (indices into the detections, fp32 boxes in the input frame, windows)
        for i, o in enumerate(outs):
            c = int(host[i, -1])
            inst = Instances(o)
            inst.pred_boxes = Boxes(host[i, :4 * k].view(k, 4)[:c].clone())
            inst.scores = host[i, 4 * k:5 * k][:c].clone()
            inst.pred_classes = host[i, 5 * k:6 * k][:c].to(torch.int64)
            inst.candidate_flags = host[i, 6 * k:7 * k][:c].to(torch.int8)
            sel = torch.nonzero(inst.candidate_flags == 1).flatten()
            fb, win, ok = reference_crop_boxes(inst.pred_boxes.tensor[sel].numpy(), o, sizes[i])
            sel, fb, win = sel[torch.from_numpy(ok)], fb[ok], win[ok]
            cand.append((sel, fb, win))
            results.append({"instances": inst})

        m = [len(c[0]) for c in cand]
        total = sum(m)
        self.stats = {"detections": int(host[:, -1].sum()), "candidates": total}
        if total:
            crops = torch.cat([self._crops(images[i], cand[i][2]) for i in range(n) if m[i]])
            feats = self._describe(crops)
            qcls = torch.cat([results[i]["instances"].pred_classes[cand[i][0]] for i in range(n)]).to(dev)
            res = self.bank.verify(feats, qcls, topk=10, knn=self.knn)
            keep, votes = res["keep"].to(torch.int64).cpu(), res["votes"].cpu()
            feats_h = feats.cpu()
        off = 0
        kept_boxes = []
        for i, (sel, fb, win) in enumerate(cand):
            inst = results[i]["instances"]
            ci = Instances(sizes[i])
            ci.gt_boxes = Boxes(torch.from_numpy(fb).reshape(-1, 4))
            ci.gt_classes = inst.pred_classes[sel]
            ci.scores = inst.scores[sel]
            ci.det_index = sel
            if total:
                ci.crop_feats = feats_h[off:off + m[i]]
                ci.top10_shots = votes[off:off + m[i]]
                ci.keep = keep[off:off + m[i]]
            else:
                ci.crop_feats = torch.zeros((0, getattr(self.bank, "D", 0)))
                ci.top10_shots = torch.zeros((0, 10), dtype=torch.int64)
                ci.keep = torch.zeros(0, dtype=torch.int64)
            off += m[i]
            results[i]["candidates"] = ci
            kept_boxes.append(ci.gt_boxes.tensor[ci.keep.bool()])
        self.stats["verified"] = int(sum(len(b) for b in kept_boxes))

        corrected: List[Optional[torch.Tensor]] = [None] * n
        if self.corrector is not None and self.stats["verified"]:
            # train_net_reg_qe.py --eval-only on the verified boxes (GeneralizedRCNNRegOnly.inference, rcnn.py:372-410), same device images
            pyramid, _ = self.corrector.engine.run_features(images)
            planes = [pyramid[l] for l in (2, 3, 4, 5)]
            reg = self.corrector.head(planes, [b.to(dev) for b in kept_boxes], sizes)
            corrected = [r.cpu() for r in reg]
        for i, o in enumerate(outs):
            inst, ci = results[i]["instances"], results[i]["candidates"]
            kb = ci.keep.bool()
            pl = Instances(o)
            if corrected[i] is not None:
                sx, sy = o[1] / sizes[i][1], o[0] / sizes[i][0]                 # detector_postprocess, postprocessing.py:37-59
                b = corrected[i].clone()
                b[:, 0::2] *= sx
                b[:, 1::2] *= sy
                bx = Boxes(b)
                bx.clip(o)
                pl.pred_boxes = bx
            else:
                pl.pred_boxes = Boxes(inst.pred_boxes.tensor[ci.det_index[kb]].clone())
            pl.pred_classes = ci.gt_classes[kb]
            pl.scores = ci.scores[kb]
            results[i]["pseudo_labels"] = pl
        return results
