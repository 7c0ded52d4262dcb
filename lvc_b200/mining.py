"""Label -> Verify (-> Correct) in one pass over a batch, without the JSON files the reference puts between its stages.

The reference mines pseudo-labels with three tools run one after the other, each re-reading the images:

    tools/train_net.py --eval-only                     detections of the base detector -> coco_instances_results.json
    tools/create_coco_dataset_from_dets_all.py          score / class / area filter (get_ret_anns :129-238) -> candidate dataset
    tools/run_nearest_neighbours.py --eval-only         DatasetMapperQE crops (dataset_mapper.py:407-409) -> DINO ViT -> kNN vote -> keep
    tools/train_net_reg_qe.py --eval-only               GeneralizedRCNNRegOnly on the kept boxes -> corrected boxes

``PseudoLabelMiner`` chains the same operators on the device for one batch: the detector's output block stays in HBM, the
candidate filter runs on it, the crops are cut from the device-resident detector input (the image ``DatasetMapperQE`` would
re-load and resize the same way), the descriptors go straight into the kNN verifier, and the corrector reads the same images
(only those that kept a pseudo-label, like the reference's corrector data set).
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
        self._host_ring, self._img_ring = {}, {}

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

    # ------------------------------------------------------------------------------------------------------------- pipeline
    # label(): everything up to the packed D2H of the detections is enqueued without waiting; verify(): needs the host copy (the number
    # of candidates sizes the ViT batch), enqueues crops -> ViT -> kNN (-> corrector) and their D2H; finish(): assembles the result.
    def _label(self, batched_inputs: List[dict], images=None):
        det = self.detector
        dev = det._device
        if images is None:
            images = det.to_device(batched_inputs)
        sizes = [tuple(im.shape[-2:]) for im in images]
        outs = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)]
        boxes, scores, classes, rows, counts = det.engine.run(images, outs)
        ids = [x.get("image_id") for x in batched_inputs]
        flags, _ = self.filter(boxes, scores, classes, counts, outs, ids if all(i is not None for i in ids) else None)
        n, k = scores.shape
        # the one host round trip in front of the verifier: detections + flags, packed
        packed = torch.cat([boxes.reshape(n, 4 * k), scores, classes.float(), flags.float(), counts.float()[:, None]], dim=1)
        key = tuple(packed.shape)
        ring = self._host_ring.setdefault(key, [[torch.empty(packed.shape, dtype=packed.dtype, pin_memory=True) for _ in range(4)], 0])
        host = ring[0][ring[1] % 4]
        ring[1] += 1
        host.copy_(packed, non_blocking=True)
        return dict(images=images, sizes=sizes, outs=outs, host=host, k=k, ready=torch.cuda.current_stream(dev).record_event())

    def _verify(self, st):
        dev = self.detector._device
        st["ready"].synchronize()
        host, k, images, sizes, outs = st["host"], st["k"], st["images"], st["sizes"], st["outs"]
        n = len(outs)
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
        st.update(results=results, cand=cand, m=m, total=total, detections=int(host[:, -1].sum()), out=None)
        if total:
            crops = torch.cat([self._crops(images[i], cand[i][2]) for i in range(n) if m[i]])
            feats = self._describe(crops)
            qcls = torch.cat([results[i]["instances"].pred_classes[cand[i][0]] for i in range(n)]).to(dev, non_blocking=True)
            res = self.bank.verify(feats, qcls, topk=10, knn=self.knn)
            D = feats.shape[1]
            out = torch.cat([feats, res["votes"].float(), res["keep"].float()[:, None]], dim=1)       # [total, D + 10 + 1], one D2H
            hout = torch.empty(out.shape, dtype=torch.float32, pin_memory=True)
            hout.copy_(out, non_blocking=True)
            st.update(out=hout, D=D)
        st["done"] = torch.cuda.current_stream(dev).record_event()
        return st

    def _correct(self, st):
        """train_net_reg_qe.py --eval-only (GeneralizedRCNNRegOnly.inference, rcnn.py:372-410) on the verified boxes.  Like the reference,
        whose corrector data set holds only the images that kept a pseudo-label, the corrector's backbone runs on those images only: the
        sub-batch is padded (by repeating its first image) to the next power of two so that the engine keeps at most log2(batch) + 1
        activation sets per image shape."""
        st["done"].synchronize()
        dev = self.detector._device
        cand, m, total, images, sizes = st["cand"], st["m"], st["total"], st["images"], st["sizes"]
        st["corr"] = None
        if self.corrector is None or not total:
            return st
        D = st["D"]
        keep = st["out"][:, D + 10].bool()
        off, idx, kept = 0, [], []
        for i, (sel, fb, win) in enumerate(cand):
            kb = keep[off:off + m[i]].numpy()
            off += m[i]
            if kb.any():
                idx.append(i)
                kept.append(torch.from_numpy(fb[kb]).reshape(-1, 4))
        if not idx:
            return st
        nb = 1
        while nb < len(idx):
            nb *= 2
        nb = min(nb, max(len(images), len(idx)))
        pad = nb - len(idx)
        sub_images = [images[i] for i in idx] + [images[idx[0]]] * pad
        sub_sizes = [sizes[i] for i in idx] + [sizes[idx[0]]] * pad
        pyramid, _ = self.corrector.engine.run_features(sub_images)
        planes = [pyramid[l] for l in (2, 3, 4, 5)]
        boxes = [b.to(dev, non_blocking=True) for b in kept] + [torch.zeros((0, 4), device=dev)] * pad
        reg = torch.cat(self.corrector.head(planes, boxes, sub_sizes))
        hreg = torch.empty(reg.shape, dtype=torch.float32, pin_memory=True)
        hreg.copy_(reg, non_blocking=True)
        st["corr"] = (idx, [len(b) for b in kept], hreg)
        st["done"] = torch.cuda.current_stream(dev).record_event()
        return st

    def _finish(self, st) -> List[dict]:
        st["done"].synchronize()
        results, cand, m, total, sizes, outs = st["results"], st["cand"], st["m"], st["total"], st["sizes"], st["outs"]
        hout = st["out"]
        D = st.get("D", getattr(self.bank, "D", 0))
        off = 0
        verified = 0
        corrected = {}
        if st.get("corr") is not None:
            idx, lens, hreg = st["corr"]
            o2 = 0
            for i, n_i in zip(idx, lens):
                corrected[i] = hreg[o2:o2 + n_i]
                o2 += n_i
        for i, (sel, fb, win) in enumerate(cand):
            inst, o = results[i]["instances"], outs[i]
            ci = Instances(sizes[i])
            ci.gt_boxes = Boxes(torch.from_numpy(fb).reshape(-1, 4))
            ci.gt_classes = inst.pred_classes[sel]
            ci.scores = inst.scores[sel]
            ci.det_index = sel
            rows = hout[off:off + m[i]] if total else torch.zeros((0, D + 11))
            ci.crop_feats = rows[:, :D].clone()
            ci.top10_shots = rows[:, D:D + 10].to(torch.int64)
            ci.keep = rows[:, D + 10].to(torch.int64)
            off += m[i]
            results[i]["candidates"] = ci
            kb = ci.keep.bool()
            verified += int(kb.sum())
            pl = Instances(o)
            if i in corrected:
                b = corrected[i].clone()
                b[:, 0::2] *= o[1] / sizes[i][1]                              # detector_postprocess, postprocessing.py:37-59
                b[:, 1::2] *= o[0] / sizes[i][0]
                bx = Boxes(b)
                bx.clip(o)
                pl.pred_boxes = bx
            else:
                pl.pred_boxes = Boxes(inst.pred_boxes.tensor[sel[kb]].clone())
            pl.pred_classes = ci.gt_classes[kb]
            pl.scores = ci.scores[kb]
            results[i]["pseudo_labels"] = pl
        self.stats = {"detections": st["detections"], "candidates": total, "verified": verified}
        return results

    @torch.no_grad()
    def __call__(self, batched_inputs: List[dict]) -> List[dict]:
        return self._finish(self._correct(self._verify(self._label(batched_inputs))))

    def inference_stream(self, batches):
        """The name ``lvc_b200.evaluation.inference_on_dataset`` looks for: ``inference_on_dataset(miner, loader, PseudoLabelCollector())``."""
        return self.stream(batches)

    @torch.no_grad()
    def stream(self, batches):
        """``miner(inputs)`` for every batch of the iterable, in order, software-pipelined: while the GPU runs the detector on batch i
        the host selects the candidates of batch i - 1 and enqueues their crops / ViT / kNN behind it, and assembles the result of
        batch i - 2 -- no stage waits for the device with nothing queued behind it.  The H2D copy of the next batch runs on a copy
        stream, into persistent device image sets (four batches are in flight -- label, verify, correct, assemble: six sets per shape)."""
        from collections import deque
        dev = self.detector._device
        main, copy_stream = torch.cuda.current_stream(dev), torch.cuda.Stream(device=dev)

        def stage(batched_inputs):
            key = tuple((tuple(x["image"].shape), x["image"].dtype) for x in batched_inputs)
            ring = self._img_ring.get(key)
            if ring is None:
                while len(self._img_ring) >= 3:
                    self._img_ring.pop(next(iter(self._img_ring)))
                ring = self._img_ring[key] = {"sets": [[torch.empty(sh, dtype=dt, device=dev) for sh, dt in key] for _ in range(6)],
                                              "free": [None] * 6, "n": 0}
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_stream(main)               # new blocks may recycle memory a queued kernel still reads
            slot = ring["n"] % 6
            ring["n"] += 1
            images = ring["sets"][slot]
            with torch.cuda.stream(copy_stream):
                if ring["free"][slot] is not None:
                    copy_stream.wait_event(ring["free"][slot])  # the set's previous user (six batches ago) is through the corrector
                for dst, x in zip(images, batched_inputs):
                    dst.copy_(x["image"], non_blocking=True)
                ready = copy_stream.record_event()
            return batched_inputs, images, ready, (ring, slot)

        it = iter(batches)
        nxt = next(it, None)
        staged = stage(nxt) if nxt is not None else None
        labelled, verified, corrected = deque(), deque(), deque()
        while staged is not None or labelled or verified or corrected:
            if staged is not None:
                batched_inputs, images, ready, ring = staged
                nxt = next(it, None)
                staged_next = stage(nxt) if nxt is not None else None
                main.wait_event(ready)
                labelled.append((self._label(batched_inputs, images), ring))
                staged = staged_next
            if labelled and (len(labelled) > 1 or staged is None):
                st, ring = labelled.popleft()
                verified.append((self._verify(st), ring))
            if verified and (len(verified) > 1 or (staged is None and not labelled)):
                st, ring = verified.popleft()
                corrected.append(self._correct(st))
                ring[0]["free"][ring[1]] = st["done"]          # the image set is free once the corrector (its last reader) has run
            if corrected and (len(corrected) > 1 or (staged is None and not labelled and not verified)):
                yield self._finish(corrected.popleft())
