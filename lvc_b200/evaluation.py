"""Callers and wire format either side of the detection path, mirrored from the reference:

* ``instances_to_coco_json``  -- lvc/evaluation/coco_evaluation.py:566-603 (XYXY -> XYWH, python lists, one dict per detection)
* ``DatasetEvaluator`` / ``DatasetEvaluators`` / ``COCOResultCollector`` -- the reset / process / evaluate protocol of
  lvc/evaluation/evaluator.py:13-82 and the prediction-gathering half of COCOEvaluator (coco_evaluation.py:96-147, 302-312)
* ``inference_on_dataset``    -- lvc/evaluation/evaluator.py:85-161, same contract (returns ``evaluator.evaluate()``), but
  the loop is software-pipelined through ``GeneralizedRCNN.inference_stream`` (H2D of batch i+1 overlaps the forward of batch i)
  and there is no per-image ``cuda.synchronize()``; any batch size the loader yields is accepted (the reference pins 1).
"""
import collections
import datetime
import itertools
import json
import logging
import os
import time
from collections import OrderedDict

import torch
import torch.distributed as dist


def instances_to_coco_json(instances, img_id):
    """Dump an ``Instances`` to the reference's COCO-result dicts (coco_evaluation.py:566-603)."""
    n = len(instances)
    if n == 0:
        return []
    boxes = instances.pred_boxes.tensor.clone()
    boxes[:, 2] -= boxes[:, 0]          # BoxMode.convert(XYXY_ABS -> XYWH_ABS), detectron2/structures/boxes.py:43-129
    boxes[:, 3] -= boxes[:, 1]
    boxes = boxes.tolist()
    scores = instances.scores.tolist()
    classes = instances.pred_classes.tolist()
    return [{"image_id": img_id, "category_id": classes[k], "bbox": boxes[k], "score": scores[k]} for k in range(n)]


class DatasetEvaluator:
    def reset(self):
        pass

    def process(self, inputs, outputs):
        pass

    def evaluate(self):
        pass


class DatasetEvaluators(DatasetEvaluator):
    def __init__(self, evaluators):
        self._evaluators = evaluators

    def reset(self):
        for e in self._evaluators:
            e.reset()

    def process(self, inputs, outputs):
        for e in self._evaluators:
            e.process(inputs, outputs)

    def evaluate(self):
        results = OrderedDict()
        for e in self._evaluators:
            r = e.evaluate()
            if r is not None and (not dist.is_initialized() or dist.get_rank() == 0):
                for k, v in r.items():
                    assert k not in results, f"Different evaluators produce results with the same key {k}"
                    results[k] = v
        return results


class COCOResultCollector(DatasetEvaluator):
    """The candidate-sourcing half of COCOEvaluator: collect per-image predictions, gather them on rank 0 in rank order
    (== itertools.chain(*comm.gather(...)), coco_evaluation.py:119-126), write ``coco_instances_results.json`` with the
    contiguous -> dataset category-id mapping applied (coco_evaluation.py:288-312)."""

    def __init__(self, output_dir=None, contiguous_id_to_dataset_id=None, file_name="coco_instances_results.json"):
        self._output_dir, self._map, self._file_name = output_dir, contiguous_id_to_dataset_id, file_name
        self._predictions = []

    def reset(self):
        self._predictions = []

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            pred = {"image_id": inp["image_id"]}
            if "instances" in out:
                pred["instances"] = instances_to_coco_json(out["instances"].to("cpu"), inp["image_id"])
            self._predictions.append(pred)

    def evaluate(self):
        preds = self._predictions
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
            dist.gather_object(preds, gathered, dst=0)
            if dist.get_rank() != 0:
                return {}
            preds = list(itertools.chain(*gathered))
        results = list(itertools.chain(*[p["instances"] for p in preds]))
        if self._map is not None:
            for r in results:
                r["category_id"] = self._map[r["category_id"]]
        if self._output_dir:
            os.makedirs(self._output_dir, exist_ok=True)
            with open(os.path.join(self._output_dir, self._file_name), "w") as f:
                f.write(json.dumps(results))
        return {"num_images": len(preds), "num_detections": len(results), "results": results}


class CandidateCollector(DatasetEvaluator):
    """Pseudo-label candidates straight from the detector (rows a15 / f2): consumes results that carry ``candidate_flags``
    (``model.candidate_filter``, the device-side form of get_ret_anns' score mode) and emits the annotations the reference's
    offline pass would write (tools/create_coco_dataset_from_dets_all.py:129-238): COCO-result dicts with ``ignore_qe`` /
    ``iscrowd`` set, ``area`` = w * h and ``id`` = pycocotools ``loadRes``' running index over ALL detections in dataset order
    (the id the next stage reuses as ``instances.ids``, dataset_mapper.py:383-401).  Ranks gather in rank order
    (== ``chain(*comm.gather(...))``, coco_evaluation.py:119-126): with ``inference_shard`` that is dataset order."""

    def __init__(self, contiguous_id_to_dataset_id=None):
        self._map = contiguous_id_to_dataset_id
        self._records = []

    def reset(self):
        self._records = []

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            inst = out["instances"]
            n = len(inst)
            kept = []
            if n and inst.has("candidate_flags"):
                flags = inst.candidate_flags
                sel = flags.nonzero().flatten().tolist()
                if sel:
                    dets = instances_to_coco_json(inst, inp["image_id"])
                    for j in sel:
                        d = dets[j]
                        ig = int(flags[j] == 2)
                        d.update(ignore_qe=ig, iscrowd=ig, area=d["bbox"][2] * d["bbox"][3], local_index=j)
                        kept.append(d)
            self._records.append((inp["image_id"], n, kept))

    def evaluate(self):
        recs = self._records
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
            dist.gather_object(recs, gathered, dst=0)
            if dist.get_rank() != 0:
                return {}
            recs = list(itertools.chain(*gathered))
        anns, offset = [], 0
        for image_id, n, kept in recs:
            for d in kept:
                d["id"] = offset + d.pop("local_index") + 1
                if self._map is not None:
                    d["category_id"] = self._map[d["category_id"]]
                anns.append(d)
            offset += n
        return {"num_images": len(recs), "num_detections": offset, "annotations": anns,
                "num_candidates": sum(1 for a in anns if a["ignore_qe"] == 0)}


class PseudoLabelCollector(DatasetEvaluator):
    """End of the chained pipeline (``lvc_b200.mining.PseudoLabelMiner``): consumes results that carry ``pseudo_labels`` (the verified, and
    if a corrector ran corrected, boxes in the output frame) and emits them as the annotation list the reference's last stage writes
    (tools/train_net_reg_qe.py -> ``UBBRSaver``: COCO-style dicts with ``image_id``, ``category_id``, XYWH ``bbox``, ``score``, ``area``,
    ``iscrowd`` 0 and a running ``id``).  Ranks gather in rank order (== ``chain(*comm.gather(...))``): with ``inference_shard`` that is
    dataset order.  Also counts what went through the stages (detections, candidates, verified)."""

    def __init__(self, contiguous_id_to_dataset_id=None):
        self._map = contiguous_id_to_dataset_id
        self._records = []

    def reset(self):
        self._records = []

    def process(self, inputs, outputs):
        for inp, out in zip(inputs, outputs):
            pl = out["pseudo_labels"]
            rows = instances_to_coco_json(pl, inp["image_id"]) if len(pl.pred_boxes) else []
            n_det = len(out["instances"]) if "instances" in out else 0
            n_cand = len(out["candidates"].gt_boxes) if "candidates" in out else 0
            self._records.append((inp["image_id"], n_det, n_cand, rows))

    def evaluate(self):
        recs = self._records
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
            dist.gather_object(recs, gathered, dst=0)
            if dist.get_rank() != 0:
                return {}
            recs = list(itertools.chain(*gathered))
        anns = []
        for image_id, n_det, n_cand, rows in recs:
            for d in rows:
                d.update(id=len(anns) + 1, area=d["bbox"][2] * d["bbox"][3], iscrowd=0)
                if self._map is not None:
                    d["category_id"] = self._map[d["category_id"]]
                anns.append(d)
        return {"num_images": len(recs), "num_detections": sum(r[1] for r in recs), "num_candidates": sum(r[2] for r in recs),
                "num_pseudo_labels": len(anns), "annotations": anns}


def inference_shard(n, rank=None, world_size=None):
    """The reference's InferenceSampler rule (detectron2/data/samplers/distributed_sampler.py:191-194): rank r of W takes the
    contiguous block [r * ceil(n / W), min((r + 1) * ceil(n / W), n)) -- so that chain(*gather(...)) on rank 0 is in dataset order.
    Images shard with NO data-path collective; only the results are gathered."""
    if rank is None or world_size is None:
        on = dist.is_available() and dist.is_initialized()
        rank, world_size = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    per = (n - 1) // world_size + 1 if n > 0 else 0
    begin = per * rank
    return range(min(begin, n), min(begin + per, n))


def inference_on_dataset(model, data_loader, evaluator):
    """Same contract as the reference's ``inference_on_dataset``: runs ``model`` over ``data_loader`` (an iterable with a
    length or a plain iterator, yielding list[dict] batches), feeds every (inputs, outputs) pair to ``evaluator.process`` and returns
    ``evaluator.evaluate()``; logs the reference's two timing lines."""
    logger = logging.getLogger(__name__)
    num_devices = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    total = len(data_loader) if hasattr(data_loader, "__len__") else "?"
    logger.info("Start inference on {} batches".format(total))
    evaluator.reset()
    start = time.time()
    n_img = 0
    # The loader is consumed LAZILY (the mining path runs over the whole training set: ~100 k images of ~3 MB each): the
    # pipelined stream pulls one batch ahead, and only the inputs of in-flight batches are remembered here.
    in_flight = collections.deque()

    def feed():
        for inputs in data_loader:
            in_flight.append(inputs)
            yield inputs

    stream = model.inference_stream(feed()) if hasattr(model, "inference_stream") else (model(b) for b in feed())
    with torch.no_grad():
        for outputs in stream:
            inputs = in_flight.popleft()
            evaluator.process(inputs, outputs)
            n_img += len(inputs)
    total_time = time.time() - start
    logger.info("Total inference time: {} ({:.6f} s / img per device, on {} devices)".format(
        str(datetime.timedelta(seconds=int(total_time))), total_time / max(n_img, 1), num_devices))
    results = evaluator.evaluate()
    return {} if results is None else results
