"""Wire format and driver mirrors, CPU: against the reference's own instances_to_coco_json when the
reference tree is present, and protocol checks with a stub model otherwise."""
import json
import os

import pytest
import torch

from lvc_b200.evaluation import COCOResultCollector, DatasetEvaluators, inference_on_dataset, instances_to_coco_json
from lvc_b200.structures import Boxes, Instances


def _inst(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    b = torch.rand(n, 4, generator=g) * 100
    b[:, 2:] += b[:, :2]
    i = Instances((200, 300))
    i.pred_boxes = Boxes(b)
    i.scores = torch.rand(n, generator=g)
    i.pred_classes = torch.randint(0, 80, (n,), generator=g)
    return i


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_instances_to_coco_json_matches_reference():
    from oracle import ref_shim
    ref_shim.install()
    from detectron2.structures import Boxes as RB, Instances as RI
    from lvc.evaluation.coco_evaluation import instances_to_coco_json as ref_fn
    for n in (0, 1, 17):
        mine = _inst(n)
        ri = RI((200, 300))
        ri.pred_boxes = RB(mine.pred_boxes.tensor.clone())
        ri.scores = mine.scores.clone()
        ri.pred_classes = mine.pred_classes.clone()
        assert instances_to_coco_json(mine, 42) == ref_fn(ri, 42)


def test_driver_protocol_and_result_file(tmp_path):
    class StubModel:
        def __call__(self, batch):
            return [{"instances": _inst(3, seed=x["image_id"])} for x in batch]

    loader = [[{"image_id": 10 * b + k} for k in range(2)] for b in range(3)]
    ev = COCOResultCollector(str(tmp_path), contiguous_id_to_dataset_id={c: c + 1 for c in range(80)})
    res = inference_on_dataset(StubModel(), loader, DatasetEvaluators([ev]))
    assert res["num_images"] == 6 and res["num_detections"] == 18
    on_disk = json.load(open(tmp_path / "coco_instances_results.json"))
    assert len(on_disk) == 18 and on_disk[0]["image_id"] == 0
    want = instances_to_coco_json(_inst(3, seed=0), 0)[0]
    assert on_disk[0]["bbox"] == pytest.approx(want["bbox"]) and on_disk[0]["category_id"] == want["category_id"] + 1
    w = on_disk[0]["bbox"]
    assert w[2] > 0 and w[3] > 0      # XYWH
