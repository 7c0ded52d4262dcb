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


# ---------------------------------------------------------------- N > 1: images shard, results gather (gloo, world size 2, CPU)
class _StubModel:
    """Stands in for GeneralizedRCNN: one deterministic Instances per image id."""

    def __call__(self, batch):
        return [{"instances": _inst(1 + x["image_id"] % 3, seed=x["image_id"])} for x in batch]


def _shard_worker(rank, world, port, n_images, ret):
    import torch.distributed as dist
    from lvc_b200.evaluation import inference_shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = inference_shard(n_images)
    batches = [[{"image_id": i} for i in list(mine)[j:j + 2]] for j in range(0, len(mine), 2)]
    out = inference_on_dataset(_StubModel(), batches, COCOResultCollector())
    ret[rank] = (list(mine), out.get("num_images"), [r["image_id"] for r in out.get("results", [])])
    dist.destroy_process_group()


def test_sharded_mining_gathers_in_dataset_order_gloo_world2():
    import torch.multiprocessing as mp
    from lvc_b200.evaluation import inference_shard
    assert [list(inference_shard(7, r, 2)) for r in (0, 1)] == [[0, 1, 2, 3], [4, 5, 6]]          # contiguous ceil(n / W) blocks
    assert [list(inference_shard(5, r, 8)) for r in range(8)] == [[0], [1], [2], [3], [4], [], [], []]
    assert list(inference_shard(0, 0, 2)) == []
    port = 29500 + (os.getpid() % 2000) + 7
    ret = mp.Manager().dict()
    mp.spawn(_shard_worker, args=(2, port, 7, ret), nprocs=2, join=True)
    assert ret[0][0] == [0, 1, 2, 3] and ret[1][0] == [4, 5, 6]
    assert ret[1][1] is None                                              # only rank 0 holds the gathered results
    want = [i for i in range(7) for _ in range(1 + i % 3)]               # every detection, in dataset order
    assert ret[0][1] == 7 and ret[0][2] == want


# ---------------------------------------------------------------- the chained pipeline's collector (stub miner, gloo world 2)
class _StubMiner:
    """Stands in for lvc_b200.mining.PseudoLabelMiner: per image 1 + id % 3 detections, the first id % 2 of them verified.  Has the
    ``inference_stream`` entry point inference_on_dataset looks for."""

    def __call__(self, batch):
        out = []
        for x in batch:
            inst = _inst(1 + x["image_id"] % 3, seed=x["image_id"])
            k = x["image_id"] % 2
            pl = Instances(inst.image_size)
            pl.pred_boxes = Boxes(inst.pred_boxes.tensor[:k].clone())
            pl.pred_classes, pl.scores = inst.pred_classes[:k], inst.scores[:k]
            cand = Instances(inst.image_size)
            cand.gt_boxes = Boxes(inst.pred_boxes.tensor[:1].clone())
            out.append({"instances": inst, "candidates": cand, "pseudo_labels": pl})
        return out

    def inference_stream(self, batches):
        for b in batches:
            yield self(b)


def _miner_worker(rank, world, port, n_images, ret):
    import torch.distributed as dist
    from lvc_b200.evaluation import PseudoLabelCollector, inference_shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = inference_shard(n_images)
    batches = [[{"image_id": i} for i in list(mine)[j:j + 2]] for j in range(0, len(mine), 2)]
    out = inference_on_dataset(_StubMiner(), iter(batches), PseudoLabelCollector({c: c + 1 for c in range(80)}))
    ret[rank] = out
    dist.destroy_process_group()


def test_pseudo_label_collector_gathers_in_dataset_order_gloo_world2():
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + 11
    ret = mp.Manager().dict()
    mp.spawn(_miner_worker, args=(2, port, 9, ret), nprocs=2, join=True)
    assert ret[1] == {}
    r = ret[0]
    assert r["num_images"] == 9 and r["num_detections"] == sum(1 + i % 3 for i in range(9)) and r["num_candidates"] == 9
    anns = r["annotations"]
    assert [a["image_id"] for a in anns] == [1, 3, 5, 7] and [a["id"] for a in anns] == [1, 2, 3, 4] and r["num_pseudo_labels"] == 4
    want = instances_to_coco_json(_inst(2, seed=1), 1)[0]
    a = anns[0]
    assert a["bbox"] == pytest.approx(want["bbox"]) and a["category_id"] == want["category_id"] + 1 and a["iscrowd"] == 0
    assert a["area"] == pytest.approx(a["bbox"][2] * a["bbox"][3])
