"""Host-side mirror of the reference's model-level interface for candidate sourcing:
``GeneralizedRCNN`` (lvc/modeling/meta_arch/rcnn.py:25-333).  Same call contract --

    model(batched_inputs: list[dict(image: Tensor[3,H,W] (BGR, 0..255), height, width, ...)])
        -> list[dict("instances": Instances(pred_boxes: Boxes, scores, pred_classes))]

-- so that ``inference_on_dataset`` (lvc/evaluation/evaluator.py:117-126) can drive it unchanged; the arithmetic is the
DetectorEngine's liblvcb200 launches.  Inference only (``model.training`` is always False): the training branches of the
reference are out of scope."""
from typing import Dict, List

import torch

from ..config import DetectorConfig
from ..structures import Boxes, Instances
from .engine import DetectorEngine


class GeneralizedRCNN:
    def __init__(self, cfg: DetectorConfig, state_dict: Dict[str, torch.Tensor], device="cuda", use_cuda_graph=True):
        self.cfg = cfg
        self.device = torch.device(device)
        self.engine = DetectorEngine(cfg, state_dict, device, use_cuda_graph=use_cuda_graph)
        self.training = False
        self._host_ring = {}
        self._img_ring = {}

    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("lvc_b200.GeneralizedRCNN is inference-only (pseudo-label mining path)")
        return self

    def to_device(self, batched_inputs):
        """H2D of the batch (rcnn.py:328): uint8 or float images, pinned sources copy asynchronously."""
        return [x["image"].to(self.device, non_blocking=True) for x in batched_inputs]

    @torch.no_grad()
    def __call__(self, batched_inputs: List[dict]):
        return self.inference(batched_inputs)

    @torch.no_grad()
    def inference_stream(self, batches):
        """Software-pipelined form of the reference's inference loop (``for inputs in data_loader: outputs = model(inputs)``,
        lvc/evaluation/evaluator.py:117-126): yields ``model(inputs)`` for every batch of the iterable, in order, while the
        H2D copy of batch i+1 (copy stream) overlaps the forward of batch i and the packed D2H of its detections."""
        copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)

        def stage(batched_inputs):   # H2D of a batch on the copy stream, into one of three persistent device image sets
            # (no allocator traffic in the steady state: a cudaMalloc / cudaFree in the loop stalls the whole device for milliseconds;
            # three sets because the copy of batch i + 2 is enqueued while batch i may still be reading its images)
            key = tuple((tuple(x["image"].shape), x["image"].dtype) for x in batched_inputs)
            ring = self._img_ring.get(key)
            if ring is None:
                ring = self._img_ring[key] = [[[torch.empty(sh, dtype=dt, device=self.device) for sh, dt in key] for _ in range(3)], 0]
            images = ring[0][ring[1] % 3]
            ring[1] += 1
            with torch.cuda.stream(copy_stream):
                for dst, x in zip(images, batched_inputs):
                    dst.copy_(x["image"], non_blocking=True)
                ready = copy_stream.record_event()
            return batched_inputs, images, ready

        it = iter(batches)
        first = next(it, None)
        staged = stage(first) if first is not None else None
        pending = None   # (host tensor, event, outs, k, keepalive)
        while staged is not None:
            batched_inputs, images, ready = staged
            # the NEXT batch's copies are enqueued before this batch's forward is launched: if the two streams happen to share a hardware
            # queue, the copy then sits ahead of the forward's ~70 graph nodes instead of behind them (that false serialisation showed
            # up as 8-10 ms steps in some runs)
            nxt = next(it, None)
            staged = stage(nxt) if nxt is not None else None
            sizes = [tuple(im.shape[-2:]) for im in images]
            outs = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)]
            main.wait_event(ready)
            boxes, scores, classes, rows, counts = self.engine.run(images, outs)
            packed = torch.cat([boxes.view(len(images), -1), scores, classes.float(), counts.float()[:, None]], dim=1)
            # two persistent pinned result buffers per shape, used alternately (a fresh pinned allocation per batch is a cudaHostAlloc:
            # milliseconds, and it serialises with the device on some hosts)
            key = (tuple(packed.shape), packed.dtype)
            ring = self._host_ring.setdefault(key, [torch.empty(packed.shape, dtype=packed.dtype, pin_memory=True) for _ in range(2)] + [0])
            host = ring[ring[2] & 1]
            ring[2] += 1
            host.copy_(packed, non_blocking=True)
            done = main.record_event()
            if pending is not None:
                yield self._unpack(*pending[:4])
            pending = (host, done, outs, scores.shape[1], images)
        if pending is not None:
            yield self._unpack(*pending[:4])

    @staticmethod
    def _unpack(host, done, outs, k):
        done.synchronize()
        res = []
        for i, o in enumerate(outs):
            c = int(host[i, -1])
            inst = Instances(o)
            inst.pred_boxes = Boxes(host[i, : 4 * k].view(k, 4)[:c].clone())
            inst.scores = host[i, 4 * k: 5 * k][:c].clone()
            inst.pred_classes = host[i, 5 * k: 6 * k][:c].to(torch.int64)
            res.append({"instances": inst})
        return res

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], do_postprocess=True):
        images = self.to_device(batched_inputs)
        sizes = [tuple(im.shape[-2:]) for im in images]
        outs = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)] if do_postprocess else sizes
        boxes, scores, classes, rows, counts = self.engine.run(images, outs)
        # one packed D2H per batch
        host = torch.cat([boxes.view(len(images), -1), scores, classes.float(), counts.float()[:, None]], dim=1).cpu()
        k = scores.shape[1]
        res = []
        for i, o in enumerate(outs):
            c = int(host[i, -1])
            inst = Instances(o)
            inst.pred_boxes = Boxes(host[i, : 4 * k].view(k, 4)[:c].clone())
            inst.scores = host[i, 4 * k: 5 * k][:c].clone()
            inst.pred_classes = host[i, 5 * k: 6 * k][:c].to(torch.int64)
            res.append({"instances": inst})
        return res
