"""Host-side mirrors of the reference's meta-architectures on the mining path (lvc/modeling/meta_arch/rcnn.py):

* ``GeneralizedRCNN``        (rcnn.py:25-333)   candidate sourcing: backbone + RPN + box head -> detections
* ``GeneralizedRCNNRegOnly`` (rcnn.py:336-410)  box corrector: backbone + cascade regression heads on given boxes
* ``ProposalNetwork``        (rcnn.py:413-488)  backbone + RPN -> proposals

Same construction and call contract as the reference --

    model = META_ARCH_REGISTRY.get(name)(cfg)            # cfg: the reference's CfgNode (or a DetectorConfig)
    DetectionCheckpointer(model).load(path)              # -> model.load_state_dict(...): reference parameter names
    model.eval(); model(batched_inputs: list[dict(image: Tensor[3,H,W] (BGR, 0..255), height, width, ...)])
        -> list[dict("instances": Instances(pred_boxes: Boxes, scores, pred_classes))]

-- so that ``inference_on_dataset`` (lvc/evaluation/evaluator.py:117-126) and the tools drive them unchanged.  They are
``nn.Module``s whose parameters / buffers carry the reference's state-dict names; the arithmetic is
the DetectorEngine's liblvcb200 launches, built lazily from the current parameters on the first forward and rebuilt after
``load_state_dict`` / ``.to()``.  Inference only: the training branches of the reference are out of scope.
"""
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
from torch import nn

from .. import ops
from ..config import DetectorConfig
from ..structures import Boxes, Instances
from ..weights import synthetic_corrector_head, synthetic_state_dict
from .corrector import BoxCorrectorHead
from .engine import DetectorEngine
from .postprocessing import detector_postprocess

_BUFFER_SUFFIXES = (".norm.weight", ".norm.bias", ".norm.running_mean", ".norm.running_var")   # FrozenBatchNorm2d buffers, batch_norm.py:34-43


class _Holder(nn.Module):
    """Parameter container standing where a reference sub-module stands (only its state-dict names matter here)."""


def _install(root: nn.Module, sd: Dict[str, torch.Tensor]):
    for name, t in sd.items():
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Holder())
            mod = mod._modules[p]
        t = t.detach().clone()
        if name.endswith(_BUFFER_SUFFIXES) or ".cell_anchors." in name:
            mod.register_buffer(parts[-1], t)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))


def _cell_anchor_buffers(cfg: DetectorConfig):
    """proposal_generator.anchor_generator.cell_anchors.{i} (anchor_generator.py:20-28, 173-208): part of the reference's state dict."""
    return {f"proposal_generator.anchor_generator.cell_anchors.{i}":
            torch.tensor(ops.cell_anchors(s, cfg.anchor_ratios), dtype=torch.float32).view(-1, 4) for i, s in enumerate(cfg.anchor_sizes)}


class _MiningModel(nn.Module):
    """Common part: configuration, reference-named parameters, lazily built engine, device handling."""

    _needs = ("backbone.", "proposal_generator.")      # state-dict prefixes this architecture owns (besides its heads)

    def __init__(self, cfg, state_dict=None, device=None, use_cuda_graph=True, precision="bf16"):
        super().__init__()
        if hasattr(cfg, "MODEL"):                       # the reference's CfgNode (lvc/config/defaults.py)
            device = device or str(cfg.MODEL.DEVICE)
            self.cfg = DetectorConfig.from_reference_cfg(cfg)
        else:
            self.cfg = cfg
        self._device = torch.device(device or "cuda")
        self._use_cuda_graph, self._precision = use_cuda_graph, precision
        self._engine = None
        sd = self._default_state(self.cfg) if state_dict is None else dict(state_dict)
        for k, v in _cell_anchor_buffers(self.cfg).items():
            sd.setdefault(k, v)
        _install(self, sd)
        self.training = False
        if self._device.type == "cuda":
            nn.Module.to(self, self._device)

    # ------------------------------------------------------------------ nn.Module plumbing
    def _default_state(self, cfg):
        return synthetic_state_dict(cfg, 0)             # random init with the reference's init laws, like a freshly built model

    @property
    def device(self):
        return self._device

    def train(self, mode=True):
        if mode:
            raise NotImplementedError(f"lvc_b200.{type(self).__name__} is inference-only (pseudo-label mining path)")
        return super().train(False)

    def _apply(self, fn, *a, **k):                      # .to() / .cuda() / .cpu(): parameters move, the engine is rebuilt on demand
        out = super()._apply(fn, *a, **k)
        p = next(self.parameters(), None)
        if p is not None:
            self._device = p.device
        self._engine = None
        return out

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = dict(state_dict)
        if strict:                                      # checkpoints written before the buffers existed lack them; they are recomputable
            for k, v in _cell_anchor_buffers(self.cfg).items():
                sd.setdefault(k, v)
        res = super().load_state_dict(sd, strict=strict, **kw)
        self._engine = None
        return res

    def _engine_state(self):
        return {k: v.detach().cpu() for k, v in self.state_dict().items()}

    @property
    def engine(self) -> DetectorEngine:
        if self._engine is None:
            self._engine = DetectorEngine(self.cfg, self._engine_state(), self._device, use_cuda_graph=self._use_cuda_graph,
                                          precision=self._precision)
        return self._engine

    def to_device(self, batched_inputs):
        """H2D of the batch (rcnn.py:328): uint8 or float images, pinned sources copy asynchronously."""
        return [x["image"].to(self._device, non_blocking=True) for x in batched_inputs]

    def preprocess_image(self, batched_inputs):
        """rcnn.py:324-333.  Normalisation, padding and batching happen inside the engine's stem kernel; this returns the device images."""
        return self.to_device(batched_inputs)

    @torch.no_grad()
    def forward(self, batched_inputs: List[dict], *args, **kwargs):
        return self.inference(batched_inputs, *args, **kwargs)


class GeneralizedRCNN(_MiningModel):
    def __init__(self, cfg, state_dict=None, device=None, use_cuda_graph=True, precision="bf16"):
        super().__init__(cfg, state_dict, device, use_cuda_graph, precision)
        self._host_ring = {}
        self._img_ring = OrderedDict()
        self.candidate_filter = None    # a lvc_b200.candidates.CandidateFilter: results then carry `candidate_flags` (row a15 / f2)

    def _pack(self, batched_inputs, images, outs, boxes, scores, classes, counts):
        """One [n, 6k+1 (+k)] fp32 block per batch for the single D2H: boxes | scores | classes | (candidate flags) | count."""
        cols = [boxes.view(len(images), -1), scores, classes.float()]
        if self.candidate_filter is not None:
            flags, _ = self.candidate_filter(boxes, scores, classes, counts, outs, [x.get("image_id") for x in batched_inputs])
            cols.append(flags.float())
        cols.append(counts.float()[:, None])
        return torch.cat(cols, dim=1)

    @torch.no_grad()
    def inference_stream(self, batches):
        """Software-pipelined form of the reference's inference loop (``for inputs in data_loader: outputs = model(inputs)``,
        lvc/evaluation/evaluator.py:117-126): yields ``model(inputs)`` for every batch of the iterable, in order, while the
        H2D copy of batch i+1 (copy stream) overlaps the forward of batch i and the packed D2H of its detections.  ``batches``
        is consumed lazily: at most two batches are alive at any time."""
        dev = self._device
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        engine = self.engine

        def stage(batched_inputs):   # H2D of a batch on the copy stream, into one of three persistent device image sets
            # (no allocator traffic in the steady state: a cudaMalloc / cudaFree in the loop stalls the whole device for milliseconds;
            # three sets because the copy of batch i + 2 is enqueued while batch i may still be reading its images).  Sets are keyed
            # by the batch's image shapes and LRU-bounded: a real dataset has many shapes.
            key = tuple((tuple(x["image"].shape), x["image"].dtype) for x in batched_inputs)
            ring = self._img_ring.get(key)
            fresh = ring is None
            if fresh:
                while len(self._img_ring) >= 4:
                    self._img_ring.popitem(last=False)
                ring = self._img_ring[key] = [[[torch.empty(sh, dtype=dt, device=dev) for sh, dt in key] for _ in range(3)], 0]
            else:
                self._img_ring.move_to_end(key)
            images = ring[0][ring[1] % 3]
            ring[1] += 1
            with torch.cuda.stream(copy_stream):
                if fresh:                               # new blocks may recycle an evicted set that a queued forward still reads
                    copy_stream.wait_stream(main)
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
            boxes, scores, classes, rows, counts = engine.run(images, outs)
            packed = self._pack(batched_inputs, images, outs, boxes, scores, classes, counts)
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
            if host.shape[1] > 6 * k + 1:
                inst.candidate_flags = host[i, 6 * k: 7 * k][:c].to(torch.int8)
            res.append({"instances": inst})
        return res

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], detected_instances=None, do_postprocess=True):
        """rcnn.py:177-322 (the ``detected_instances is None`` branch: the path of ``--eval-only``)."""
        if detected_instances is not None:
            raise NotImplementedError("forward_with_given_boxes is not on the mining path")
        images = self.to_device(batched_inputs)
        sizes = [tuple(im.shape[-2:]) for im in images]
        outs = [(int(x.get("height", s[0])), int(x.get("width", s[1]))) for x, s in zip(batched_inputs, sizes)] if do_postprocess else sizes
        boxes, scores, classes, rows, counts = self.engine.run(images, outs)
        # one packed D2H per batch
        host = self._pack(batched_inputs, images, outs, boxes, scores, classes, counts).cpu()
        return self._unpack(host, _Done(), outs, scores.shape[1])


class _Done:
    def synchronize(self):
        pass


class ProposalNetwork(_MiningModel):
    """rcnn.py:413-488: backbone + proposal generator; ``forward`` returns ``[{"proposals": Instances(proposal_boxes,
    objectness_logits)}]`` rescaled to the requested output size (``no_post=True``: the raw per-image proposals)."""

    def _default_state(self, cfg):
        sd = synthetic_state_dict(cfg, 0)
        return {k: v for k, v in sd.items() if not k.startswith("roi_heads.")}

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], no_post=False):
        images = self.to_device(batched_inputs)
        sizes = [tuple(im.shape[-2:]) for im in images]
        props, logits, counts = self.engine.run_proposals(images)
        counts = counts.cpu().tolist()
        results = []
        for i, s in enumerate(sizes):
            inst = Instances(s)
            inst.proposal_boxes = Boxes(props[i, : counts[i]].clone())
            inst.objectness_logits = logits[i, : counts[i]].clone()
            results.append(inst)
        if no_post:
            return results, sizes
        out = []
        for r, x, s in zip(results, batched_inputs, sizes):
            out.append({"proposals": detector_postprocess(r, x.get("height", s[0]), x.get("width", s[1]))})
        return out


class GeneralizedRCNNRegOnly(_MiningModel):
    """rcnn.py:336-410 with ``CascadeROIHeads`` / ``BoxOnlyLayersCascade`` heads (cascade_rcnn.py:167-203): the box corrector.
    ``inference`` regresses the ``gt_boxes`` of every input's ``instances`` through the three cascade stages, stores them as
    ``pred_boxes`` (``pred_classes`` = ``gt_classes``), rescales to the output size and returns the INPUT dicts (minus "image"),
    like the reference does."""

    def __init__(self, cfg, state_dict=None, device=None, use_cuda_graph=False, precision="bf16", num_fc=3, stages=3):
        self._num_fc, self._stages = num_fc, stages
        if hasattr(cfg, "MODEL"):
            self._num_fc = int(cfg.MODEL.ROI_BOX_HEAD.NUM_FC)
            self._stages = len(cfg.MODEL.ROI_BOX_CASCADE_HEAD.IOUS) if hasattr(cfg.MODEL, "ROI_BOX_CASCADE_HEAD") else stages
        super().__init__(cfg, state_dict, device, use_cuda_graph, precision)
        self._head = None

    def _default_state(self, cfg):
        sd = {k: v for k, v in synthetic_state_dict(cfg, 0).items() if not k.startswith("roi_heads.")}
        sd.update(synthetic_corrector_head(cfg, 0, self._num_fc, self._stages))
        return sd

    def _apply(self, fn, *a, **k):
        self._head = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._head = None
        return super().load_state_dict(state_dict, strict, **kw)

    @property
    def head(self) -> BoxCorrectorHead:
        if self._head is None:
            self._head = BoxCorrectorHead(self.cfg, self._engine_state(), self._device, num_fc=self._num_fc, stages=self._stages)
        return self._head

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], detected_instances=None, do_postprocess=True):
        if detected_instances is not None:
            raise NotImplementedError("forward_with_given_boxes is not on the mining path")
        images = self.to_device(batched_inputs)
        sizes = [tuple(im.shape[-2:]) for im in images]
        pyramid, _ = self.engine.run_features(images)
        planes = [pyramid[l] for l in (2, 3, 4, 5)]
        gts = [x["instances"].to(self._device) for x in batched_inputs]
        boxes = self.head(planes, [g.gt_boxes.tensor for g in gts], sizes)
        out = []
        for b, x, s in zip(boxes, batched_inputs, sizes):
            inst = x["instances"]
            inst.set("pred_boxes", Boxes(b.to(inst.gt_classes.device)))
            inst.set("pred_classes", inst.gt_classes)
            r = Instances(s, **inst.get_fields())
            x["instances"] = detector_postprocess(r, x.get("height", s[0]), x.get("width", s[1]))
            x.pop("image", None)
            out.append(x)
        return out


META_ARCHITECTURES = {"GeneralizedRCNN": GeneralizedRCNN, "GeneralizedRCNNRegOnly": GeneralizedRCNNRegOnly,
                      "ProposalNetwork": ProposalNetwork}
