"""Drop-in for detectron2/layers/nms.py:7,10-29 (``nms`` / ``batched_nms``) on liblvcb200's CUDA path."""
import torch

from .. import _lib

TRICK, VANILLA, REFERENCE_CUDA = 0, 1, -1

_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, )
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _run(boxes, scores, idxs, iou_threshold, mode):
    _lib.require_cuda(boxes, scores, idxs)
    assert boxes.shape[-1] == 4
    lib = _lib.load()
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    b = boxes.detach().to(torch.float32).contiguous()
    s = scores.detach().to(torch.float32).contiguous()
    i = idxs.detach().to(torch.int64).contiguous() if idxs is not None else None
    keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
    num = torch.empty(1, dtype=torch.int64, device=boxes.device)
    nbytes = lib.lvcb200_batched_nms_workspace(n)
    ws = _workspace(nbytes, boxes.device)
    with torch.cuda.device(boxes.device):
        rc = lib.lvcb200_batched_nms(_lib.ptr(b), _lib.ptr(s), _lib.ptr(i), n, float(iou_threshold), int(mode), _lib.ptr(keep),
                                     _lib.ptr(num), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_batched_nms")
    return keep[: int(num.item())]  # the reference API returns a variable-length tensor: one sync, as torchvision does


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms semantics: indices kept, sorted by decreasing score; IoU > threshold suppresses."""
    return _run(boxes, scores, None, iou_threshold, VANILLA)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                mode: int = REFERENCE_CUDA) -> torch.Tensor:
    """Same as detectron2.layers.batched_nms.  ``mode`` selects the reference branch to reproduce bit-exactly:
    REFERENCE_CUDA (default) = what the reference does on a CUDA device (coordinate trick up to 25000 boxes,
    per-class loop above); TRICK / VANILLA force one (the CPU reference run takes TRICK only up to 1000 boxes)."""
    return _run(boxes, scores, idxs, iou_threshold, mode)
