"""Crop front end of the kNN descriptors (the part that is index arithmetic): the mirror of
``get_crops_qe`` (lvc/data/utils.py:485-519) and ``preprocess_crops`` (tools/run_nearest_neighbours.py:102-105).
The DINO ViT forward that consumes the crops is lvc_b200.modeling.DinoViT (lvc_b200.knn.get_descriptors chains the two)."""
import numpy as np
import torch

from . import _lib


def get_padding(H, W):
    """lvc/data/utils.py:468-482: pad (H, W) to a square, the odd pixel goes left / top."""
    max_d = max(H, W)
    h_padding = (max_d - W) / 2
    v_padding = (max_d - H) / 2
    l_pad = h_padding if h_padding % 1 == 0 else h_padding + 0.5
    t_pad = v_padding if v_padding % 1 == 0 else v_padding + 0.5
    r_pad = h_padding if h_padding % 1 == 0 else h_padding - 0.5
    b_pad = v_padding if v_padding % 1 == 0 else v_padding - 0.5
    return int(l_pad), int(r_pad), int(t_pad), int(b_pad)


def crop_geometry(boxes, H, W, operation="context"):
    """Per box (x1, y1, x2, y2 as produced by ``box.tensor.long()``): the source window after python slicing and the padding
    of the reference, as the int32 [n, 8] table lvcb200_crops_qe consumes."""
    geom = np.zeros((len(boxes), 8), np.int32)
    for i, (x1, y1, x2, y2) in enumerate(np.asarray(boxes, np.int64).tolist()):
        if operation == "pad":
            l_p, r_p, t_p, b_p = get_padding(y2 - y1 + 1, x2 - x1 + 1)
            ys, ye, xs, xe = y1, y2 + 1, x1, x2 + 1
        elif operation == "context":
            l_p, r_p, t_p, b_p = get_padding(y2 - y1 + 1, x2 - x1 + 1)
            y1n, x1n = max(0, y1 - t_p), max(0, x1 - l_p)
            y2n, x2n = min(H, y2 + b_p), min(W, x2 + r_p)
            l_p, r_p, t_p, b_p = get_padding(y2n - y1n + 1, x2n - x1n + 1)
            ys, ye, xs, xe = y1n, y2n + 1, x1n, x2n + 1
        else:
            raise ValueError(operation)
        ys, xs = max(ys, 0), max(xs, 0)                   # boxes are clipped detections: non-negative; slicing clamps the end
        ah, aw = max(min(ye, H) - ys, 0), max(min(xe, W) - xs, 0)
        geom[i] = (ys, xs, ah, aw, t_p, l_p, ah + t_p + b_p, aw + l_p + r_p)
    return geom


def get_crops_qe(image, boxes, operation="context", size=224, mean=None, std=None):
    """image: [3,H,W] (or [1,3,H,W]) uint8 / fp32 CUDA tensor; boxes: [n,4] integer boxes.  Returns [n,3,size,size] fp32
    (normalised when mean / std are given, like preprocess_crops)."""
    _lib.require_cuda(image)
    img = image[0] if image.dim() == 4 else image
    img = img.contiguous() if img.dtype == torch.uint8 else img.float().contiguous()
    _, H, W = img.shape
    geom = torch.from_numpy(crop_geometry(boxes.cpu().numpy() if torch.is_tensor(boxes) else boxes, H, W, operation)).to(img.device)
    n = geom.shape[0]
    out = torch.empty((n, 3, size, size), dtype=torch.float32, device=img.device)
    m = torch.as_tensor(mean, dtype=torch.float32, device=img.device) if mean is not None else None
    s = (1.0 / torch.as_tensor(std, dtype=torch.float32, device=img.device)) if std is not None else None
    rc = _lib.load().lvcb200_crops_qe(_lib.ptr(img), _lib.U8 if img.dtype == torch.uint8 else _lib.F32, H, W, _lib.ptr(geom), n, size,
                                      _lib.ptr(m), _lib.ptr(s), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_crops_qe")
    return out
