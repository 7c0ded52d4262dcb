"""Host logic of lvc_b200/mining.py (CPU): the detection -> crop-window conversion chain of the reference, restated once with
numpy (reference_crop_boxes) and once with the very calls the reference makes."""
import numpy as np
import torch

from lvc_b200.mining import reference_crop_boxes


def _chain_with_library_calls(boxes_out, out_hw, in_hw):
    """The reference's conversion chain written with the very calls it makes (torch.tensor on python lists, numpy float64),
    independent of lvc_b200.mining.reference_crop_boxes."""
    res = []
    for b in boxes_out.tolist():
        arr = np.asarray([b], np.float32)
        arr[:, 2] -= arr[:, 0]; arr[:, 3] -= arr[:, 1]                      # BoxMode.convert XYXY->XYWH on the fp32 array
        js = arr[0].tolist()                                               # JSON
        t = torch.tensor(js)[None, :]                                      # BoxMode.convert on a list: float32 tensor
        t[:, 2] += t[:, 0]; t[:, 3] += t[:, 1]
        xyxy = np.array([t.flatten().tolist()])                            # float64
        sx, sy = in_hw[1] * 1.0 / out_hw[1], in_hw[0] * 1.0 / out_hw[0]
        pts = np.array([(xyxy[0, 0], xyxy[0, 1]), (xyxy[0, 2], xyxy[0, 1]), (xyxy[0, 0], xyxy[0, 3]), (xyxy[0, 2], xyxy[0, 3])])
        pts[:, 0] = pts[:, 0] * sx; pts[:, 1] = pts[:, 1] * sy
        bb = np.concatenate([pts.min(0), pts.max(0)]).clip(min=0)
        bb = np.minimum(bb, [in_hw[1], in_hw[0], in_hw[1], in_hw[0]])
        res.append(torch.as_tensor(bb, dtype=torch.float32))
    f = torch.stack(res) if res else torch.zeros((0, 4))
    return f, f.long()


def test_reference_crop_boxes_chain():
    rng = np.random.default_rng(0)
    p = rng.uniform(0, 1100, (500, 4)).astype(np.float32)
    p[:, 0::2] = np.minimum(p[:, 0::2], 937.0); p[:, 1::2] = np.minimum(p[:, 1::2], 1000.0)      # clipped detections, some on the border
    b = np.stack([np.minimum(p[:, 0], p[:, 2]), np.minimum(p[:, 1], p[:, 3]), np.maximum(p[:, 0], p[:, 2]), np.maximum(p[:, 1], p[:, 3])], 1)
    f, w, ok = reference_crop_boxes(b, (1000, 937), (800, 750))
    fw, ww = _chain_with_library_calls(b, (1000, 937), (800, 750))
    assert np.array_equal(f, fw.numpy()) and np.array_equal(w, ww.numpy())
    assert ok.sum() >= 490 and np.array_equal(ok, ((fw[:, 2] - fw[:, 0] > 1e-5) & (fw[:, 3] - fw[:, 1] > 1e-5)).numpy())


