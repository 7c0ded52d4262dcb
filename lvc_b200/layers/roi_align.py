"""Drop-in for detectron2/layers/roi_align.py:14-15,63-108 (``roi_align`` / ``ROIAlign``) on liblvcb200."""
import torch
from torch import nn

from .. import _lib


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    """torchvision.ops.roi_align signature.  input [N,C,H,W] fp32 CUDA, rois [K,5] -> [K,C,ph,pw] fp32."""
    _lib.require_cuda(input, rois)
    assert rois.dim() == 2 and rois.size(1) == 5
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    x = input.detach().to(torch.float32).contiguous()
    r = rois.detach().to(torch.float32).contiguous()
    N, C, H, W = x.shape
    out = torch.empty((r.shape[0], C, ph, pw), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        rc = _lib.load().lvcb200_roi_align_nchw_f32(_lib.ptr(x), N, C, H, W, _lib.ptr(r), r.shape[0], ph, pw, float(spatial_scale),
                                                    int(sampling_ratio), int(bool(aligned)), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_roi_align_nchw_f32")
    return out.to(input.dtype)


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        return roi_align(input, rois.to(dtype=input.dtype), self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")
