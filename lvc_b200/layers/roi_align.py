"""Drop-in for detectron2/layers/roi_align.py:14-15,63-108 (``roi_align`` / ``ROIAlign``) on liblvcb200."""
import torch
from torch import nn

from .. import _lib


def _forward(x, r, ph, pw, spatial_scale, sampling_ratio, aligned):
    N, C, H, W = x.shape
    out = torch.empty((r.shape[0], C, ph, pw), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(x.device):
        rc = _lib.load().lvcb200_roi_align_nchw_f32(_lib.ptr(x), N, C, H, W, _lib.ptr(r), r.shape[0], ph, pw, float(spatial_scale),
                                                    int(sampling_ratio), int(bool(aligned)), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_roi_align_nchw_f32")
    return out


class _ROIAlignFn(torch.autograd.Function):
    """Forward + backward on liblvcb200 (the reference reaches torchvision's autograd kernel pair through
    detectron2/layers/roi_align.py:15; arithmetic spec ROIAlign_cuda.cu:64-139 / :141-306)."""

    @staticmethod
    def forward(ctx, x, r, ph, pw, spatial_scale, sampling_ratio, aligned):
        ctx.save_for_backward(r)
        ctx.cfg = (tuple(x.shape), ph, pw, float(spatial_scale), int(sampling_ratio), bool(aligned))
        return _forward(x, r, ph, pw, spatial_scale, sampling_ratio, aligned)

    @staticmethod
    def backward(ctx, grad_out):
        (r,) = ctx.saved_tensors
        (N, C, H, W), ph, pw, scale, ratio, aligned = ctx.cfg
        g = grad_out.detach().to(torch.float32).contiguous()
        gin = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            rc = _lib.load().lvcb200_roi_align_backward_nchw_f32(_lib.ptr(g), _lib.ptr(r), r.shape[0], N, C, H, W, ph, pw, scale, ratio,
                                                                 int(aligned), _lib.ptr(gin), _lib.stream_ptr())
        _lib.check(rc, "lvcb200_roi_align_backward_nchw_f32")
        return gin, None, None, None, None, None, None


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    """torchvision.ops.roi_align signature.  input [N,C,H,W] fp32 CUDA, rois [K,5] -> [K,C,ph,pw] fp32; differentiable w.r.t. input."""
    _lib.require_cuda(input, rois)
    assert rois.dim() == 2 and rois.size(1) == 5
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    r = rois.detach().to(torch.float32).contiguous()
    if input.requires_grad and torch.is_grad_enabled():
        return _ROIAlignFn.apply(input.to(torch.float32).contiguous(), r, ph, pw, spatial_scale, sampling_ratio, aligned).to(input.dtype)
    x = input.detach().to(torch.float32).contiguous()
    return _forward(x, r, ph, pw, spatial_scale, sampling_ratio, aligned).to(input.dtype)


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
