"""Mirrors of the small layer wrappers detectron2/layers/__init__.py exports next to the operators ("operator
signatures to keep"): ``ShapeSpec`` (shape_spec.py), ``cat`` / ``nonzero_tuple`` / ``Conv2d`` / ``Linear`` (wrappers.py:14-118),
``FrozenBatchNorm2d`` / ``get_norm`` (batch_norm.py:13-135).  Same constructor arguments, parameter / buffer names and ``_version``, so
reference state dicts load; ``forward`` runs on liblvcb200:

* ``Conv2d.forward`` (wrappers.py:94-98: conv -> norm -> activation) = one shift-GEMM launch on a zero-bordered channels-last plane,
  FrozenBN folded into weights and bias, ReLU in the epilogue; stride 2 = the stride-1 result subsampled (what a strided conv computes).
  Supported: 1x1 (padding 0) and 3x3 (padding 1), stride 1 / 2, dilation 1, groups 1, channels a multiple of 8 -- every conv of
  ResNet-FPN / RPN / the box heads except the 7x7 stem, which the DetectorEngine runs through its space-to-depth kernel.
  Anything else raises (no CPU or cuDNN fallback).
* ``Linear.forward`` = the same GEMM.
* ``precision``: "bf16" (bf16 operands, fp32 accumulate) or "strict" (bf16 hi / lo pair operands, three-term product: fp32-grade, DESIGN.md
  section 2); a class-level default that an instance may override.

Inference only: parameters are read, never differentiated (the training branches of the reference are out of scope).
The modules are the op-level drop-in; whole models run through DetectorEngine, which keeps activations in plane form between layers
instead of converting NCHW <-> planes around every conv as these wrappers must.
"""
from collections import namedtuple
from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib, ops


class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
    """detectron2/layers/shape_spec.py: a simple structure that contains basic shape specification about a tensor."""

    def __new__(cls, channels=None, height=None, width=None, stride=None):
        return super().__new__(cls, channels, height, width, stride)


def cat(tensors: List[torch.Tensor], dim: int = 0):
    """wrappers.py:14-21: torch.cat, but avoiding the copy for a single-element list."""
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def nonzero_tuple(x: torch.Tensor):
    """wrappers.py:108-118: a 'as_tuple=True' version of torch.nonzero."""
    if x.dim() == 0:
        return x.unsqueeze(0).nonzero().unbind(1)
    return x.nonzero().unbind(1)


class FrozenBatchNorm2d(nn.Module):
    """batch_norm.py:13-92: BatchNorm2d with fixed statistics and affine parameters, all four stored as buffers
    (``weight``, ``bias``, ``running_mean``, ``running_var``; ``_version = 3``).  On the hot path the layer never runs by itself:
    ``Conv2d`` folds ``scale_shift()`` into its GEMM.  ``forward`` on its own applies the same fp32 scale / shift with two torch
    element-wise ops (a pure streaming op, kept for module-level completeness)."""

    _version = 3

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = num_features, eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def scale_shift(self):
        """batch_norm.py:45-52 (the fp32 form): y = x * scale + shift."""
        scale = self.weight.float() * (self.running_var.float() + self.eps).rsqrt()
        return scale, self.bias.float() - self.running_mean.float() * scale

    def forward(self, x):
        scale, shift = self.scale_shift()
        return x * scale.view(1, -1, 1, 1).to(x.dtype) + shift.view(1, -1, 1, 1).to(x.dtype)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        version = local_metadata.get("version", None)
        if version is None or version < 2:      # batch_norm.py:68-76: no running stats before version 2
            state_dict.setdefault(prefix + "running_mean", torch.zeros_like(self.running_mean))
            state_dict.setdefault(prefix + "running_var", torch.ones_like(self.running_var))
        if version is not None and version < 3:  # batch_norm.py:78-84: eps moved out of running_var in version 3
            state_dict[prefix + "running_var"] = state_dict[prefix + "running_var"] - self.eps
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def __repr__(self):
        return "FrozenBatchNorm2d(num_features={}, eps={})".format(self.num_features, self.eps)


def get_norm(norm, out_channels):
    """batch_norm.py:113-135 for the norms of the mining path: "" / None -> None, "FrozenBN" -> FrozenBatchNorm2d, a callable is called.
    The training-time norms (BN, SyncBN, GN) are out of scope and raise."""
    if norm is None or (isinstance(norm, str) and len(norm) == 0):
        return None
    if isinstance(norm, str):
        if norm != "FrozenBN":
            raise _lib.LvcB200Error(f"get_norm: '{norm}' is a training-time norm; the mining path uses FrozenBN only")
        return FrozenBatchNorm2d(out_channels)
    return norm(out_channels)


def _is_relu(act):
    return act is F.relu or act is torch.relu or isinstance(act, nn.ReLU) or act is F.relu_


class _Dense:
    """GEMM operands of one layer, rebuilt when a parameter / buffer changes (version counters)."""

    def __init__(self):
        self.key, self.w, self.b = None, None, None


class Conv2d(nn.Conv2d):
    """wrappers.py:41-99: ``torch.nn.Conv2d`` with the extra keyword arguments ``norm`` and ``activation``."""

    precision = "bf16"

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation
        self._dense = _Dense()

    def _check(self, x):
        k, s, p = self.kernel_size, self.stride, self.padding
        ok = (k in ((1, 1), (3, 3)) and p == (k[0] // 2, k[0] // 2) and s in ((1, 1), (2, 2)) and self.dilation == (1, 1) and self.groups == 1
              and self.in_channels % 8 == 0 and self.out_channels % 8 == 0 and self.padding_mode == "zeros")
        if not ok:
            raise _lib.LvcB200Error(f"Conv2d: unsupported configuration {self.extra_repr()} (supported: 1x1 / 3x3, padding k // 2, stride 1 / 2, "
                                    "dilation 1, groups 1, channels % 8 == 0); no fallback")
        if self.norm is not None and not isinstance(self.norm, FrozenBatchNorm2d):
            raise _lib.LvcB200Error("Conv2d: only FrozenBatchNorm2d can be folded (inference path)")
        if self.training and self.norm is not None:
            pass  # FrozenBN behaves the same in train mode; nothing to differentiate here anyway
        _lib.require_cuda(x)
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"Conv2d: expected [N, {self.in_channels}, H, W], got {tuple(x.shape)}")

    def _operands(self, device):
        srcs = [self.weight, self.bias] + ([self.norm.weight, self.norm.bias, self.norm.running_mean, self.norm.running_var] if self.norm is not None else [])
        key = (self.precision, str(device)) + tuple((t.data_ptr(), t._version) if t is not None else None for t in srcs)
        d = self._dense
        if d.key != key:
            w = self.weight.detach().float()
            b = self.bias.detach().float() if self.bias is not None else torch.zeros(self.out_channels, device=w.device)
            if self.norm is not None:
                scale, shift = self.norm.scale_shift()
                w = w * scale.view(-1, 1, 1, 1).to(w.device)
                b = b * scale.to(b.device) + shift.to(b.device)
            taps = self.kernel_size[0] * self.kernel_size[1]
            wg = w.permute(0, 2, 3, 1).reshape(self.out_channels, -1).contiguous()          # OIHW -> [O, (kh, kw, I)]
            d.w = (ops.split_weight(wg, taps) if self.precision == "strict" else wg.to(torch.bfloat16)).to(device)
            d.b = b.to(device)
            d.key = key
        return d.w, d.b

    @torch.no_grad()
    def forward(self, x):
        self._check(x)
        if x.numel() == 0:      # wrappers.py:66-92: empty inputs keep their (empty) output shape
            ho = (x.shape[2] + 2 * self.padding[0] - self.kernel_size[0]) // self.stride[0] + 1
            wo = (x.shape[3] + 2 * self.padding[1] - self.kernel_size[1]) // self.stride[1] + 1
            return x.new_empty((x.shape[0], self.out_channels, max(ho, 0), max(wo, 0)))
        w, b = self._operands(x.device)
        relu = _is_relu(self.activation)
        strict = self.precision == "strict"
        n, _, H, W = x.shape
        taps = self.kernel_size[0] * self.kernel_size[1]
        if strict:
            p = ops.PairPlane.from_nchw(x)
            out = ops.PairPlane(torch.zeros((2 * p.split_rows, self.out_channels), dtype=torch.bfloat16, device=x.device), n, H, W, self.out_channels)
        else:
            p = ops.Plane.from_nchw(x)
            out = ops.Plane(torch.empty((n, p.PH, p.PW, self.out_channels), dtype=torch.bfloat16, device=x.device), H, W, self.out_channels)
        shifts = [(kh - 1) * p.PW + (kw - 1) for kh in range(3) for kw in range(3)] if taps == 9 else (0,)
        if strict:
            ops.gemm(p.full, w, bias=b, out=out.full, relu=relu, taps=taps, shifts=shifts, K=self.in_channels, M=p.M, plane_hw=(p.PH, p.PW),
                     split_rows=p.split_rows)
        else:
            ops.gemm(p.t.view(-1, self.in_channels), w, bias=b, out=out.t.view(-1, self.out_channels), relu=relu, taps=taps, shifts=shifts,
                     K=self.in_channels, plane_hw=(p.PH, p.PW))
        y = out.to_nchw()
        if self.stride == (2, 2):
            y = y[:, :, ::2, ::2].contiguous()
        y = y.to(x.dtype) if x.dtype.is_floating_point else y
        if self.activation is not None and not relu:
            y = self.activation(y)
        return y


class Linear(nn.Linear):
    """wrappers.py:101-105 (``Linear = torch.nn.Linear``): the same module, its forward on the tcgen05 GEMM (features a multiple of 8)."""

    precision = "bf16"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._dense = _Dense()

    @torch.no_grad()
    def forward(self, x):
        _lib.require_cuda(x)
        if self.in_features % 8 or self.out_features % 8:
            raise _lib.LvcB200Error("Linear: in_features and out_features must be multiples of 8; no fallback")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, self.in_features)
        if x2.shape[0] == 0:
            return x.new_empty(lead + (self.out_features,))
        key = (self.precision, str(x.device), self.weight.data_ptr(), self.weight._version,
               (self.bias.data_ptr(), self.bias._version) if self.bias is not None else None)
        d = self._dense
        if d.key != key:
            w = self.weight.detach().float()
            d.w = (ops.split_weight(w, 1) if self.precision == "strict" else w.to(torch.bfloat16)).to(x.device).contiguous()
            d.b = (self.bias.detach().float() if self.bias is not None else torch.zeros(self.out_features)).to(x.device)
            d.key = key
        if self.precision == "strict":
            a, S = ops.pair_split(x2.float())
            y = ops.gemm(a, d.w, bias=d.b, out_dtype=torch.float32, M=x2.shape[0], split_rows=S)
        else:
            y = ops.gemm(x2.to(torch.bfloat16).contiguous(), d.w, bias=d.b, out_dtype=torch.float32)
        return y.view(lead + (self.out_features,)).to(x.dtype if x.dtype.is_floating_point else torch.float32)
