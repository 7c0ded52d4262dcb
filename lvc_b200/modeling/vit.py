"""Descriptor front end of the label-verification step: the DINO ViT-S/8 forward that turns candidate crops into the
descriptors the kNN consumes -- ``crop_features = model(crops)`` in tools/run_nearest_neighbours.py:108-128, where ``model`` is
``torch.hub.load('facebookresearch/dino:main', cfg.QUERY_EXPAND.NN_MODEL)`` (:292-295; ``dino_vits8``: patch 8, dim 384, depth 12,
6 heads, MLP ratio 4, qkv bias, LayerNorm eps 1e-6, erf GELU; forward = CLS token of the final norm).

DINO is a third-party dependency that is not in the reference tree (unpinned ``main``, fetched over the network), so the architecture is
restated from its published definition; the state-dict names are DINO's (``cls_token``, ``pos_embed``, ``patch_embed.proj``,
``blocks.{i}.{norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2}``, ``norm``), so a real checkpoint loads unchanged.  Every linear layer
is a ``lvcb200_gemm_bf16`` launch (bias / residual fused); LayerNorm, GELU, the fused attention and the token assembly are kernels of
``csrc/vit.cu``.  bf16 activations (the residual stream included), fp32 accumulation and statistics.
"""
import math
from typing import Dict

import torch

from .. import _lib, ops


def synthetic_vit_state_dict(dim=384, depth=12, heads=6, patch=8, img=224, mlp_ratio=4, seed=0):
    """Random weights with DINO's init law (trunc_normal std 0.02 for weights / tokens, zero biases, unit LayerNorm)."""
    g = torch.Generator().manual_seed(seed)
    tn = lambda *s: torch.randn(*s, generator=g).clamp_(-2, 2) * 0.02
    n = (img // patch) ** 2
    sd = {"cls_token": tn(1, 1, dim), "pos_embed": tn(1, n + 1, dim),
          "patch_embed.proj.weight": tn(dim, 3, patch, patch) * 5, "patch_embed.proj.bias": torch.zeros(dim),
          "norm.weight": torch.ones(dim), "norm.bias": torch.zeros(dim)}
    for i in range(depth):
        p = f"blocks.{i}."
        sd.update({p + "norm1.weight": torch.ones(dim) + 0.1 * tn(dim) * 50, p + "norm1.bias": tn(dim),
                   p + "attn.qkv.weight": tn(3 * dim, dim) * 3, p + "attn.qkv.bias": tn(3 * dim),
                   p + "attn.proj.weight": tn(dim, dim) * 3, p + "attn.proj.bias": tn(dim),
                   p + "norm2.weight": torch.ones(dim) + 0.1 * tn(dim) * 50, p + "norm2.bias": tn(dim),
                   p + "mlp.fc1.weight": tn(mlp_ratio * dim, dim) * 3, p + "mlp.fc1.bias": tn(mlp_ratio * dim),
                   p + "mlp.fc2.weight": tn(dim, mlp_ratio * dim) * 3, p + "mlp.fc2.bias": tn(dim)})
    return sd


class DinoViT:
    """``model(crops [B,3,S,S] fp32 CUDA, normalised) -> [B, dim] fp32`` like the hub model's forward (inference only)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", heads=6, patch=8, eps=1e-6, attention="tc"):
        _lib.load()
        self.device = torch.device(device)
        sd = {k: v.detach().float() for k, v in state_dict.items()}
        self.dim = sd["cls_token"].shape[-1]
        self.heads, self.patch, self.eps = heads, patch, eps
        if self.dim // heads != 64:
            raise _lib.LvcB200Error("DinoViT: the fused attention kernel is built for head_dim 64 (ViT-S: 384 / 6)")
        self.depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
        dev, bf = self.device, torch.bfloat16
        w = lambda k: sd[k].to(dev, bf).contiguous()
        f = lambda k: sd[k].to(dev).contiguous()
        self.cls, self.pos = f("cls_token").view(-1), f("pos_embed").view(-1, self.dim)
        self.pe_w, self.pe_b = sd["patch_embed.proj.weight"].reshape(self.dim, -1).to(dev, bf).contiguous(), f("patch_embed.proj.bias")
        self.blocks = []
        for i in range(self.depth):
            p = f"blocks.{i}."
            self.blocks.append(dict(n1=(f(p + "norm1.weight"), f(p + "norm1.bias")), qkv=(w(p + "attn.qkv.weight"), f(p + "attn.qkv.bias")),
                                    proj=(w(p + "attn.proj.weight"), f(p + "attn.proj.bias")), n2=(f(p + "norm2.weight"), f(p + "norm2.bias")),
                                    fc1=(w(p + "mlp.fc1.weight"), f(p + "mlp.fc1.bias")), fc2=(w(p + "mlp.fc2.weight"), f(p + "mlp.fc2.bias"))))
        self.norm = (f("norm.weight"), f("norm.bias"))
        self.debug = None
        if attention not in ("tc", "mma"):
            raise ValueError("DinoViT: attention must be 'tc' (tcgen05 / TMEM kernel) or 'mma' (mma.sync kernel)")
        self.attention = attention

    def eval(self):
        return self

    def _ln(self, x, wb, rows=None, ldx=None, out_dtype=torch.bfloat16):
        rows = x.shape[0] if rows is None else rows
        out = torch.empty((rows, self.dim), dtype=out_dtype, device=x.device)
        _lib.check(_lib.load().lvcb200_layernorm(_lib.ptr(x), rows, self.dim, ldx or x.stride(0), _lib.ptr(wb[0]), _lib.ptr(wb[1]), self.eps,
                                                 _lib.ptr(out), _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32, self.dim, _lib.stream_ptr()),
                   "lvcb200_layernorm")
        return out

    @torch.no_grad()
    def __call__(self, crops: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(crops)
        lib = _lib.load()
        crops = crops.detach().float().contiguous()
        B, _, S, _ = crops.shape
        if S % self.patch or (S // self.patch) ** 2 + 1 != self.pos.shape[0]:
            raise _lib.LvcB200Error("DinoViT: crop size does not match pos_embed (the mining path uses 224 x 224 crops; no interpolation)")
        Np, D, H = (S // self.patch) ** 2, self.dim, self.heads
        N = Np + 1
        dev, bf = crops.device, torch.bfloat16
        if B == 0:
            return torch.empty((0, D), dtype=torch.float32, device=dev)
        patches = torch.empty((B * Np, 3 * self.patch ** 2), dtype=bf, device=dev)
        _lib.check(lib.lvcb200_vit_patchify(_lib.ptr(crops), B, S, self.patch, _lib.ptr(patches), _lib.stream_ptr()), "lvcb200_vit_patchify")
        tok = ops.gemm(patches, self.pe_w, bias=self.pe_b)                       # the 8 x 8 / stride-8 conv as one GEMM over patch rows
        x = torch.empty((B * N, D), dtype=bf, device=dev)
        _lib.check(lib.lvcb200_vit_assemble(_lib.ptr(tok), _lib.ptr(self.cls), _lib.ptr(self.pos), B, Np, D, _lib.ptr(x), _lib.stream_ptr()),
                   "lvcb200_vit_assemble")
        scale = 1.0 / math.sqrt(D // H)
        for blk in self.blocks:
            h = self._ln(x, blk["n1"])
            qkv = ops.gemm(h, blk["qkv"][0], bias=blk["qkv"][1])
            a = torch.empty((B * N, D), dtype=bf, device=dev)
            if self.attention == "tc":
                _lib.check(lib.lvcb200_attention_tc(_lib.ptr(qkv), B, N, H, D // H, scale, _lib.ptr(a), _lib.stream_ptr()), "lvcb200_attention_tc")
            else:
                _lib.check(lib.lvcb200_attention(_lib.ptr(qkv), B, N, H, D // H, scale, _lib.ptr(a), _lib.stream_ptr()), "lvcb200_attention")
            x = ops.gemm(a, blk["proj"][0], bias=blk["proj"][1], residual=x)     # x + proj(attn(norm1(x)))
            h = self._ln(x, blk["n2"])
            m = ops.gemm(h, blk["fc1"][0], bias=blk["fc1"][1], relu="gelu")         # erf GELU in the GEMM's epilogue
            x = ops.gemm(m, blk["fc2"][0], bias=blk["fc2"][1], residual=x)       # x + fc2(gelu(fc1(norm2(x))))
        if self.debug is not None:
            self.debug["tokens"] = x.view(B, N, D)
        return self._ln(x, self.norm, rows=B, ldx=N * D, out_dtype=torch.float32)  # final norm on the CLS rows only
