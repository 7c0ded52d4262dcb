"""fp32 CPU restatement of the DINO ViT forward (TEST INFRASTRUCTURE ONLY).

The model is a third-party dependency of the reference that is NOT in /root/reference: tools/run_nearest_neighbours.py:292-293 fetches
``torch.hub.load('facebookresearch/dino:main', 'dino_vits8')`` (unpinned ``main``) over the network, which is unavailable here.  This
restates its published architecture (vision_transformer.py of facebookresearch/dino: PatchEmbed conv k = s = patch, cls token, learned
pos_embed, pre-norm blocks ``x + attn(norm1(x))``, ``x + mlp(norm2(x))`` with erf GELU, LayerNorm eps 1e-6, output = CLS token of the
final norm).  PARITY UNPINNED for this sub-path: no golden vectors of the real checkpoint exist offline; the CUDA path is checked against
this restatement with synthetic weights of DINO's names / shapes, and its building blocks (attention, LayerNorm, GELU) against torch's.
"""
import torch
import torch.nn.functional as F


def vit_forward(sd, crops, heads=6, patch=8, eps=1e-6, collect=None):
    """sd: DINO-named state dict (fp32); crops [B,3,S,S] fp32 -> [B, dim]."""
    with torch.no_grad():
        sd = {k: v.float() for k, v in sd.items()}
        dim = sd["cls_token"].shape[-1]
        x = F.conv2d(crops.float(), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch)     # [B, dim, gp, gp]
        B = x.shape[0]
        x = x.flatten(2).transpose(1, 2)
        x = torch.cat([sd["cls_token"].expand(B, -1, -1), x], dim=1) + sd["pos_embed"]
        depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
        hd = dim // heads
        for i in range(depth):
            p = f"blocks.{i}."
            h = F.layer_norm(x, (dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
            qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)
            attn = (qkv[0] @ qkv[1].transpose(-2, -1)) * hd ** -0.5
            a = (attn.softmax(dim=-1) @ qkv[2]).transpose(1, 2).reshape(B, -1, dim)
            x = x + F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
            h = F.layer_norm(x, (dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
            x = x + F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        if collect is not None:
            collect["tokens"] = x
        return F.layer_norm(x, (dim,), sd["norm.weight"], sd["norm.bias"], eps)[:, 0]
