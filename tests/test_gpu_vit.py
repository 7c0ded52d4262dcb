"""f1: the DINO ViT-S/8 descriptor front end on the B200 (-m gpu).  DINO is a third-party dependency outside the reference tree
(oracle/vit.py: parity unpinned); the CUDA path is checked against the fp32 restatement with synthetic weights of DINO's names and
shapes, and each new kernel against torch's own op.  bf16 activations: tolerances are bf16-rounding level, stated per check."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lvc_b200 import _lib, ops
from lvc_b200.modeling import DinoViT, synthetic_vit_state_dict
from oracle import oracle as O
from oracle.vit import vit_forward

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _attention(qkv, B, N, H, kernel):
    out = torch.empty((B * N, H * 64), dtype=torch.bfloat16, device=DEV)
    lib = _lib.load()
    if kernel == "tc":
        _lib.check(lib.lvcb200_attention_tc(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "attention_tc")
    else:
        _lib.check(lib.lvcb200_attention(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "attention")
    return out


@pytest.mark.parametrize("kernel", ["tc", "mma"])
@pytest.mark.parametrize("B,N,H", [(2, 785, 6), (3, 100, 6), (1, 64, 2), (2, 17, 1), (5, 128, 1), (2, 257, 3)])
def test_attention_kernel_vs_torch(B, N, H, kernel):
    """Both attention kernels (tcgen05 / TMEM: attention_tc.cu; mma.sync: vit.cu) against torch's fp32 softmax(Q K^T / 8) V on the same bf16
    inputs, incl. ragged token counts (key masking, dropped query rows, tiles that run into the next crop's rows or past the matrix)."""
    g = torch.Generator().manual_seed(B * 1000 + N)
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).bfloat16().to(DEV)
    out = _attention(qkv, B, N, H, kernel)
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    want = ((q @ k.transpose(-2, -1) * 0.125).softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    err = float((out.float() - want).abs().max())
    assert err <= 2 ** -7 * float(want.abs().max()) + 1e-3, err     # P is rounded to bf16 before the P V product, the output once more


def test_attention_tc_large_scores_and_determinism():
    """Peaked rows (one dominant key, scores up to +-60 before the softmax) and repeated launches: no overflow, identical results."""
    B, N, H = 3, 300, 2
    g = torch.Generator().manual_seed(3)
    qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 4).bfloat16().to(DEV)
    a = _attention(qkv, B, N, H, "tc")
    b = _attention(qkv, B, N, H, "tc")
    assert torch.equal(a, b) and bool(torch.isfinite(a.float()).all())
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    want = ((q @ k.transpose(-2, -1) * 0.125).softmax(-1) @ v).transpose(1, 2).reshape(B * N, H * 64)
    assert float((a.float() - want).abs().max()) <= 2 ** -6 * float(want.abs().max()) + 1e-3


def test_layernorm_and_gelu_kernels():
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(1000, 384, generator=g) * 3 + 1).bfloat16().to(DEV)
    gam, bet = torch.randn(384, generator=g).to(DEV), torch.randn(384, generator=g).to(DEV)
    lib = _lib.load()
    out = torch.empty((1000, 384), dtype=torch.float32, device=DEV)
    _lib.check(lib.lvcb200_layernorm(_lib.ptr(x), 1000, 384, 384, _lib.ptr(gam), _lib.ptr(bet), 1e-6, _lib.ptr(out), _lib.F32, 384, _lib.stream_ptr()), "ln")
    torch.testing.assert_close(out, F.layer_norm(x.float(), (384,), gam, bet, 1e-6), rtol=1e-4, atol=1e-4)
    # strided rows (the final norm reads one row per image) and bf16 output
    ob = torch.empty((10, 384), dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.lvcb200_layernorm(_lib.ptr(x), 10, 384, 100 * 384, _lib.ptr(gam), _lib.ptr(bet), 1e-6, _lib.ptr(ob), _lib.BF16, 384, _lib.stream_ptr()), "ln")
    torch.testing.assert_close(ob.float(), F.layer_norm(x.float()[::100], (384,), gam, bet, 1e-6), rtol=2 ** -7, atol=2e-2)
    # another width takes the generic (two-pass) kernel
    x2 = (torch.randn(77, 256, generator=g) * 2 - 1).bfloat16().to(DEV)
    o2 = torch.empty((77, 256), dtype=torch.float32, device=DEV)
    _lib.check(lib.lvcb200_layernorm(_lib.ptr(x2), 77, 256, 256, _lib.ptr(gam), _lib.ptr(bet), 1e-6, _lib.ptr(o2), _lib.F32, 256, _lib.stream_ptr()), "ln")
    torch.testing.assert_close(o2, F.layer_norm(x2.float(), (256,), gam[:256], bet[:256], 1e-6), rtol=1e-4, atol=1e-4)
    y = x.clone().view(-1)
    _lib.check(lib.lvcb200_gelu(_lib.ptr(y), y.numel(), _lib.stream_ptr()), "gelu")
    torch.testing.assert_close(y.view_as(x).float(), F.gelu(x.float()), rtol=2 ** -7, atol=1e-3)


@pytest.mark.parametrize("attention", ["tc", "mma"])
@pytest.mark.parametrize("depth,B", [(2, 4), (12, 3)])
def test_vit_forward_vs_oracle(depth, B, attention):
    """Whole forward (224 x 224 crops -> 384-d descriptors) against the fp32 restatement: bf16 weights / activations through
    `depth` blocks; descriptors agree to < 3e-2 relative L2 and > 0.999 cosine, the token matrix before the final norm to < 3e-2."""
    sd = synthetic_vit_state_dict(depth=depth, seed=depth)
    crops = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(9))
    model = DinoViT(sd, attention=attention)
    model.debug = {}
    got = model(crops.to(DEV))
    col = {}
    want = vit_forward(sd, crops, collect=col)
    rel_t = float((model.debug["tokens"].float().cpu() - col["tokens"]).norm() / col["tokens"].norm())
    rel = float((got.cpu() - want).norm() / want.norm())
    cos = float(F.cosine_similarity(got.cpu(), want, dim=1).min())
    print(f"ViT depth {depth}: tokens rel L2 {rel_t:.2e}, descriptors rel L2 {rel:.2e}, min cosine {cos:.5f}")
    assert got.shape == (B, 384) and rel_t < 3e-2 and rel < 3e-2 and cos > 0.999
    assert model(crops[:0].to(DEV)).shape == (0, 384)


def test_crops_to_descriptors_to_knn_stays_on_device(golden):
    """The label-verification front end end to end on the device (run_nearest_neighbours.py:102-128, 142-162): candidate boxes ->
    context crops (lvcb200_crops_qe) -> ViT descriptors -> KnnBank; the nearest bank row of every query agrees with the oracle
    pipeline (fp32 ViT restatement + C kNN on the same crops) except where the oracle's two best similarities are within 2e-2."""
    from lvc_b200.crops import get_crops_qe
    rng = np.random.default_rng(3)
    img = torch.from_numpy(rng.integers(0, 256, (3, 240, 320)).astype(np.uint8))
    boxes = np.concatenate([rng.integers(0, 150, (24, 2)), rng.integers(160, 239, (24, 2))], 1).astype(np.int64)[:, [0, 1, 2, 3]]
    boxes[:, 2] = np.minimum(boxes[:, 0] + rng.integers(20, 150, 24), 319)
    boxes[:, 3] = np.minimum(boxes[:, 1] + rng.integers(20, 80, 24), 239)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    crops = get_crops_qe(img.to(DEV), torch.from_numpy(boxes), "context", 224, mean, std)
    sd = synthetic_vit_state_dict(depth=4, seed=1)
    model = DinoViT(sd)
    feats = model(crops)
    assert feats.is_cuda and feats.shape == (24, 384)
    bank, q = feats[:16], feats[16:]
    cls = torch.arange(16, device=DEV) % 4
    res = ops.KnnBank(bank, cls).verify(q, cls[:8], topk=10, knn=5)
    want_f = vit_forward(sd, crops.cpu())
    w = O.knn_verify(want_f[:16].numpy(), (np.arange(16) % 4), want_f[16:].numpy(), np.arange(8) % 4, topk=10, knn=5)
    gi, wi = res["top_idx"].cpu().numpy()[:, 0], w["top_idx"][:, 0]
    close = (w["top_sim"][:, 0] - w["top_sim"][:, 1]) < 2e-2
    assert np.all((gi == wi) | close)


def test_patchify_and_assemble_kernels_exact():
    """vit_patchify (crops -> patch rows in the conv weight's (c, iy, ix) column order) and vit_assemble ([cls | tokens] + pos_embed), fast
    (patch 8 / D % 8 == 0) and generic paths, against the same rearrangement in torch: identical bf16 values."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(12)
    for S, P in [(224, 8), (64, 8), (48, 16)]:
        B, gp = 3, S // P
        crops = torch.randn(B, 3, S, S, generator=g).to(DEV)
        out = torch.empty((B * gp * gp, 3 * P * P), dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.lvcb200_vit_patchify(_lib.ptr(crops), B, S, P, _lib.ptr(out), _lib.stream_ptr()), "patchify")
        want = crops.view(B, 3, gp, P, gp, P).permute(0, 2, 4, 1, 3, 5).reshape(B * gp * gp, 3 * P * P).bfloat16()
        assert torch.equal(out, want), (S, P)
    for D in (384, 36):
        B, Np = 2, 49
        tok = torch.randn(B * Np, D, generator=g).bfloat16().to(DEV)
        cls, pos = torch.randn(D, generator=g).to(DEV), torch.randn(Np + 1, D, generator=g).to(DEV)
        x = torch.empty((B * (Np + 1), D), dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.lvcb200_vit_assemble(_lib.ptr(tok), _lib.ptr(cls), _lib.ptr(pos), B, Np, D, _lib.ptr(x), _lib.stream_ptr()), "assemble")
        full = torch.cat([cls.view(1, 1, D).expand(B, 1, D), tok.float().view(B, Np, D)], dim=1) + pos.view(1, Np + 1, D)
        assert torch.equal(x, full.bfloat16().view(-1, D)), D
