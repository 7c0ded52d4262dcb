"""tcgen05 shift-GEMM / conv engine parity (run on the B200 with -m gpu), through the C ABI.
Floating-point kernel => the checker is a plain fp32 torch reference of the same op evaluated on the SAME bf16-rounded
operands (fp32 accumulate); tolerance: fp32 outputs 1e-3 of the output scale (north_star's 1e-3 rel), bf16 outputs one
bf16 rounding step (2^-8) of the output scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lvc_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_mm(a, w):
    torch.backends.cuda.matmul.allow_tf32 = False
    return a.float() @ w.float().t()


def _close(got, want, bf16_out):
    scale = float(want.abs().max()) + 1e-12
    tol = (2.0 ** -8 if bf16_out else 1e-3) * scale
    err = float((got.float() - want).abs().max())
    assert err <= tol, f"max err {err:.3e} > tol {tol:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 128, 256), (1000, 1024, 1024), (257, 64, 64), (4096, 512, 2048),
                                   (200, 416, 1024), (5000, 16, 256), (333, 32, 128), (2048, 2048, 512)])
@pytest.mark.parametrize("f32_out", [False, True])
def test_gemm_plain(M, N, K, f32_out):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = (torch.randn(M, K, generator=g)).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias=bias, out_dtype=torch.float32 if f32_out else torch.bfloat16)
    _close(out, _ref_mm(a, w) + bias, not f32_out)


def test_gemm_residual_relu_and_pitches():
    g = torch.Generator(device="cpu").manual_seed(3)
    M, N, K = 777, 256, 192
    abuf = torch.randn(M, K + 64, generator=g).bfloat16().to(DEV)        # A with a row pitch larger than K
    a = abuf[:, :K]
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    res = torch.randn(M, N, generator=g).bfloat16().to(DEV)
    dbuf = torch.full((M, N + 32), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.gemm(a, w, residual=res, relu=True, out=dbuf[:, :N])
    _close(dbuf[:, :N], torch.relu(_ref_mm(a, w) + res.float()), True)
    assert float((dbuf[:, N:] - 7.0).abs().max()) == 0.0                 # nothing written outside D
    # residual with fp32 output and a partial last N tile (residual rides the tensor core as an identity-MMA K step)
    N2 = 416
    w3 = (torch.randn(N2, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    res3 = torch.randn(M, N2, generator=g).bfloat16().to(DEV)
    _close(ops.gemm(a, w3, residual=res3, out_dtype=torch.float32), _ref_mm(a, w3) + res3.float(), False)
    _close(ops.gemm(a, w3, residual=res3, relu=True), torch.relu(_ref_mm(a, w3) + res3.float()), True)
    # K not a multiple of 64 (single tap): TMA zero-fills the tail of the last K block
    a2 = torch.randn(M, 152, generator=g).bfloat16().to(DEV)
    w2 = torch.randn(N, 152, generator=g).bfloat16().to(DEV)
    _close(ops.gemm(a2, w2, out_dtype=torch.float32), _ref_mm(a2, w2), False)


def _conv_weight_to_gemm(w):
    """OIHW -> [O, (kh, kw, I)] K-major, the order of the shift-GEMM taps."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("n,H,W,Cin,Cout", [(2, 25, 42, 128, 256), (1, 50, 84, 256, 256), (3, 13, 21, 64, 64), (2, 16, 16, 512, 512),
                                                 (2, 50, 84, 64, 64), (2, 40, 60, 128, 64)])   # the last two take the TAP3 path (shared A tile)
def test_conv3x3_as_shift_gemm(n, H, W, Cin, Cout):
    g = torch.Generator(device="cpu").manual_seed(H * W)
    x = torch.randn(n, Cin, H, W, generator=g).bfloat16().to(DEV)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).bfloat16().to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    pl = ops.Plane.from_nchw(x)
    PH, PW = H + 2, W + 2
    shifts = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]
    A = pl.t.view(-1, Cin)
    out = ops.gemm(A, _conv_weight_to_gemm(w), bias=bias, relu=True, taps=9, shifts=shifts, K=Cin, plane_hw=(PH, PW))
    o = ops.Plane(out.view(n, PH, PW, Cout), H, W, Cout)
    want = torch.relu(F.conv2d(x.float(), w.float(), bias, padding=1))
    _close(o.to_nchw(), want, True)
    full = out.view(n, PH, PW, Cout).float()
    assert float(full[:, 0].abs().max()) == 0 and float(full[:, -1].abs().max()) == 0        # border re-zeroed
    assert float(full[:, :, 0].abs().max()) == 0 and float(full[:, :, -1].abs().max()) == 0


def test_conv1x1_residual_block_tail():
    """conv3 of a bottleneck: 1x1 conv + folded-BN bias + residual + ReLU on a bordered plane (resnet.py:195-211)."""
    g = torch.Generator(device="cpu").manual_seed(5)
    n, H, W, Cin, Cout = 2, 20, 28, 256, 1024
    x = torch.randn(n, Cin, H, W, generator=g).bfloat16().to(DEV)
    sc = torch.randn(n, Cout, H, W, generator=g).bfloat16().to(DEV)
    w = (torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5).bfloat16().to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    px, ps = ops.Plane.from_nchw(x), ops.Plane.from_nchw(sc)
    out = ops.gemm(px.t.view(-1, Cin), w.view(Cout, Cin), bias=bias, residual=ps.t.view(-1, Cout), relu=True, plane_hw=(H + 2, W + 2))
    want = torch.relu(F.conv2d(x.float(), w.float(), bias) + sc.float())
    _close(ops.Plane(out.view(n, H + 2, W + 2, Cout), H, W, Cout).to_nchw(), want, True)


@pytest.mark.parametrize("M,N,K", [(1000, 600, 1024), (4096, 256, 384), (130, 2400, 384)])
def test_gemm_tf32_operands_f16_out(M, N, K):
    """kNN score GEMM: fp32 operands consumed as TF32 (kind::tf32), fp16 output, bias = -c_s.  Checker: fp32 matmul of the
    tf32-truncated... the tensor core may round or truncate to 10 mantissa bits, so the bound is the rigorous one used by the
    re-rank stage: |err| <= 2^-9 * |a| * |w| per row/col pair (+ fp16 output rounding 2^-11 * |value|)."""
    g = torch.Generator(device="cpu").manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias=bias, out_dtype=torch.float16)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = a @ w.t() + bias
    bound = 2.0 ** -9 * a.norm(dim=1, keepdim=True) * w.norm(dim=1)[None, :] + 2.0 ** -10 * ref.abs() + 1e-6
    assert bool(((out.float() - ref).abs() <= bound).all()), float(((out.float() - ref).abs() / bound).max())
    # and it is genuinely more accurate than bf16 operands would be (typical error ~ 2^-11 / sqrt(K) scale)
    assert float((out.float() - ref).abs().mean()) < 2e-3 * float(ref.abs().mean())


def _bottleneck_stage(n, H, W, cin, bott, cout, nblocks, seed, chain):
    """A res stage as the engine issues it (resnet.py:195-211): per block 1x1 -> 3x3 -> 1x1 + residual (+ 1x1 shortcut conv on the
    first block), all on zero-bordered planes.  Returns every layer output (bf16)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    PH, PW = H + 2, W + 2
    x = ops.Plane.from_nchw(torch.randn(n, cin, H, W, generator=g).bfloat16().to(DEV)).t.view(-1, cin)
    sh = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]

    def wt(o, i):
        return (torch.randn(o, i, generator=g) / i ** 0.5).bfloat16().to(DEV), (torch.randn(o, generator=g) * 0.1).to(DEV)

    params = []
    for b in range(nblocks):
        ci = cin if b == 0 else cout
        params.append(dict(c1=wt(bott, ci), c2=wt(bott, 9 * bott), c3=wt(cout, bott), sc=wt(cout, ci) if b == 0 else None))
    outs = []
    with ops.gemm_chain(chain):
        for p in params:
            o1 = ops.gemm(x, p["c1"][0], bias=p["c1"][1], relu=True, plane_hw=(PH, PW))
            o2 = ops.gemm(o1, p["c2"][0], bias=p["c2"][1], relu=True, taps=9, shifts=sh, K=bott, plane_hw=(PH, PW))
            sc = ops.gemm(x, p["sc"][0], bias=p["sc"][1], plane_hw=(PH, PW)) if p["sc"] is not None else x
            x = ops.gemm(o2, p["c3"][0], bias=p["c3"][1], residual=sc, relu=True, plane_hw=(PH, PW))
            outs += [o1, o2, x] + ([sc] if p["sc"] is not None else [])
    torch.cuda.synchronize()
    return outs


@pytest.mark.parametrize("n,H,W,cin,bott,cout,nblocks", [(2, 25, 42, 256, 64, 256, 3),      # res2-like (BLOCK_N 64 / 256)
                                                         (2, 30, 44, 256, 128, 512, 4),     # res3-like (BLOCK_N 128 / 256)
                                                         (3, 50, 84, 512, 256, 1024, 5),    # res4-like, > 148 tiles per layer
                                                         (1, 13, 21, 1024, 512, 2048, 3),   # res5-like, fewer tiles than SMs
                                                         (2, 9, 11, 64, 64, 192, 2)])       # tiny, partial N tile (192 of 256)
def test_layer_chain_matches_per_layer_launches(n, H, W, cin, bott, cout, nblocks):
    """gemm_chain.cu: one persistent launch with tile-granular dependencies == the same layers launched one by one, bit for bit
    (identical MMA order per tile), and both within bf16 rounding of the fp32 reference (checked by the tests above)."""
    want = _bottleneck_stage(n, H, W, cin, bott, cout, nblocks, 11, chain=False)
    for rep in range(3):                                   # repeated runs reuse the cached plan (counters reset by the run)
        got = _bottleneck_stage(n, H, W, cin, bott, cout, nblocks, 11, chain=True)
        assert len(got) == len(want)
        for i, (a, b) in enumerate(zip(got, want)):
            assert torch.equal(a, b), f"layer output {i} differs (rep {rep}): max |d| {float((a.float() - b.float()).abs().max())}"


def test_layer_chain_rejects_in_chain_overwrite():
    from lvc_b200 import _lib
    a = torch.zeros(256, 64, dtype=torch.bfloat16, device=DEV)
    w = torch.zeros(64, 64, dtype=torch.bfloat16, device=DEV)
    b = torch.zeros(256, 64, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(_lib.LvcB200Error):
        with ops.gemm_chain():
            ops.gemm(a, w, out=b)
            ops.gemm(b, w, out=a)        # overwrites a buffer layer 0 reads


def test_lateral_conv_with_fused_topdown_add():
    """FPN lateral 1x1 conv with the nearest-2x upsampled coarser level added in the epilogue (fpn.py:128-134): against an fp32 torch
    reference on the same bf16 operands (one bf16 rounding), borders re-zeroed."""
    g = torch.Generator(device="cpu").manual_seed(21)
    n, Ht, Wt, Cin, Cout = 2, 13, 21, 512, 256
    H, W = 2 * Ht, 2 * Wt
    x = torch.randn(n, Cin, H, W, generator=g).bfloat16().to(DEV)
    top = torch.randn(n, Cout, Ht, Wt, generator=g).bfloat16().to(DEV)
    w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).bfloat16().to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    px, pt = ops.Plane.from_nchw(x), ops.Plane.from_nchw(top)
    out = ops.gemm(px.t.view(-1, Cin), w, bias=bias, plane_hw=(H + 2, W + 2), upsample_add=pt)
    po = ops.Plane(out.view(n, H + 2, W + 2, Cout), H, W, Cout)
    want = F.conv2d(x.float(), w.float().view(Cout, Cin, 1, 1), bias) + F.interpolate(top.float(), scale_factor=2, mode="nearest")
    _close(po.to_nchw(), want, True)
    full = out.view(n, H + 2, W + 2, Cout).float()
    assert float(full[:, 0].abs().max()) == 0 and float(full[:, -1].abs().max()) == 0 and float(full[:, :, 0].abs().max()) == 0


# ---------------------------------------------------------------------------------------------- strict mode (SPLIT GEMM)
def _pair(x):
    full, S = ops.pair_split(x.float().contiguous())
    return full, S


def _merge(full, S, M):
    return full[:M].float() + full[S:S + M].float()


def _rel64(got, want):
    return float((got.double() - want).norm() / (want.norm() + 1e-300))


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 128, 256), (1000, 1024, 1024), (257, 64, 64), (5000, 16, 256), (200, 416, 1024)])
@pytest.mark.parametrize("f32_out", [False, True])
def test_split_gemm_matches_fp64(M, N, K, f32_out):
    """Strict mode: fp32 operands carried as bf16 hi/lo pairs, three-term product on the bf16 tensor pipe.  Checker: the same
    product in fp64 on the ORIGINAL fp32 operands; tolerance 3e-5 relative L2 (pair representation 2^-18 per operand, dropped
    lo*lo term 2^-18, fp32 accumulation) -- two orders inside north_star's 1e-3, which a single bf16 product (4e-3) misses."""
    if N < 64 and not f32_out:
        pytest.skip("bf16 outputs use the staged epilogue (N >= 64)")
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    af, S = _pair(a)
    out = ops.gemm(af, ops.split_weight(w).to(DEV), bias=bias, M=M, split_rows=S, out_dtype=torch.float32 if f32_out else torch.bfloat16)
    got = out if f32_out else _merge(out, S, M)
    want = a.double() @ w.double().t() + bias.double()
    assert _rel64(got, want) < 3e-5
    plain = ops.gemm(a.bfloat16(), w.bfloat16(), bias=bias, out_dtype=torch.float32)
    assert _rel64(plain, want) > 20 * _rel64(got, want)      # the pair product really is that much tighter than one bf16 product


def test_split_conv3x3_residual_relu_vs_fp32_conv():
    g = torch.Generator(device="cpu").manual_seed(11)
    n, C, H, W, O = 2, 64, 19, 27, 128
    x = torch.randn(n, C, H, W, generator=g).to(DEV)
    w = (torch.randn(O, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(DEV)
    b = torch.randn(O, generator=g).to(DEV)
    r = torch.randn(n, O, H, W, generator=g).to(DEV)
    xp, rp = ops.PairPlane.from_nchw(x), ops.PairPlane.from_nchw(r)
    out = ops.PairPlane(torch.zeros((2 * xp.split_rows, O), dtype=torch.bfloat16, device=DEV), n, H, W, O)
    PW = xp.PW
    ops.gemm(xp.full, ops.split_weight(w.permute(0, 2, 3, 1).reshape(O, -1), 9).to(DEV), bias=b, residual=rp.full, out=out.full, relu=True,
             taps=9, shifts=[(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)], K=C, M=xp.M, plane_hw=(xp.PH, xp.PW),
             split_rows=xp.split_rows)
    want = torch.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1) + r.double())
    assert _rel64(out.to_nchw(), want) < 3e-5
    # the result is again a valid zero-bordered pair plane
    assert float(out.t[:, 0].abs().max()) == 0 and float(out.lo[:, :, 0].abs().max()) == 0 and float(out.t[:, -1].abs().max()) == 0
    # pair helpers round-trip: merge(split(x)) == x to 2^-17
    m = xp.merged()
    assert _rel64(m[:, 1:-1, 1:-1].permute(0, 3, 1, 2), x.double()) < 2 ** -17


def test_pair_elementwise_kernels():
    """maxpool / upsample-add / row_inv_norm / make_rois helper kernels of the strict path against torch."""
    from lvc_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(5)
    n, H, W, C = 2, 10, 14, 64
    top = torch.randn(n, C, H, W, generator=g).to(DEV)
    fine = torch.randn(n, C, 2 * H, 2 * W, generator=g).to(DEV)
    tp, fp = ops.PairPlane.from_nchw(top), ops.PairPlane.from_nchw(fine)
    _lib.check(lib.lvcb200_upsample2_add_pair(_lib.ptr(tp.full), tp.split_rows, n, H, W, C, _lib.ptr(fp.full), fp.split_rows, 2 * H, 2 * W,
                                              _lib.stream_ptr()), "up")
    want = fine.double() + F.interpolate(top, scale_factor=2, mode="nearest").double()
    assert _rel64(fp.to_nchw(), want) < 1e-5
    x = torch.randn(300, 1024, generator=g).to(DEV)
    xf, S = _pair(x)
    got = ops.row_inv_norm(xf, 20.0, 1e-5, lo_off=S * 1024, rows=300)
    torch.testing.assert_close(got, 20.0 / (x.norm(dim=1) + 1e-5), rtol=1e-5, atol=0)
    got = ops.row_inv_norm(x.bfloat16(), 20.0)
    torch.testing.assert_close(got, 20.0 / (x.bfloat16().float().norm(dim=1) + 1e-5), rtol=1e-5, atol=0)
    props = torch.randn(3, 50, 4, generator=g).to(DEV)
    counts = torch.tensor([50, 0, 17], dtype=torch.int32, device=DEV)
    rois, rimg = ops.make_rois(props, counts)
    assert torch.equal(rois[:, 1:], props.view(-1, 4)) and torch.equal(rois[:, 0], torch.arange(3, device=DEV).repeat_interleave(50).float())
    assert torch.equal(rimg.view(3, 50)[2], torch.where(torch.arange(50, device=DEV) < 17, 2, -1).int()) and int(rimg.view(3, 50)[1].max()) == -1


def test_gemm_gelu_epilogue():
    """relu="gelu": exact (erf) GELU fused into the staged epilogue (fc1 of the ViT MLP) against torch's nn.GELU on the fp32 product."""
    g = torch.Generator(device="cpu").manual_seed(12)
    M, N, K = 1570, 1536, 384
    a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5 * 2).bfloat16().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias=bias, relu="gelu")
    _close(out, F.gelu(_ref_mm(a, w) + bias), True)
    with pytest.raises(Exception):
        ops.gemm(a, w[:64], bias=bias[:64], relu="gelu")           # narrow layers have no GELU instantiation


@pytest.mark.parametrize("precision,tol", [("bf16", 1.5e-2), ("strict", 5e-5)])
def test_layer_wrappers_conv2d_linear_vs_torch(precision, tol):
    """lvc_b200.layers.Conv2d / Linear (mirrors of detectron2/layers/wrappers.py:41-105 on the shift-GEMM) against torch's fp32 conv2d /
    linear with the same parameters: 1x1 and 3x3, stride 1 and 2, FrozenBN folded, ReLU fused; relative L2 <= tol (bf16 operands: 1.5e-2,
    strict pair operands: 5e-5; the reference values are computed in fp64)."""
    import torch.nn.functional as F
    from lvc_b200.layers import Conv2d, FrozenBatchNorm2d, Linear
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 64, 37, 45, generator=g).cuda()
    for k, s, bn, act in [(1, 1, False, None), (3, 1, True, F.relu), (3, 2, True, torch.nn.ReLU()), (1, 2, False, torch.tanh)]:
        norm = None
        if bn:
            norm = FrozenBatchNorm2d(96)
            norm.weight.copy_(torch.rand(96, generator=g) + 0.5); norm.bias.copy_(torch.randn(96, generator=g) * 0.1)
            norm.running_mean.copy_(torch.randn(96, generator=g) * 0.1); norm.running_var.copy_(torch.rand(96, generator=g) + 0.5)
        conv = Conv2d(64, 96, kernel_size=k, stride=s, padding=k // 2, bias=not bn, norm=norm, activation=act).cuda()
        conv.precision = precision
        got = conv(x)
        dd = lambda t: None if t is None else t.detach().double()       # fp64 reference (torch's fp32 CUDA conv may itself run in TF32)
        want = F.conv2d(x.double(), dd(conv.weight), dd(conv.bias), stride=s, padding=k // 2)
        if bn:
            want = F.batch_norm(want, dd(norm.running_mean), dd(norm.running_var), dd(norm.weight), dd(norm.bias), False, 0.0, norm.eps)
        if act is not None:
            want = act(want)
        assert got.shape == want.shape and got.dtype == x.dtype
        rel = float((got.double() - want).norm() / want.norm())
        assert rel <= tol, (k, s, bn, rel)
        with torch.no_grad():
            conv.weight.mul_(2.0)                        # in-place parameter change: the packed operands are rebuilt
        want2 = F.conv2d(x.double(), dd(conv.weight), dd(conv.bias), stride=s, padding=k // 2)
        if not bn and act is None:
            assert float((conv(x).double() - want2).norm() / want2.norm()) <= tol
    assert conv(x[:0]).shape == (0, 96, 19, 23)
    lin = Linear(1024, 416).cuda()
    lin.precision = precision
    xf = torch.randn(3, 50, 1024, generator=g).cuda()
    got, want = lin(xf), F.linear(xf.double(), lin.weight.detach().double(), lin.bias.detach().double())
    assert got.shape == (3, 50, 416) and float((got.double() - want).norm() / want.norm()) <= tol
