"""SURVEY 8(f) row 1 (crop front end of the kNN descriptors): oracle and crop geometry against the reference's get_crops_qe
output (tests/golden/crops.npz); the CUDA kernel against both (gpu)."""
import numpy as np
import pytest
import torch

from lvc_b200.crops import crop_geometry, get_crops_qe, get_padding
from oracle import oracle as O


@pytest.mark.parametrize("op", ["pad", "context"])
def test_oracle_matches_reference(golden, op):
    g = golden("crops")
    assert np.array_equal(O.get_crops_qe(g["img"], g["boxes"], op, 32), g["crops_" + op])


def test_padding_rule():
    assert get_padding(5, 8) == (0, 0, 2, 1)      # odd difference: the extra pixel goes on top / left
    assert get_padding(8, 8) == (0, 0, 0, 0) and get_padding(10, 4) == (3, 3, 0, 0)


def test_geometry_is_square_inside_the_image():
    geom = crop_geometry(np.array([[10, 12, 30, 40]]), 60, 90, "pad")
    y0, x0, ah, aw, tp, lp, Hp, Wp = geom[0]
    assert (y0, x0, ah, aw) == (12, 10, 29, 21) and Hp == Wp == 29


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["pad", "context"])
def test_kernel_matches_reference_and_normalises(golden, op):
    g = golden("crops")
    img = torch.from_numpy(g["img"]).cuda()
    got = get_crops_qe(img, g["boxes"], op, size=32)
    assert np.array_equal(got.cpu().numpy(), g["crops_" + op])                     # gather only: bit-exact
    got_f = get_crops_qe(img.float(), g["boxes"], op, size=32)
    assert torch.equal(got_f, got)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    nrm = get_crops_qe(img, g["boxes"], op, size=32, mean=mean, std=std).cpu()
    want = (torch.from_numpy(g["crops_" + op]) - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    torch.testing.assert_close(nrm, want, rtol=1e-5, atol=1e-5)
    big = get_crops_qe(img, g["boxes"], op, size=224)                               # the reference's real crop size
    assert np.array_equal(big.cpu().numpy(), O.get_crops_qe(g["img"], g["boxes"], op, 224))
    odd = get_crops_qe(img, g["boxes"], op, size=30)                                # size % 4 != 0: the generic (per-pixel) kernel
    assert np.array_equal(odd.cpu().numpy(), O.get_crops_qe(g["img"], g["boxes"], op, 30))
    odd_n = get_crops_qe(img, g["boxes"], op, size=30, mean=mean, std=std)
    inv = (1.0 / torch.tensor(std)).view(1, 3, 1, 1).cuda()
    assert torch.equal(odd_n, (odd - torch.tensor(mean).view(1, 3, 1, 1).cuda()) * inv)   # the same (x - mean) * (1 / std) in both kernels
    fast_n = get_crops_qe(img, g["boxes"], op, size=32, mean=mean, std=std)
    assert torch.equal(fast_n, (got - torch.tensor(mean).view(1, 3, 1, 1).cuda()) * inv)
