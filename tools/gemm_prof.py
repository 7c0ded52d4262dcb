"""Launch the three characteristic res4 GEMM shapes (for ncu --set full): conv1 1x1 (mode 1), conv3 1x1 + residual (mode 2),
conv2 3x3 as 9 shifted taps."""
import sys

import torch

sys.path.insert(0, ".")
from lvc_b200 import ops  # noqa: E402

n, H, W = 8, 50, 84
PH, PW = H + 2, W + 2
M = n * PH * PW
g = torch.Generator().manual_seed(0)
x1024 = torch.randn(M, 1024, generator=g).bfloat16().cuda()
x256 = torch.randn(M, 256, generator=g).bfloat16().cuda()
w1 = torch.randn(256, 1024, generator=g).bfloat16().cuda()
w2 = torch.randn(256, 2304, generator=g).bfloat16().cuda()
w3 = torch.randn(1024, 256, generator=g).bfloat16().cuda()
b256 = torch.zeros(256).cuda()
b1024 = torch.zeros(1024).cuda()
o256 = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
o1024 = torch.empty(M, 1024, dtype=torch.bfloat16, device="cuda")
shifts = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]
for it in range(3):
    ops.gemm(x1024, w1, bias=b256, out=o256, relu=True, plane_hw=(PH, PW))
    ops.gemm(x256, w2, bias=b256, out=o256, relu=True, taps=9, shifts=shifts, K=256, plane_hw=(PH, PW))
    ops.gemm(x256, w3, bias=b1024, residual=x1024, out=o1024, relu=True, plane_hw=(PH, PW))
torch.cuda.synchronize()
print("done")
