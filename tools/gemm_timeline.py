"""Per-role clock64 timeline of CTA 0 for one GEMM launch, from an instrumented debug build (LVCB200_LIB=.../liblvcb200_dbg.so; the
instrumentation lives only in that scratch build).  usage: LVCB200_LIB=... python tools/gemm_timeline.py conv3|conv1|c2"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from lvc_b200 import _lib, ops

kind = sys.argv[1] if len(sys.argv) > 1 else "conv3"
M = 35776
PH, PW = 52, 86
g = torch.Generator().manual_seed(0)
if kind == "conv3":
    a = torch.randn(M, 256, generator=g).bfloat16().cuda(); w = torch.randn(1024, 256, generator=g).bfloat16().cuda()
    r = torch.randn(M, 1024, generator=g).bfloat16().cuda(); o = torch.empty(M, 1024, dtype=torch.bfloat16, device="cuda")
    f = lambda: ops.gemm(a, w, residual=r, out=o, relu=True, plane_hw=(PH, PW))
elif kind == "conv1":
    a = torch.randn(M, 1024, generator=g).bfloat16().cuda(); w = torch.randn(256, 1024, generator=g).bfloat16().cuda()
    o = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
    f = lambda: ops.gemm(a, w, out=o, relu=True, plane_hw=(PH, PW))
else:
    a = torch.randn(M, 256, generator=g).bfloat16().cuda(); w = torch.randn(256, 2304, generator=g).bfloat16().cuda()
    o = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
    sh = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]
    f = lambda: ops.gemm(a, w, out=o, relu=True, taps=9, shifts=sh, K=256, plane_hw=(PH, PW))
for _ in range(3):
    f()
torch.cuda.synchronize()
f()
torch.cuda.synchronize()
buf = np.zeros((4, 64, 8), np.uint64)
lib = _lib.load()
lib.lvcb200_debug_dump.argtypes = [ctypes.c_void_p]
assert lib.lvcb200_debug_dump(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t0 = int(buf[0, 0, 0])
rel = lambda v: (int(v) - t0) if int(v) else -1
print(kind, "cycles relative to the producer's first tile start (CTA 0)")
print("tile | prod start, loads issued | mma: wait-tempty start, got tempty, MMAs issued, tfull committed | epi: wait start, got tfull, released | "
      "phase0: wait_read done, bar1, tmem loaded, staged, fenced, bar2")
for t in range(9):
    if not int(buf[1, t, 0]):
        break
    print(t, "|", rel(buf[0, t, 0]), rel(buf[0, t, 1]), "|", *[rel(buf[1, t, i]) for i in range(4)], "|", *[rel(buf[2, t, i]) for i in range(3)],
          "|", *[rel(buf[3, t, i]) for i in range(6)])
