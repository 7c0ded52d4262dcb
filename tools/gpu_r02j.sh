#!/bin/bash
# tcgen05 attention: tests (bounded), then the descriptor bench with both kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_vit.py -x -q 2>&1 | tail -15
timeout 300 python - <<'PY' 2>&1 | tail -12
import torch, time
from lvc_b200 import _lib
lib = _lib.load()
B, N, H = 1024, 785, 6
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.0).bfloat16()
out = torch.empty((B * N, H * 64), dtype=torch.bfloat16, device="cuda")
def tc(): _lib.check(lib.lvcb200_attention_tc(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "tc")
def mma(): _lib.check(lib.lvcb200_attention(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "mma")
for name, fn in (("tc", tc), ("mma", mma)):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 4.0 * B * H * N * N * 64
    print(f"attention {name}: {ms:.3f} ms per layer call ({fl / ms / 1e9:.0f} TFLOP/s algorithmic)")
PY
timeout 300 python tools/vit_prof.py 2>&1 | tail -12
