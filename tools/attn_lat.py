"""Latency structure of attention_tc: one CTA per SM (147 CTAs) vs two (294) vs many; N chosen for 1, 3 and 7 key blocks."""
import sys
import torch
sys.path.insert(0, ".")
from lvc_b200 import _lib
lib = _lib.load()
def t(B, N, H, reps=20):
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    out = torch.empty((B * N, H * 64), dtype=torch.bfloat16, device="cuda")
    f = lambda: _lib.check(lib.lvcb200_attention_tc(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "tc")
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ctas = ((N + 127) // 128) * H * B
    print(f"B={B} N={N} H={H}: {ctas} CTAs, {(N + 127) // 128} key blocks: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
for N in (128, 384, 896):
    qt = (N + 127) // 128
    for ctas in (147, 294, 588, 1176):
        t(ctas // qt, N, 1)
