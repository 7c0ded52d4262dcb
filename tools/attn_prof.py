"""One call of each attention kernel at B crops (for ncu): python tools/attn_prof.py [B]"""
import sys
import torch
sys.path.insert(0, ".")
from lvc_b200 import _lib
lib = _lib.load()
B, N, H = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 785, 6
qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty((B * N, H * 64), dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    _lib.check(lib.lvcb200_attention_tc(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "tc")
    _lib.check(lib.lvcb200_attention(_lib.ptr(qkv), B, N, H, 64, 0.125, _lib.ptr(out), _lib.stream_ptr()), "mma")
torch.cuda.synchronize()
