"""Is the tcgen05 GEMM bound by its epilogue at small K?  Time D[M, N] = A[M, K] W^T for several K at fixed M, N (bf16 out, bias)."""
import sys
import torch
sys.path.insert(0, ".")
from lvc_b200 import ops
M = 803840
for N in (384, 1152, 1536):
    for K in (64, 128, 384, 768, 1536):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        b = torch.randn(N, device="cuda")
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        for _ in range(2):
            ops.gemm(a, w, bias=b, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            ops.gemm(a, w, bias=b, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tiles = ((M + 127) // 128) * ((N + 255) // 256)
        print(f"N={N:5d} K={K:5d}: {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s  {ms * 1e3 / (tiles / 148):6.2f} us per tile per CTA  out {M * N * 2 / ms / 1e6:6.0f} GB/s")
        del a, w, out
