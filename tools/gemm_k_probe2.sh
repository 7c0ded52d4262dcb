#!/bin/bash
# knobs of the GEMM on the ViT shapes (K = 384): epilogue phase width / ring depth
cat > /tmp/p.py <<'PY'
import sys, os, torch
sys.path.insert(0, ".")
from lvc_b200 import ops
M = 803840
for N, K in ((1152, 384), (384, 384), (1536, 384), (384, 1536)):
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    for _ in range(2): ops.gemm(a, w, bias=b, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.gemm(a, w, bias=b, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"  N={N:5d} K={K:5d}: {ms:7.3f} ms {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s")
PY
for cfg in "" "LVCB200_GEMM_PHASE=64" "LVCB200_GEMM_PHASE=64 LVCB200_GEMM_STAGES=4" "LVCB200_GEMM_PHASE=128 LVCB200_GEMM_STAGES=3" "LVCB200_GEMM_PHASE=256" "LVCB200_GEMM_WEPI=1" "LVCB200_GEMM_2CTA=1"; do
  echo "== $cfg"; env $cfg python /tmp/p.py
done
