"""GPU diagnostic for the tcgen05 shift-GEMM: per-shape error statistics and timing (prints; never asserts).
Run on the GPU box before pytest so that a descriptor / swizzle mistake is characterised, not just detected."""
import sys
import time

import torch

sys.path.insert(0, ".")
from lvc_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False


def run(M, N, K, f32=True):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().cuda()
    try:
        out = ops.gemm(a, w, out_dtype=torch.float32 if f32 else torch.bfloat16)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"M={M} N={N} K={K}: EXCEPTION {e}")
        return False
    ref = a.float() @ w.float().t()
    err = (out.float() - ref).abs()
    scale = float(ref.abs().max())
    mx = float(err.max())
    print(f"M={M} N={N} K={K} f32={f32}: max_err={mx:.3e} scale={scale:.3e} rel={mx / scale:.2e}", end="")
    if mx / scale > 1e-2:
        bad = (err > 1e-2 * scale)
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"  BAD frac={float(bad.float().mean()):.3f} rows[{int(rows.min())}..{int(rows.max())}] n={len(rows)} "
              f"cols[{int(cols.min())}..{int(cols.max())}] n={len(cols)}")
        print("   out[0,:8]", out[0, :8].tolist())
        print("   ref[0,:8]", ref[0, :8].tolist())
        return False
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.gemm(a, w, out=out)
    t0.record()
    for _ in range(10):
        ops.gemm(a, w, out=out)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"  {ms * 1e3:.1f} us  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    return True


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    ok = True
    for shape in [(128, 256, 64), (128, 256, 256), (256, 256, 512), (128, 16, 64), (128, 64, 64), (128, 128, 128),
                  (1000, 1024, 1024), (8192, 256, 2304), (33600, 1024, 256), (33600, 256, 1024), (8000, 1024, 12544),
                  (8192, 8192, 8192)]:
        ok = run(*shape) and ok
    print("GEMM_DIAG", "OK" if ok else "FAILED")
