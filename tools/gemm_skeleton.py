import sys, torch
sys.path.insert(0, ".")
from lvc_b200 import ops
M, N, K = 546208, 256, 256
PW = 338
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, 9 * K, device="cuda").bfloat16(); o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
sh = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]
f = lambda: ops.gemm(a, w, out=o, taps=9, shifts=sh, K=K)
for _ in range(3): f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10 * 1e3
iters = 4268 * 36 / 148
print(f"{t:.1f} us per launch, {t / iters * 1965:.0f} cycles per K block per CTA")
