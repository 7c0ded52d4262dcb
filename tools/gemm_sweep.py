"""Sweep tile / pipeline knobs of the shift-GEMM on the characteristic res-stage shapes (graph-timed, back to back x20)."""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ".")
    from lvc_b200 import ops

    def bench(name, fn):
        for _ in range(3):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"  {name:34s} {e0.elapsed_time(e1) / 100 * 1e3:7.1f} us")

    gen = torch.Generator().manual_seed(0)

    def t(*shape):
        return torch.randn(*shape, generator=gen).bfloat16().cuda()

    for tag, (n, H, W, cin, bott, cout) in {"res4": (8, 50, 84, 1024, 256, 1024), "res3": (8, 100, 168, 512, 128, 512),
                                            "res2": (8, 200, 336, 256, 64, 256), "res5": (8, 25, 42, 2048, 512, 2048)}.items():
        PH, PW = H + 2, W + 2
        M = n * PH * PW
        x, y1, y2, o = t(M, cin), t(M, bott), t(M, bott), torch.empty(M, cout, dtype=torch.bfloat16, device="cuda")
        w1, w2, w3 = t(bott, cin), t(bott, 9 * bott), t(cout, bott)
        b1, b3 = torch.zeros(bott).cuda(), torch.zeros(cout).cuda()
        sh = [(kh - 1) * PW + (kw - 1) for kh in range(3) for kw in range(3)]
        bench(f"{tag} c1 1x1 K={cin} N={bott}", lambda: ops.gemm(x, w1, bias=b1, out=y1, relu=True, plane_hw=(PH, PW)))
        bench(f"{tag} c2 3x3 K={9 * bott} N={bott}", lambda: ops.gemm(y1, w2, bias=b1, out=y2, relu=True, taps=9, shifts=sh, K=bott, plane_hw=(PH, PW)))
        bench(f"{tag} c3 1x1+res K={bott} N={cout}", lambda: ops.gemm(y2, w3, bias=b3, residual=x, out=o, relu=True, plane_hw=(PH, PW)))
else:
    configs = [{}, {"LVCB200_GEMM_BN": "128"}, {"LVCB200_GEMM_PHASE": "64"}, {"LVCB200_GEMM_STAGES": "2"}, {"LVCB200_GEMM_STAGES": "3"},
               {"LVCB200_GEMM_BN": "128", "LVCB200_GEMM_PHASE": "64"}, {"LVCB200_GEMM_BN": "64"}]
    if len(sys.argv) > 1 and sys.argv[1] == "2cta":
        configs = [{}, {"LVCB200_GEMM_2CTA": "1"}]
    if len(sys.argv) > 1 and sys.argv[1] == "debug":
        configs = [{}, {"LVCB200_GEMM_DEBUG": "1"}, {"LVCB200_GEMM_DEBUG": "2"}, {"LVCB200_GEMM_BN": "64"},
                   {"LVCB200_GEMM_BN": "64", "LVCB200_GEMM_DEBUG": "1"}, {"LVCB200_GEMM_BN": "64", "LVCB200_GEMM_DEBUG": "2"}]
    for env in configs:
        print("== config", env or "default", flush=True)
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, __file__, "child"], env=e)
