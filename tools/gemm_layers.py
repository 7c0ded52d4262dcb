"""Per-layer table of the detection step's dense launches (batch 8 R101-FPN): for every recorded GEMM descriptor
  seq_us  = duration inside the real launch sequence (CUDA events around each launch, all queued behind a device-side sleep so
            the host never starves the stream; L2 state as in the real step)
  solo_us = the same launch repeated 20x back to back in a CUDA graph (operands L2-warm where they fit)
next to its tensor bound (flops / sustained bf16 peak) and HBM bound (compulsory bytes / copy bandwidth).
Usage: python tools/gemm_layers.py [out.md]"""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lvc_b200 import ops  # noqa: E402
from lvc_b200.modeling import DetectorEngine  # noqa: E402
from lvc_b200.weights import synthetic_state_dict  # noqa: E402

TF, GBS, _, _ = bench.measured_peaks()

cfg = bench.bench_cfg()
eng = DetectorEngine(cfg, synthetic_state_dict(cfg, 0))
ims = bench.make_images(0, bench.BATCH, device="cuda")
for _ in range(2):
    eng.run(ims)
torch.cuda.synchronize()
ops.GEMM_RECORD = []
eng.run(ims)
torch.cuda.synchronize()
rec, ops.GEMM_RECORD = ops.GEMM_RECORD, None


def esz(code):
    return 4 if code == ops.F32 else 2


def unit_rows(ent):
    d, tag = ent
    ds = [x for x, _ in d] if isinstance(d, list) else [d]
    M = N = K = taps = res = 0
    flops = byt = tb = hb = 0.0
    for x in ds:
        M, N, K, taps, res = x.M, x.N, x.K, x.taps, bool(x.residual)
        f = 2.0 * M * N * K * taps
        b = M * K * esz(x.a_dtype) + N * K * taps * esz(x.a_dtype) + M * N * esz(x.d_dtype) + (M * N * 2 if x.residual else 0)
        flops += f; byt += b
        tb += f / TF / 1e6; hb += b / GBS / 1e3          # per-layer bounds add up (layers are dependent)
    if len(ds) > 1:
        return dict(M=ds[0].M, N=-len(ds), K=0, taps=0, res=False, flops=flops, bytes=byt, tb=tb, hb=hb, mb=sum(
            max(2.0 * x.M * x.N * x.K * x.taps / TF / 1e6, (x.M * x.K * 2 + x.N * x.K * x.taps * 2 + x.M * x.N * 2 + (x.M * x.N * 2 if x.residual else 0)) / GBS / 1e3) for x in ds))
    return dict(M=M, N=N, K=K, taps=taps, res=res, flops=flops, bytes=byt, tb=tb, hb=hb, mb=max(tb, hb))


rows = [unit_rows(e) for e in rec]

# in-sequence timing
REPS = 5
acc = [0.0] * len(rec)
for r in range(REPS):
    evs = []
    torch.cuda._sleep(int(40e6))                         # ~20 ms head start for the host
    for d, _ in rec:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.replay_gemms([(d, _)])
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    for i, (e0, e1) in enumerate(evs):
        acc[i] += e0.elapsed_time(e1) * 1e3 / REPS
for i, r in enumerate(rows):
    r["seq_us"] = acc[i]

# solo timing (unique shapes only)
solo = {}
for i, (d, _) in enumerate(rec):
    r = rows[i]
    key = (r["M"], r["N"], r["K"], r["taps"], r["res"], 0 if isinstance(d, list) else d.d_dtype)
    if key not in solo:
        for _ in range(2):
            ops.replay_gemms([(d, _)])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.replay_gemms([(d, _)])
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        solo[key] = e0.elapsed_time(e1) * 1e3 / 60
    r["solo_us"] = solo[key]

# whole-sequence graph replay for reference
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ops.replay_gemms(rec)
g.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
graph_ms = e0.elapsed_time(e1) / 10

lines = [f"# per-layer dense launches, batch {bench.BATCH} R101-FPN (peaks: {TF:.0f} TFLOP/s bf16 sustained, {GBS:.0f} GB/s copy)", "",
         f"{len(rec)} launches; graph replay of the sequence {graph_ms:.3f} ms; sum seq {sum(r['seq_us'] for r in rows) / 1e3:.3f} ms; "
         f"sum solo {sum(r['solo_us'] for r in rows) / 1e3:.3f} ms; sum max(tensor,hbm) bound {sum(r['mb'] for r in rows) / 1e3:.3f} ms; "
         f"sum tensor bound {sum(r['flops'] for r in rows) / TF / 1e9:.3f} ms", "",
         "| # | M | N | K | taps | res | GFLOP | MB | tensor us | hbm us | seq us | solo us | seq TFLOP/s | seq/bound |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
groups = {}
for i, r in enumerate(rows):
    tb, hb = r["tb"], r["hb"]
    lines.append(f"| {i} | {r['M']} | {r['N']} | {r['K']} | {r['taps']} | {int(r['res'])} | {r['flops'] / 1e9:.2f} | {r['bytes'] / 1e6:.1f} | {tb:.1f} | {hb:.1f} | "
                 f"{r['seq_us']:.1f} | {r['solo_us']:.1f} | {r['flops'] / r['seq_us'] / 1e6:.0f} | {r['seq_us'] / r['mb']:.2f} |")
    key = (r["M"], r["N"], r["K"], r["taps"], r["res"])
    gacc = groups.setdefault(key, dict(n=0, seq=0.0, solo=0.0, tb=0.0, hb=0.0, mb=0.0))
    gacc["n"] += 1; gacc["seq"] += r["seq_us"]; gacc["solo"] += r["solo_us"]; gacc["tb"] += tb; gacc["hb"] += hb; gacc["mb"] += r["mb"]
lines += ["", "## by shape", "", "| M | N | K | taps | res | count | seq us total | solo us total | tensor bound | hbm bound | excess over max-bound us |", "|---|---|---|---|---|---|---|---|---|---|---|"]
for key, ga in sorted(groups.items(), key=lambda kv: -kv[1]["seq"]):
    mb = ga["mb"]
    lines.append(f"| {key[0]} | {key[1]} | {key[2]} | {key[3]} | {int(key[4])} | {ga['n']} | {ga['seq']:.0f} | {ga['solo']:.0f} | {ga['tb']:.0f} | {ga['hb']:.0f} | {ga['seq'] - mb:.0f} |")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 1:
    os.makedirs(os.path.dirname(sys.argv[1]) or ".", exist_ok=True)
    open(sys.argv[1], "w").write(out + "\n")
