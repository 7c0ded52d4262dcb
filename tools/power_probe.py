"""Is the dense stack power-capped?  Replays the step's recorded dense launches (CUDA graph) (a) as isolated bursts after idle gaps and
(b) back to back for ~2 s while sampling SM clock and board power through NVML.  Usage: python tools/power_probe.py"""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lvc_b200 import ops  # noqa: E402
from lvc_b200.modeling import DetectorEngine  # noqa: E402
from lvc_b200.weights import synthetic_state_dict  # noqa: E402

import pynvml  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
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
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    ops.replay_gemms(rec)
g.replay()
torch.cuda.synchronize()


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


bursts = []
for _ in range(8):
    time.sleep(0.2)
    bursts.append(timed(1))
print("burst (1 replay after 200 ms idle) ms:", [round(b, 3) for b in bursts])

samples = []
stop = False


def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.02)


th = threading.Thread(target=sampler)
th.start()
for n in (10, 30, 100, 300):
    t = timed(n)
    print(f"sustained x{n}: {t:.3f} ms / replay")
stop = True
th.join()
clk = sorted(s[0] for s in samples)
pw = sorted(s[1] for s in samples)
print(f"NVML during sustained: sm clock median {clk[len(clk) // 2]} min {clk[0]} max {clk[-1]} MHz; power median {pw[len(pw) // 2]:.0f} max {pw[-1]:.0f} W; {len(samples)} samples")
print("power limit W:", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3)
