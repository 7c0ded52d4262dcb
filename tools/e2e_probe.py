"""Where does an end-to-end step go?  Times H2D (events on the copy stream), the forward (events on the main stream) and the host
loop per batch for GeneralizedRCNN.inference_stream-like pipelining.  usage: python tools/e2e_probe.py [batches]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lvc_b200.modeling import GeneralizedRCNN  # noqa: E402
from lvc_b200.weights import synthetic_state_dict  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cfg = bench.bench_cfg()
model = GeneralizedRCNN(cfg, synthetic_state_dict(cfg, 0), "cuda", use_cuda_graph=True)
host_images = [im.to(torch.uint8).pin_memory() for im in bench.make_images(0, bench.BATCH)]
batched = [{"image": im, "height": bench.H, "width": bench.W} for im in host_images]
for _ in range(2):
    model(batched)
for _ in model.inference_stream([batched] * 3):
    pass
torch.cuda.synchronize()
# raw H2D bandwidth of one batch, alone
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
imgs = [im.cuda(non_blocking=True) for im in host_images]
e1.record()
torch.cuda.synchronize()
print(f"H2D of one batch alone: {e0.elapsed_time(e1):.3f} ms ({sum(i.numel() for i in host_images) / e0.elapsed_time(e1) / 1e6:.1f} GB/s)")
t_iter = []
t0 = time.perf_counter()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
last = t0
for res in model.inference_stream([batched] * K):
    now = time.perf_counter()
    t_iter.append((now - last) * 1e3)
    last = now
f1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
t_iter.sort()
print(f"e2e {wall / K:.3f} ms/batch (events {f0.elapsed_time(f1) / K:.3f}); host loop per yield: min {t_iter[0]:.2f} median {t_iter[len(t_iter) // 2]:.2f} "
      f"p90 {t_iter[int(len(t_iter) * 0.9)]:.2f} max {t_iter[-1]:.2f}")
