"""Timing of lvcb200_subsample_labels at a few sizes (CUDA events, 20 calls)."""
import numpy as np
import torch
from lvc_b200.modeling import subsample_labels_batched

def t(V, N, p):
    lab = torch.from_numpy(np.random.default_rng(0).choice(np.array([-1, 0, 1], np.int8), size=(V, N), p=p)).cuda()
    keys = torch.randint(0, 2 ** 32, lab.shape, dtype=torch.int64, device="cuda").to(torch.uint32)
    for _ in range(3):
        subsample_labels_batched(lab, 256, 0.5, 0, keys=keys)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        subsample_labels_batched(lab, 256, 0.5, 0, keys=keys)
    e1.record(); torch.cuda.synchronize()
    print(f"V={V} N={N} p={p}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")

for V, N, p in [(8, 268569, [0.3, 0.699, 0.001]), (1, 268569, [0.3, 0.699, 0.001]), (8, 26856, [0.3, 0.699, 0.001]), (8, 268569, [0.98, 0.019, 0.001]),
                (8, 1000, [0.3, 0.6, 0.1])]:
    t(V, N, p)
