"""kNN config #4 workload, a few launches of each tensor-core path (for ncu launch lists / counters), plus the fraction of queries
the v2 path hands to the exact resolve / SIMT fallback.  usage: [ncu ...] python tools/knn_prof.py [Q]"""
import sys

import torch

sys.path.insert(0, ".")
from lvc_b200 import ops  # noqa: E402

dev = torch.device("cuda")
S, D, ncls = 600, 1024, 20
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
g = torch.Generator(device=dev).manual_seed(1)
means = torch.zeros(ncls, D, device=dev)
means[torch.arange(ncls), torch.arange(ncls)] = 4.0
cls_all = torch.arange(ncls, device=dev).repeat_interleave(S // ncls)
bank = torch.randn(S, D, generator=g, device=dev) + means[cls_all]
g2 = torch.Generator(device=dev).manual_seed(2)
qcls = torch.randint(0, ncls, (Q,), generator=g2, device=dev)
queries = torch.randn(Q, D, generator=g2, device=dev) + means[qcls]
kb = ops.KnnBank(bank, cls_all)
for path in ("tc3", "tc3", "tc1", "tc1"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = kb.verify(queries, qcls, topk=10, knn=10, path=path)
    e1.record()
    torch.cuda.synchronize()
    print(path, f"{e0.elapsed_time(e1):.3f} ms")
    if path == "tc3":
        al = lambda x: (x + 255) // 256 * 256
        Sq = (Q + 127) // 128 * 128
        off = al(2 * Sq * D * 2) + 2 * al(Q * 4) + 2 * al(Q * 13 * 4) + al(Q * 2)
        flag = ops._ws[("knn", dev.index if dev.index is not None else torch.cuda.current_device())][off:off + Q]
        print("  flagged for exact resolve:", float((flag == 1).float().mean()), " SIMT fallback:", float((flag == 2).float().mean()))
