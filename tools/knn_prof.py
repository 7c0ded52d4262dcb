"""kNN config #4 driver for ncu (launch list / full captures)."""
import sys

import torch

sys.path.insert(0, ".")
from lvc_b200 import ops  # noqa: E402

S, D, Q, ncls = 600, 1024, 200_000, 20
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
means = torch.zeros(ncls, D, device=dev)
means[torch.arange(ncls), torch.arange(ncls)] = 4.0
cls = torch.arange(ncls, device=dev).repeat_interleave(S // ncls)
bank = torch.randn(S, D, generator=g, device=dev) + means[cls]
qcls = torch.randint(0, ncls, (Q,), generator=g, device=dev)
q = torch.randn(Q, D, generator=g, device=dev) + means[qcls]
path = sys.argv[1] if len(sys.argv) > 1 else "tc"
for _ in range(3):
    kb = ops.KnnBank(bank, cls)
    out = kb.verify(q, qcls, path=path)
torch.cuda.synchronize()
print("keep", float(out["keep"].float().mean()))
