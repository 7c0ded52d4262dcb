"""torchrun check of all_gather_bank(backend="p2p") against the NCCL path: identical classes / descriptors for even and ragged shards,
repeated calls (buffer alternation), then the kNN section of the bench with both exchanges.
usage: python -m torch.distributed.run --nproc-per-node N tools/p2p_gather_check.py"""
import json
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
from lvc_b200.evaluation import inference_shard
from lvc_b200.knn import all_gather_bank
dev = torch.device("cuda")
ok = True
for total, D in ((600, 1024), (7, 384), (2393, 384), (world, 64)):
    g = torch.Generator().manual_seed(total)
    desc = torch.randn(total, D, generator=g).to(dev)
    cls = torch.randint(0, 80, (total,), generator=g).to(dev)
    sh = inference_shard(total, rank, world)
    for it in range(3):
        d_in = desc[sh.start:sh.stop] + it
        c0, d0 = all_gather_bank(cls[sh.start:sh.stop], d_in, total=total, backend="nccl")
        c1, d1 = all_gather_bank(cls[sh.start:sh.stop], d_in, total=total, backend="p2p")
        same = torch.equal(c0, c1) and torch.equal(d0, d1) and torch.equal(d1, desc + it) and torch.equal(c1, cls)
        ok = ok and same
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("p2p == nccl on all ranks:", bool(t[0]))
import bench
r = bench.knn_section(rank, world, dev, dist, with_cpu=False)
if rank == 0:
    print(json.dumps(r["tensor_core_path"], indent=1))
dist.destroy_process_group()
