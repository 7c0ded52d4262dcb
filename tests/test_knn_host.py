"""Host logic of the kNN mirror (CPU): vote rule and the world_size-2 gloo all-gather of uneven bank shards."""
import os

import numpy as np
import torch
import torch.multiprocessing as mp

from lvc_b200.knn import all_gather_bank, assemble_tensors, get_nn_class_confirmatory
from lvc_b200.structures import Instances


def test_confirmatory_matches_torch_mode():
    g = torch.Generator().manual_seed(0)
    votes = torch.randint(0, 6, (500, 10), generator=g)
    dt = torch.randint(0, 6, (500,), generator=g)
    for k in (10, 5, 3, 1):
        inst = Instances((1, 1))
        inst.set("top10_shots", votes)
        inst.set("gt_classes", dt)
        get_nn_class_confirmatory([{"instances": inst}], k)
        want = (torch.mode(votes[:, :k], dim=1)[0] == dt).long()
        assert torch.equal(inst.keep, want)
    for row, want in (([3, 1, 3, 1, 2, 2, 5, 5, 7, 9], 1), ([9, 9, 1, 1, 0, 0, 4, 4, 4, 9], 4)):
        inst = Instances((1, 1))
        inst.set("top10_shots", torch.tensor([row]))
        inst.set("gt_classes", torch.tensor([want]))
        get_nn_class_confirmatory([{"instances": inst}], 10)
        assert inst.keep.tolist() == [1]


def test_assemble_tensors_sorts_by_class():
    feats = []
    for cls, n in ((3, 2), (1, 3), (2, 1)):
        inst = Instances((1, 1))
        inst.set("gt_classes", torch.full((n,), cls))
        inst.set("crop_feats", torch.full((n, 4), float(cls)))
        feats.append({"instances": inst})
    c, d = assemble_tensors(feats)
    assert c.tolist() == [1, 1, 1, 2, 3, 3] and d[:, 0].tolist() == [1, 1, 1, 2, 3, 3]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5 if rank == 0 else 2          # uneven shards
    desc = torch.arange(n * 8, dtype=torch.float32).view(n, 8) + 100 * rank
    cls = torch.arange(n, dtype=torch.int64) + (1 << 40) * rank   # exercises the 64-bit class packing
    c, d = all_gather_bank(cls, desc)
    ret[rank] = (c.clone(), d.clone())
    # single-collective form: the support set of 7 rows sharded by the InferenceSampler rule (4 + 3), sizes known to every rank
    from lvc_b200.evaluation import inference_shard
    all_d = torch.arange(7 * 8, dtype=torch.float32).view(7, 8)
    all_c = torch.arange(7, dtype=torch.int64) * (1 << 33)
    mine = inference_shard(7, rank, world)
    calls = []
    orig = dist.all_gather
    dist.all_gather = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    c2, d2 = all_gather_bank(all_c[mine.start:mine.stop], all_d[mine.start:mine.stop], total=7)
    dist.all_gather = orig
    ret[f"single{rank}"] = (c2.clone(), d2.clone(), len(calls), bool(all_gather_bank.last_check))
    try:
        all_gather_bank(all_c[:1], all_d[:1], total=7)
        ret[f"err{rank}"] = False
    except ValueError:
        ret[f"err{rank}"] = True
    dist.destroy_process_group()


def test_all_gather_bank_gloo_world2():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    want_d = torch.cat([torch.arange(40, dtype=torch.float32).view(5, 8), torch.arange(16, dtype=torch.float32).view(2, 8) + 100])
    want_c = torch.cat([torch.arange(5), torch.arange(2) + (1 << 40)])
    for r in (0, 1):
        c, d = ret[r]
        assert torch.equal(d, want_d) and torch.equal(c, want_c)
        c2, d2, ncalls, ok = ret[f"single{r}"]
        assert torch.equal(d2, torch.arange(56, dtype=torch.float32).view(7, 8)) and torch.equal(c2, torch.arange(7) * (1 << 33))
        assert ncalls == 1 and ok            # exactly one collective on the data path (gloo: the list form of all_gather)
        assert ret[f"err{r}"]                # a shard that contradicts the sampler rule is refused before any communication
