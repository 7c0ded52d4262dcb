"""Host-side mirror of the label-verification step of tools/run_nearest_neighbours.py:

    assemble_tensors (:131-139), the support-bank all-gather (:303-309), run_nearest_neighbours (:142-162),
    get_nn_class_confirmatory (:214-227)

with the same names, argument meaning and in-place mutation of the per-image ``Instances`` (fields ``crop_feats``,
``gt_classes`` in; ``top10_shots``, ``keep`` out).  The arithmetic runs in liblvcb200's kNN kernels: all images'
queries are batched into ONE device call instead of a per-image CPU broadcast.  The reference gathers pickled python
objects over gloo; here the bank shard is a device tensor all-gathered by torch.distributed (NCCL over NVLink on GPUs,
gloo in the CPU tests), uneven shards handled by a size exchange + padding.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def assemble_tensors(shot_features: List[dict]) -> Tuple[torch.Tensor, torch.Tensor]:
    """tools/run_nearest_neighbours.py:131-139: concatenate and sort the support descriptors by class."""
    classes = torch.cat([x["instances"].gt_classes for x in shot_features])
    sorter = classes.argsort()
    desc = torch.cat([x["instances"].crop_feats for x in shot_features])
    return classes[sorter], desc[sorter]


def all_gather_bank(shot_classes: torch.Tensor, shot_descriptors: torch.Tensor, group=None, total: Optional[int] = None,
                    backend: str = "auto"):
    """The ONE exchange step of the path (run_nearest_neighbours.py:303-309): every rank ends with the full bank,
    rank-major order (== torch.cat(comm.all_gather(...))).

    ``total`` = the size of the whole support set, which every rank knows (it sharded the support DATASET by the
    ``InferenceSampler`` rule, distributed_sampler.py:191-194): rank r then holds ``len(inference_shard(total, r, W))`` rows, the
    shard is padded to ``ceil(total / W)`` rows, and the exchange is a SINGLE ``all_gather_into_tensor`` of one
    ``[1 + ceil(total/W), D + 2]`` fp32 block per rank -- descriptors, bit-cast int64 classes, and a header row carrying the
    shard's row count -- with no host synchronisation anywhere (sizes are not read back: the header only feeds a device-side
    consistency flag, ``all_gather_bank.last_check``).  Without ``total`` (arbitrary uneven shards) a size exchange comes first.

    ``backend="auto"`` (default) takes the peer-memory route below when it is available on every rank (CUDA tensors, ``total`` given,
    at most 16 ranks, symmetric memory rendezvous succeeded -- agreed on once per (group, size) with an all-reduce) and NCCL otherwise.
    ``backend="p2p"`` (CUDA, needs ``total``): no NCCL launch at all -- the padded block is written into a buffer that torch's symmetric
    memory maps into every rank, a device-side barrier orders the reads behind the writes, and ``lvcb200_gather_rows_p2p`` pulls the peers'
    rows over NVLink / NVSwitch with plain loads straight into the concatenated bank (``all_gather_bank_p2p``)."""
    if backend == "p2p":
        return all_gather_bank_p2p(shot_classes, shot_descriptors, group, total)
    if backend not in ("nccl", "auto"):
        raise ValueError("all_gather_bank: backend must be 'auto', 'nccl' or 'p2p'")
    if backend == "auto" and _p2p_usable(shot_descriptors, group, total):
        return all_gather_bank_p2p(shot_classes, shot_descriptors, group, total)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return shot_classes, shot_descriptors
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = shot_descriptors.device
    D = shot_descriptors.shape[1]
    n_mine = shot_descriptors.shape[0]
    if total is not None:
        from .evaluation import inference_shard
        sizes = [len(inference_shard(total, r, world)) for r in range(world)]
        if sizes[rank] != n_mine:
            raise ValueError(f"all_gather_bank: rank {rank} holds {n_mine} rows, the InferenceSampler rule gives {sizes[rank]} of {total}")
        cap = max(sizes)
    else:
        n = torch.tensor([n_mine], dtype=torch.int64, device=dev)
        gathered = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(gathered, n, group=group)
        sizes = [int(s.item()) for s in gathered]
        cap = max(sizes)
    # one padded block carries the header, the descriptors and the (bit-cast) classes: a single collective on the data path
    pack = torch.zeros((cap + 1, D + 2), dtype=torch.float32, device=dev)
    pack[0, :1].fill_(float(n_mine))          # (a fill kernel: CUDA-graph capturable, unlike an element assignment from a host scalar)
    pack[1:1 + n_mine, :D] = shot_descriptors.float()
    if n_mine:
        pack[1:1 + n_mine, D:] = shot_classes.to(torch.int64).view(-1, 1).view(torch.float32).view(-1, 2)
    out = torch.empty((world, cap + 1, D + 2), dtype=torch.float32, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(out.view(world * (cap + 1), D + 2), pack, group=group)
    else:
        dist.all_gather(list(out.unbind(0)), pack, group=group)
    if not (dev.type == "cuda" and torch.cuda.is_current_stream_capturing()):   # (a host list cannot be uploaded inside a CUDA-graph capture)
        all_gather_bank.last_check = (out[:, 0, 0] == torch.tensor(sizes, dtype=torch.float32, device=dev)).all()   # device flag, not synced
    desc = torch.cat([out[r, 1:1 + sizes[r], :D] for r in range(world)])
    cls = torch.cat([out[r, 1:1 + sizes[r], D:].contiguous().view(torch.int64).view(-1) for r in range(world)])
    return cls, desc


all_gather_bank.last_check = None

_P2P_STATE = {}   # (group id, cap, D) -> [symmetric buffers (two, used alternately), handles, call counter]
_P2P_OK = {}      # (group id, cap, D, device) -> bool: every rank could set the peer-memory exchange up


def _p2p_usable(shot_descriptors: torch.Tensor, group, total) -> bool:
    """Whether all_gather_bank_p2p can serve this call -- decided collectively the first time a (group, size) pair is seen (every rank
    tries the symmetric-memory set-up, an all-reduce takes the minimum), never during a CUDA-graph capture of an unseen pair."""
    import os
    if total is None or not shot_descriptors.is_cuda or os.environ.get("LVCB200_KNN_P2P", "1") == "0":
        return False
    if not (dist.is_available() and dist.is_initialized()):
        return False
    grp = group if group is not None else dist.group.WORLD
    world = dist.get_world_size(grp)
    if world == 1 or world > 16 or dist.get_backend(grp) != "nccl":
        return False
    from .evaluation import inference_shard
    cap = max(len(inference_shard(total, r, world)) for r in range(world))
    key = (id(grp), cap, shot_descriptors.shape[1], shot_descriptors.device.index)
    if key in _P2P_OK:
        return _P2P_OK[key]
    if torch.cuda.is_current_stream_capturing():
        return False
    ok = 1
    try:
        _p2p_state(grp, cap, shot_descriptors.shape[1], shot_descriptors.device)
    except Exception:  # noqa: BLE001  (no symmetric memory on this platform / build: every rank falls back to NCCL together)
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=shot_descriptors.device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=grp)
    _P2P_OK[key] = bool(int(flag.item()))
    return _P2P_OK[key]


def _p2p_state(grp, cap, D, dev):
    """Two symmetric buffers [1 + cap, D + 2] (used alternately) and their handles for this (group, size) pair; allocated and
    rendezvoused on first use (a collective, host-synchronising step)."""
    import torch.distributed._symmetric_memory as symm
    key = (id(grp), cap, D, dev.index)
    st = _P2P_STATE.get(key)
    if st is None:
        bufs = [symm.empty((cap + 1, D + 2), dtype=torch.float32, device=dev) for _ in range(2)]
        hdls = [symm.rendezvous(b, grp) for b in bufs]
        for b in bufs:
            b.zero_()
        st = _P2P_STATE[key] = [bufs, hdls, 0]
        torch.cuda.synchronize(dev)
        dist.barrier(grp)
    return st


def all_gather_bank_p2p(shot_classes: torch.Tensor, shot_descriptors: torch.Tensor, group=None, total: Optional[int] = None):
    """``all_gather_bank`` over NVLink peer memory (see there).  Two symmetric buffers are used alternately so that ONE device-side
    barrier per call suffices: a rank that passes the barrier of call n + 1 has finished its reads of call n (stream order), so buffer
    n mod 2 may be rewritten in call n + 2.  The first call for a (group, size) pair allocates and rendezvouses the buffers; later
    calls are stream-ordered and CUDA-graph capturable (a captured call adds a second barrier in front of the write, because every
    replay reuses the buffer chosen at capture time)."""
    import ctypes

    from . import _lib
    from .evaluation import inference_shard
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return shot_classes, shot_descriptors
    if total is None:
        raise ValueError("all_gather_bank_p2p needs `total` (the shard sizes follow from the InferenceSampler rule)")
    _lib.require_cuda(shot_descriptors, shot_classes)
    grp = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(grp), dist.get_rank(grp)
    if world > 16:
        raise _lib.LvcB200Error("all_gather_bank_p2p: at most 16 ranks (one NVLink domain)")
    dev = shot_descriptors.device
    D, n_mine = shot_descriptors.shape[1], shot_descriptors.shape[0]
    sizes = [len(inference_shard(total, r, world)) for r in range(world)]
    if sizes[rank] != n_mine:
        raise ValueError(f"all_gather_bank_p2p: rank {rank} holds {n_mine} rows, the InferenceSampler rule gives {sizes[rank]} of {total}")
    st = _p2p_state(grp, max(sizes), D, dev)
    bufs, hdls, n = st
    st[2] = n + 1
    pack, hdl = bufs[n & 1], hdls[n & 1]
    if torch.cuda.is_current_stream_capturing():
        hdl.barrier()        # a captured call is replayed on the SAME buffer every time: wait until the peers have read the previous replay
    pack[0, :1].fill_(float(n_mine))
    pack[1:1 + n_mine, :D] = shot_descriptors.float()
    if n_mine:
        pack[1:1 + n_mine, D:] = shot_classes.to(torch.int64).view(-1, 1).view(torch.float32).view(-1, 2)
    hdl.barrier()                                         # every rank's shard is in its buffer (device-side, on the current stream)
    out = torch.empty((total, D + 2), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs[:world]])
    rows = (ctypes.c_int32 * world)(*sizes)
    _lib.check(_lib.load().lvcb200_gather_rows_p2p(ptrs, world, rows, D + 2, _lib.ptr(out), _lib.stream_ptr()), "lvcb200_gather_rows_p2p")
    desc = out[:, :D].contiguous()
    cls = out[:, D:].contiguous().view(torch.int64).view(-1)
    return cls, desc


def get_descriptors(model, image, boxes, mean, std, operation="context", size=224):
    """The per-image body of get_descriptors (tools/run_nearest_neighbours.py:108-128) without leaving the device: context crops of the
    boxes (lvc/data/utils.py:485-519 via lvcb200_crops_qe, normalised like preprocess_crops :102-105) -> ``model(crops)`` (the DINO ViT,
    lvc_b200.modeling.DinoViT) -> crop_feats [n, dim] fp32 on the device, ready for assemble_tensors / KnnBank."""
    from .crops import get_crops_qe
    return model(get_crops_qe(image, boxes, operation, size, mean, std))


def run_nearest_neighbours(shot_classes, shot_descriptors, query_features: List[dict], cosine: bool = True,
                           device: Optional[torch.device] = None, topk: int = 10):
    """tools/run_nearest_neighbours.py:142-162.  Sets ``top10_shots`` ([Qi, 10] int64 class votes, best first) on every
    image's ``Instances`` and returns the list, like the reference."""
    if not query_features:
        return query_features
    device = device or (shot_descriptors.device if shot_descriptors.is_cuda else torch.device("cuda"))
    bank = ops.KnnBank(shot_descriptors.to(device), shot_classes.to(device), cosine=cosine)   # cosine=False: -cdist ranking (:154-159)
    counts = [len(d["instances"].get("crop_feats")) for d in query_features]
    feats = torch.cat([d["instances"].get("crop_feats") for d in query_features]).to(device)
    dt = [d["instances"].gt_classes if d["instances"].has("gt_classes") else torch.zeros(c, dtype=torch.int64)
          for d, c in zip(query_features, counts)]
    qcls = torch.cat(dt).to(device)
    res = bank.verify(feats, qcls, topk=topk, knn=topk)
    votes = res["votes"].cpu()
    idx = res["top_idx"].cpu()
    off = 0
    for d, c in zip(query_features, counts):
        d["instances"].set("top10_shots", votes[off:off + c])
        d["instances"].set("top10_idx", idx[off:off + c])
        off += c
    return query_features


def get_nn_class_confirmatory(query_features: List[dict], k: int):
    """tools/run_nearest_neighbours.py:214-227: keep = (mode of the first k votes == detector class).
    torch.mode returns the smallest most-frequent value; the same rule is applied here on the stored votes."""
    for d in query_features:
        inst = d["instances"]
        votes = inst.get("top10_shots")[:, :k]
        dt = inst.gt_classes
        keep = torch.zeros(len(inst), dtype=torch.int64)
        if len(inst):
            # mode with smallest-value tie-break, vectorised: count matches per position, pick max count then min value
            eq = (votes[:, :, None] == votes[:, None, :]).sum(-1)            # [Q, k] multiplicity of each vote
            best = eq.max(dim=1, keepdim=True).values
            cand = torch.where(eq == best, votes, torch.full_like(votes, torch.iinfo(torch.int64).max))
            nn_class = cand.min(dim=1).values
            keep = (nn_class == dt.to(nn_class.device)).long()
        inst.set("keep", keep)
