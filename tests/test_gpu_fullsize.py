"""BASELINE.json's full-size configurations that the scalar oracle cannot cover in seconds, checked through size-independent properties
(-m gpu): config #4 (kNN: 200 000 x 1024 queries against a 600 x 1024 bank) and config #5 (box corrector head: 100 000 pooled RoIs).
Configs #1 / #2 (whole detector, batch 8 of 800 x 1333) are covered by the reference-generated end-to-end fixtures in test_gpu_engine.py."""
import numpy as np
import pytest
import torch

from lvc_b200 import ops
from lvc_b200.config import DetectorConfig
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _knn_case(seed=0, Q=200_000, S=600, D=1024, ncls=20):
    g = torch.Generator(device=DEV).manual_seed(seed)
    cls = torch.arange(ncls, device=DEV).repeat_interleave(S // ncls)
    means = torch.randn(ncls, D, generator=g, device=DEV) * 0.08
    bank = torch.randn(S, D, generator=g, device=DEV) + means[cls] + 3.0             # common offset: the centring matters
    qc = torch.randint(0, ncls, (Q,), generator=g, device=DEV)
    q = torch.randn(Q, D, generator=g, device=DEV) + means[qc] + 3.0
    return bank, cls, q, qc


def test_knn_config4_full_size_properties():
    """200 000 queries (BASELINE config #4) through the tensor-core path:
    * the exact SIMT path (a second, independent kernel: fp32 FMA scores, no candidate logic) returns the same top-10 indices for all but
      near-tied positions, and the same keep flags outside those rows;
    * permutation equivariance: verify(q[perm]) == verify(q)[perm] bit for bit (a query's result may not depend on its tile or position);
    * a 2 000-query random sample agrees with the C oracle (indices, except positions whose similarities tie to 1e-5);
    * self retrieval: a query that IS a bank row (plus noise far below the neighbour gap) returns that row first and keeps iff classes match."""
    bank, cls, q, qc = _knn_case()
    Q = q.shape[0]
    kb = ops.KnnBank(bank, cls)
    tc = kb.verify(q, qc, return_sim=True, path="tc3")
    ex = kb.verify(q, qc, return_sim=True, path="simt")
    bad = tc["top_idx"] != ex["top_idx"]
    assert float(bad.float().mean()) < 1e-3
    assert bool(((tc["top_sim"] - ex["top_sim"]).abs()[bad] < 1e-5).all())          # only near-ties may swap
    rows_ok = ~bad.any(1)
    assert torch.equal(tc["keep"][rows_ok], ex["keep"][rows_ok]) and torch.equal(tc["votes"][rows_ok], ex["votes"][rows_ok])
    perm = torch.randperm(Q, generator=torch.Generator(device=DEV).manual_seed(1), device=DEV)
    tp = kb.verify(q[perm].contiguous(), qc[perm].contiguous(), return_sim=True, path="tc3")
    assert torch.equal(tp["top_idx"], tc["top_idx"][perm]) and torch.equal(tp["keep"], tc["keep"][perm])
    assert torch.equal(tp["top_sim"], tc["top_sim"][perm])
    samp = torch.randperm(Q, generator=torch.Generator(device=DEV).manual_seed(2), device=DEV)[:2000]
    want = O.knn_verify(bank.cpu().numpy(), cls.cpu().numpy(), q[samp].cpu().numpy(), qc[samp].cpu().numpy())
    gi, gs = tc["top_idx"][samp].cpu().numpy(), tc["top_sim"][samp].cpu().numpy()
    b2 = gi != want["top_idx"]
    assert b2.mean() < 1e-3 and np.all(np.abs(gs[b2] - want["top_sim"][b2]) < 1e-5)
    ok = ~b2.any(1)
    assert np.array_equal(tc["keep"][samp].cpu().numpy()[ok], want["keep"][ok])
    # self retrieval at full size
    src = torch.randint(0, bank.shape[0], (Q,), generator=torch.Generator(device=DEV).manual_seed(3), device=DEV)
    q_self = bank[src] + 1e-4 * torch.randn(Q, bank.shape[1], generator=torch.Generator(device=DEV).manual_seed(4), device=DEV)
    qc_self = torch.where(torch.arange(Q, device=DEV) % 2 == 0, cls[src], (cls[src] + 1) % 20)
    rs = kb.verify(q_self, qc_self, path="tc3")
    assert torch.equal(rs["top_idx"][:, 0], src)
    del q, q_self
    torch.cuda.empty_cache()


def test_box_corrector_config5_full_size_properties():
    """100 000 pooled RoIs x 12 544 features (BASELINE config #5: 2.5 GB of bf16 input) through one stage of the corrector head (fc 12544 ->
    1024 -> 1024 -> 1024 + Linear 1024 -> 4):
    * row independence: the deltas of any subset of rows computed alone are bit-identical to the same rows of the full run (a row's result
      may not depend on which tile it falls into);
    * a 256-row sample agrees with an fp64 evaluation of the same bf16 weights with bf16-rounded activations between layers (2e-2 relative);
    * all-zero rows give exactly the bias path."""
    from lvc_b200.modeling import BoxCorrectorHead
    from lvc_b200.weights import synthetic_corrector_head
    cfg = DetectorConfig(depth=50, num_fc=3)
    sd = synthetic_corrector_head(cfg, seed=5)
    head = BoxCorrectorHead(cfg, sd)
    R = 100_000
    g = torch.Generator(device=DEV).manual_seed(6)
    pooled = torch.empty((R, 12544), dtype=torch.bfloat16, device=DEV)
    for i in range(0, R, 20_000):      # generate in slices: an fp32 temporary of the whole matrix would be 5 GB
        pooled[i:i + 20_000] = torch.relu(torch.randn(20_000, 12544, generator=g, device=DEV)).bfloat16()
    pooled[7] = 0
    full = head.head(0, pooled)                                              # [R, 16] fp32, 4 used
    assert full.shape == (R, 16) and bool(torch.isfinite(full).all())
    idx = torch.randperm(R, generator=torch.Generator(device=DEV).manual_seed(7), device=DEV)[:3000].sort().values
    sub = head.head(0, pooled[idx].contiguous())
    assert torch.equal(sub, full[idx])
    # fp64 evaluation of a sample with the head's own bf16 weights
    s = idx[:256]
    x = pooled[s].double()
    for w, b in head.fc[0]:
        x = torch.relu(x @ w.double().t() + b.double()).bfloat16().double()   # activations are stored as bf16 between the layers
    wp, bp = head.pred[0]
    want = x @ wp.double().t() + bp.double()
    rel = float((full[s].double() - want).norm() / want.norm())
    assert rel < 2e-2, rel
    zero = torch.zeros((1, 12544), dtype=torch.bfloat16, device=DEV)
    assert torch.equal(head.head(0, zero)[0], full[7])
    del pooled
    torch.cuda.empty_cache()
