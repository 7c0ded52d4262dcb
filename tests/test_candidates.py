"""a15: the candidate filter (get_ret_anns, tools/create_coco_dataset_from_dets_all.py:129-193).  The oracle restatement and the
product's vectorised mirror against the reference's own outputs (tests/golden/candidates.npz).  CPU; the mirror is pure torch
indexing, so the same code runs on CUDA tensors behind the detector."""
import numpy as np
import pytest
import torch

from lvc_b200.candidates import select_candidates
from oracle import oracle as O

CASES = [("score_full", dict(top=False, full=True, k_min=0.8, k_max=1.0, ar=0.0)),
         ("score_nofull", dict(top=False, full=False, k_min=0.5, k_max=0.9, ar=0.05)),
         ("top_full", dict(top=True, full=True, k_min=25, k_max=3, ar=0.0))]


def _train(g):
    return {int(c): set(int(v) for v in g[f"train_{int(c)}"]) for c in g["train_keys"]}


@pytest.mark.parametrize("tag,kw", CASES)
def test_oracle_matches_reference(golden, tag, kw):
    g = golden("candidates")
    flags = O.select_candidates(g["image_id"], g["category"], g["score"], g["area"], g["image_area"], _train(g),
                                [int(c) for c in g["novel"]], **kw)
    assert np.array_equal(flags, g["flags_" + tag])


@pytest.mark.parametrize("tag,kw", CASES)
def test_mirror_matches_reference(golden, tag, kw):
    g = golden("candidates")
    flags = select_candidates(torch.from_numpy(g["image_id"]), torch.from_numpy(g["category"]), torch.from_numpy(g["score"]),
                              torch.from_numpy(g["area"]), torch.from_numpy(g["image_area"]), _train(g),
                              [int(c) for c in g["novel"]], **kw)
    assert np.array_equal(flags.numpy(), g["flags_" + tag])
    assert (flags == 1).sum() > 10


def test_score_bounds_are_half_open():
    """K_min < score <= K_max (the reference's left searchsorted on -scores)."""
    sc = torch.tensor([0.5, 0.5, 0.75, 1.0, 0.25], dtype=torch.float32)          # exactly representable bounds
    f = select_candidates(torch.arange(5), torch.zeros(5, dtype=torch.int64), sc, torch.full((5,), 10.0), torch.full((5,), 100.0),
                          {}, [0], 0.5, 1.0, full=False)
    want = O.select_candidates(np.arange(5), np.zeros(5, np.int64), sc.numpy(), np.full(5, 10.0), np.full(5, 100.0), {}, [0], 0.5, 1.0,
                               full=False)
    assert f.tolist() == want.tolist() == [0, 0, 1, 1, 0]     # score == K_min is excluded, score == K_max is kept


def test_empty():
    e = torch.zeros(0)
    assert select_candidates(e.long(), e.long(), e, e, e, {}, [1], 0.8, 1.0).numel() == 0
