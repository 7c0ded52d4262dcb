"""Mirror of detectron2/modeling/sampling.py (``subsample_labels``; lvc/modeling/sampling.py holds the same function) and of
``RPN._subsample_labels`` (proposal_generator/rpn.py:249-266) on liblvcb200's ``lvcb200_subsample_labels``.

The reference samples with two ``torch.randperm`` draws whose lengths are data dependent (two host synchronisations per label
vector).  Here the randomness is one uint32 key per element -- drawn from torch's generator unless the caller passes ``keys`` --
and a class's sample is its elements with the smallest keys: the same distribution (a uniformly random subset in uniformly random
order), decided on the device for a whole batch of label vectors at once.  Forward only (the function has no gradient).
"""
from typing import Optional, Tuple

import torch

from .. import _lib


def _keys(labels: torch.Tensor, keys: Optional[torch.Tensor], generator=None) -> torch.Tensor:
    if keys is None:
        keys = torch.randint(0, 2 ** 32, labels.shape, dtype=torch.int64, device=labels.device, generator=generator)
    if keys.shape != labels.shape:
        raise ValueError("subsample_labels: keys must have the shape of labels")
    return keys.to(torch.int64).to(torch.uint32) if keys.dtype != torch.uint32 else keys


def _run(labels, keys, num_samples, positive_fraction, bg_label, want_labels):
    _lib.require_cuda(labels)
    if labels.dtype not in (torch.int64, torch.int8):
        raise TypeError("subsample_labels: labels must be int64 or int8")
    N = labels.shape[-1]
    V = 1
    for d in labels.shape[:-1]:
        V *= d
    lab = labels.contiguous().view(V, N)
    k = _keys(labels, keys).contiguous().view(V, N)
    dev = lab.device
    pos = torch.empty((V, num_samples), dtype=torch.int64, device=dev)
    neg = torch.empty((V, num_samples), dtype=torch.int64, device=dev)
    counts = torch.empty((V, 2), dtype=torch.int32, device=dev)
    out = torch.empty((V, N), dtype=torch.int8, device=dev) if want_labels else None
    rc = _lib.load().lvcb200_subsample_labels(_lib.ptr(lab), int(lab.dtype == torch.int8), _lib.ptr(k), V, N, int(num_samples),
                                              float(positive_fraction), int(bg_label), _lib.ptr(pos), _lib.ptr(neg), _lib.ptr(counts),
                                              _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "lvcb200_subsample_labels")
    return pos, neg, counts, out


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     keys: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """sampling.py:9-54.  ``labels`` [N] (CUDA, int64 / int8): returns (pos_idx, neg_idx), 1-D int64 index vectors whose total length is
    ``num_samples`` or fewer (like the reference, the lengths are read back: one synchronisation)."""
    if labels.dim() != 1:
        raise ValueError("subsample_labels: labels must be 1-D (use subsample_labels_batched for [V, N])")
    pos, neg, counts, _ = _run(labels, keys, num_samples, positive_fraction, bg_label, False)
    n_pos, n_neg = counts[0].tolist()
    return pos[0, :n_pos], neg[0, :n_neg]


def subsample_labels_batched(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                             keys: Optional[torch.Tensor] = None):
    """[V, N] label vectors in one launch, nothing read back: (pos_idx [V, num_samples], neg_idx [V, num_samples] padded with -1,
    counts [V, 2] int32)."""
    pos, neg, counts, _ = _run(labels, keys, num_samples, positive_fraction, bg_label, False)
    return pos, neg, counts


def subsample_rpn_labels(label: torch.Tensor, batch_size_per_image: int = 256, positive_fraction: float = 0.5,
                         keys: Optional[torch.Tensor] = None) -> torch.Tensor:
    """RPN._subsample_labels (rpn.py:249-266) for one ([A]) or a batch ([N, A]) of anchor label vectors of -1 / 0 / 1: everything
    outside the sample becomes -1.  Returns a new int8 tensor of the input's shape (the reference overwrites its argument)."""
    _, _, _, out = _run(label, keys, batch_size_per_image, positive_fraction, 0, True)
    return out.view(label.shape)
