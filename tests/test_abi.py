"""CPU-side checks of the drop-in boundary: the C-ABI library is built, loads, and exports every symbol that
include/lvcb200.h declares; host-side argument validation works without a GPU; ops refuse CPU tensors loudly."""
import ctypes
import os
import re

import pytest
import torch

from lvc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lvcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lvcb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lvcb200.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "python binding and header disagree"


def test_abi_version_and_error_channel():
    lib = _lib.load()
    assert lib.lvcb200_abi_version() == 1
    # argument validation happens on the host before any CUDA call
    rc = lib.lvcb200_gemm_bf16(None, None)
    assert rc == -1 and b"NULL" in lib.lvcb200_last_error()
    assert lib.lvcb200_batched_nms_workspace(1000) > 1000 * 20
    assert lib.lvcb200_knn_prepared_bytes(600, 1024) == 4 * (1024 + 640 + 600 * 1024)   # mean | -c_s (padded) | bhat


def test_no_cpu_fallback():
    from lvc_b200.layers import batched_nms, nms, roi_align
    b = torch.tensor([[0.0, 0, 1, 1]])
    s = torch.tensor([1.0])
    with pytest.raises(_lib.LvcB200Error):
        nms(b, s, 0.5)
    with pytest.raises(_lib.LvcB200Error):
        batched_nms(b, s, torch.zeros(1, dtype=torch.int64), 0.5)
    with pytest.raises(_lib.LvcB200Error):
        roi_align(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5), 2, 1.0, 0, True)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under lvc_b200/ may import it."""
    pkg = os.path.join(ROOT, "lvc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
