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
    assert lib.lvcb200_abi_version() == 2
    # argument validation happens on the host before any CUDA call
    rc = lib.lvcb200_gemm_bf16(None, None)
    assert rc == -1 and b"NULL" in lib.lvcb200_last_error()
    assert lib.lvcb200_batched_nms_workspace(1000) > 1000 * 20
    # mean | -c_s (padded) | bhat | |mu|^2 | the normalised bank as a bf16 hi/lo pair (rows padded to the 160-row chunk of the v2 kernel)
    assert lib.lvcb200_knn_prepared_bytes(600, 1024) == 4 * (1024 + 640 + 600 * 1024) + 256 + 2 * 640 * 1024 * 2
    assert lib.lvcb200_knn_tc_workspace(200_000, 600, 1024) >= 2 * 200_064 * 1024 * 2      # the queries as a bf16 hi/lo pair


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


def test_gemm_chain_host_validation():
    """lvcb200_gemm_chain_* validate the layer list on the host before any CUDA call: unsupported layers and in-chain
    overwrites are refused (workspace size 0 / error code), a well-formed chain gets a workspace size."""
    lib = _lib.load()

    def desc(A, W, D, N=64, K=64, M=256, res=0):
        d = _lib.GemmDesc()
        d.a_dtype, d.d_dtype = _lib.BF16, _lib.BF16
        d.A, d.lda, d.M_rows, d.W, d.ldw, d.D, d.ldd = A, K, M, W, K, D, N
        d.residual, d.ldr = (res, N) if res else (None, 0)
        d.M, d.N, d.K, d.taps = M, N, K, 1
        return d

    ok = (_lib.GemmDesc * 2)(desc(0x1000, 0x2000, 0x3000), desc(0x3000, 0x2000, 0x4000))
    assert lib.lvcb200_gemm_chain_workspace(ok, 2) >= 2 * 640 + 2 * 2 * 4
    bad_n = (_lib.GemmDesc * 1)(desc(0x1000, 0x2000, 0x3000, N=48))
    assert lib.lvcb200_gemm_chain_workspace(bad_n, 1) == 0 and b"multiples of 64" in lib.lvcb200_last_error()
    overwrite = (_lib.GemmDesc * 2)(desc(0x1000, 0x2000, 0x3000), desc(0x3000, 0x2000, 0x1000))
    assert lib.lvcb200_gemm_chain_workspace(overwrite, 2) == 0 and b"overwrite" in lib.lvcb200_last_error()
    f32 = desc(0x1000, 0x2000, 0x3000)
    f32.d_dtype = _lib.F32
    assert lib.lvcb200_gemm_chain_workspace((_lib.GemmDesc * 1)(f32), 1) == 0
    plan = _lib.ChainPlan()
    assert lib.lvcb200_gemm_chain_run(ctypes.byref(plan), None) == -1      # empty plan refused


def test_match_boxes_host_validation():
    """Matcher arguments are checked on the host (matcher.py:46-57 asserts) before any CUDA call."""
    lib = _lib.load()
    thr = (ctypes.c_float * 2)(0.3, 0.7)
    lab = (ctypes.c_int8 * 3)(0, -1, 1)
    assert lib.lvcb200_match_boxes(None, 0, None, None, 0, thr, 2, lab, 1, None, None, None, None, 0, None) == 0      # P == 0: nothing to do
    bad_thr = (ctypes.c_float * 2)(0.7, 0.3)
    assert lib.lvcb200_match_boxes(None, 0, None, None, 5, bad_thr, 2, lab, 0, None, None, None, None, 0, None) == -1
    assert b"sorted" in lib.lvcb200_last_error()
    bad_lab = (ctypes.c_int8 * 3)(0, 2, 1)
    assert lib.lvcb200_match_boxes(None, 0, None, None, 5, thr, 2, bad_lab, 0, None, None, None, None, 0, None) == -1
    assert b"labels" in lib.lvcb200_last_error()
    assert lib.lvcb200_match_boxes_workspace(20) >= 80
