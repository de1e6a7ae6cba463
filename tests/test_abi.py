"""CPU: the C-ABI library loads and exports exactly what include/cddmsl_b200.h declares; the product package
never touches oracle/ and has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cddmsl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cddmsl_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from cddmsl_b200 import _lib, build

    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/cddmsl_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "cddmsl_b200/_lib.py SIGNATURES out of sync with the header"
    assert _lib.lib().cddmsl_abi_version() == 1
    assert b"invalid" in _lib.lib().cddmsl_error_string(-1)


def test_sm100a_sass_is_present():
    import subprocess

    from cddmsl_b200 import _lib

    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cddmsl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_no_cpu_fallback():
    from cddmsl_b200.layers import ROIAlign, batched_nms

    with pytest.raises((RuntimeError, NotImplementedError)):
        ROIAlign((7, 7), 1.0, 0)(torch.zeros(1, 1, 8, 8), torch.tensor([[0.0, 1, 1, 5, 5]]))
    with pytest.raises((RuntimeError, NotImplementedError)):
        batched_nms(torch.tensor([[0.0, 0, 1, 1]]), torch.tensor([1.0]), torch.tensor([0]), 0.5)


def test_workspace_queries_do_not_need_a_gpu():
    from cddmsl_b200 import _lib

    L = _lib.lib()
    assert L.cddmsl_nms_workspace_bytes(12000) > 12000 * 188 * 8
    assert L.cddmsl_align_loss_workspace_bytes(8, 256, 256) >= 2048 * 2048 * 4
    assert L.cddmsl_roi_align_bwd_workspace_bytes(16, 1024, 38, 63, 8192) > 0
    # batched NMS: B images' masks; a single image through the batched entry point sizes for the segmented sort
    assert L.cddmsl_nms_batched_workspace_bytes(16, 12000) > 16 * 12000 * 188 * 8
    assert L.cddmsl_nms_batched_workspace_bytes(1, 12000) >= L.cddmsl_nms_workspace_bytes(12000)
    assert L.cddmsl_box_reg_loss_workspace_bytes(8192) >= 32 * 4
