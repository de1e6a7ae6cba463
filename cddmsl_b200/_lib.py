"""ctypes binding of the C ABI declared in include/cddmsl_b200.h.

There is deliberately no fallback: if `cddmsl_b200/lib/libcddmsl_b200.so` is missing, or a tensor is not a
CUDA tensor, the call fails loudly.  Nothing here (or anywhere in the package) imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcddmsl_b200.so")
_lib: Optional[ctypes.CDLL] = None

_vp, _i, _f, _d, _sz, _i64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t,
                              ctypes.c_int64)

# name -> (restype, argtypes); must list every symbol of include/cddmsl_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "cddmsl_abi_version": (_i, []),
    "cddmsl_error_string": (ctypes.c_char_p, [_i]),
    "cddmsl_launch_count": (ctypes.c_uint64, []),
    "cddmsl_roi_align_fwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "cddmsl_roi_align_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "cddmsl_roi_align_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "cddmsl_roi_align_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "cddmsl_roi_align_fwd2": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "cddmsl_roi_align_bwd2": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "cddmsl_nms_workspace_bytes": (_sz, [_i64]),
    "cddmsl_nms": (_i, [_vp, _vp, _vp, _i64, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_nms_topk": (_i, [_vp, _vp, _vp, _i64, _d, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_nms_batched_workspace_bytes": (_sz, [_i, _i64]),
    "cddmsl_nms_batched": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_nms_batched_topk": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _d, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_rpn_decode_topk": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _f, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp,
                                    _vp]),
    "cddmsl_clip_head_workspace_bytes": (_sz, [_i, _i, _i]),
    "cddmsl_clip_head_scores": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "cddmsl_clip_head_scores_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "cddmsl_clip_head_loss": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _f, _f, _vp, _i, _vp, _vp, _vp, _vp,
                                   _vp, _sz, _vp]),
    "cddmsl_box_reg_loss_workspace_bytes": (_sz, [_i]),
    "cddmsl_box_reg_loss": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_match_boxes_workspace_bytes": (_sz, [_i, _i]),
    "cddmsl_match_boxes": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_align_pack_normalized": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "cddmsl_align_loss_workspace_bytes": (_sz, [_i, _i, _i]),
    "cddmsl_align_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_contrastive_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_softmax_target_loss_workspace_bytes": (_sz, [_i]),
    "cddmsl_softmax_target_loss": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cddmsl_kd_l1_loss_workspace_bytes": (_sz, []),
    "cddmsl_kd_l1_loss": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
}
_INTERNAL = {"cddmsl_tune": (_i, [ctypes.c_char_p, _i])}


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library is not built (run `python -m cddmsl_b200.build`). "
                "cddmsl_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in {**SIGNATURES, **_INTERNAL}.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().cddmsl_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"cddmsl_b200: `{name}` must be a CUDA tensor (got {t.device}); there is no CPU path")


def launch_count() -> int:
    return int(lib().cddmsl_launch_count())


def tune(key: str, value: int) -> bool:
    return bool(lib().cddmsl_tune(key.encode(), int(value)))
