"""ctypes binding of libcwt_b200.so (the C ABI declared in include/cwt_b200.h).

There is no CPU fallback: if the shared object is missing or a call is made without a
CUDA device, this module raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CWT_LIB_PATH: developer hook (tools/build_variants.py) to time an alternative build of the same C ABI
LIB_PATH = os.environ.get("CWT_LIB_PATH") or os.path.join(_HERE, "lib", "libcwt_b200.so")

OK = 0
LABEL_U8, LABEL_I64 = 0, 1
FIT_AUTO, FIT_STREAM, FIT_RESIDENT, FIT_L2 = 0, 1, 2, 3
COSCLS_R, COSCLS_N, COSCLS_T = 1, 2, 4          # cwt_fit_coscls_f32 flags (CosCls cls_type 'r', 'n', 't')
ATTN_REASSOC, ATTN_TCGEN05 = 0, 1

_lib: Optional[C.CDLL] = None

_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/cwt_b200.h
SIGNATURES = {
    "cwt_version": (_i, []),
    "cwt_last_error": (C.c_char_p, []),
    "cwt_launch_count": (_ll, []),
    "cwt_prep_labels": (_i, [_vp, _i, _i, _ll, _i, _vp, _vp, _vp]),
    "cwt_fit_workspace_bytes": (_sz, [_i] * 7),
    "cwt_fit_classifier_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp] + [_i] * 8 + [_f, _i, _i, _vp, _sz, _vp]),
    "cwt_fit_status": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "cwt_fit_bias_workspace_bytes": (_sz, [_i] * 7),
    "cwt_fit_classifier_bias_f32": (_i, [_vp, _vp, _i] + [_vp] * 7 + [_i] * 8 + [_f, _f, _i, _vp, _sz, _vp]),
    "cwt_fit_coscls_workspace_bytes": (_sz, [_i] * 7),
    "cwt_fit_coscls_f32": (_i, [_vp, _vp, _i] + [_vp] * 7 + [_i] * 9 + [_f, _i, _vp, _sz, _vp]),
    "cwt_fit_multiclass_workspace_bytes": (_sz, [_i] * 7),
    "cwt_fit_multiclass_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp] + [_i] * 8 + [_f, _i, _vp, _sz, _vp]),
    "cwt_fit_dice_workspace_bytes": (_sz, [_i] * 7),
    "cwt_fit_classifier_dice_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp] + [_i] * 8 + [_f, _i, _vp, _sz, _vp]),
    "cwt_transformer_workspace_bytes": (_sz, [_i] * 6),
    "cwt_transformer_saved_bytes": (_sz, [_i] * 5),
    "cwt_transformer_fwd_f32": (_i, [_vp, _vp, _i] + [_vp] * 7 + [_f, _f, _vp, _vp] + [_i] * 6 + [_vp, _sz, _vp]),
    "cwt_transformer_bwd_f32": (_i, [_vp, _vp, _vp, _i] + [_vp] * 5 + [_f, _f, _vp] + [_vp] * 5 + [_i] * 5 + [_vp, _sz, _vp]),
    "cwt_logits_iou_workspace_bytes": (_sz, [_i] * 7),
    "cwt_logits_iou": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp] + [_i] * 8 + [_vp, _sz, _vp]),
    "cwt_upsample_argmax_iou": (_i, [_vp, _vp, _i, _vp, _vp] + [_i] * 6 + [_vp]),
    "cwt_intersection_union": (_i, [_vp, _vp, _i, _vp, _i, _ll, _i, _i, _vp]),
    "cwt_query_loss_workspace_bytes": (_sz, [_i] * 5),
    "cwt_query_loss_grad": (_i, [_vp, _vp, _i, _vp, _vp] + [_i] * 6 + [_vp, _sz, _vp]),
    "cwt_skinny_workspace_bytes": (_sz, [_i] * 4),
    "cwt_rows_times_feat": (_i, [_vp, _vp, _i, _vp] + [_i] * 4 + [_vp, _sz, _vp]),
    "cwt_feat_times_rows": (_i, [_vp, _vp, _i, _vp] + [_i] * 4 + [_vp, _sz, _vp]),
    "cwt_resize_pad_normalize_f32": (_i, [_vp, _vp] + [_i] * 5 + [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "cwt_normalize_features_f32": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "cwt_expand_zero_compressed_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, C.c_uint, _vp]),
}


# developer-only entry points declared in include/cwt_b200_debug.h (tools/ only; never on the product path)
DEBUG_SIGNATURES = {
    "cwt_debug_fit_classifier_prof_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp] + [_i] * 7 + [_f, _i, _vp, _sz, _vp, _vp]),
    "cwt_debug_l2_read": (_i, [_vp, _sz, _i, _i, _vp, _vp]),
}
FIT_BAD_LABEL, FIT_NO_FG, FIT_NONFINITE = 1, 2, 4      # bits of cwt_fit_status


def load() -> C.CDLL:
    """Load the shared object (built in-tree by ``python -m few_shot_seg_cwt_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m few_shot_seg_cwt_b200.build`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(DEBUG_SIGNATURES.items()):
        fn = getattr(lib, name)        # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().cwt_last_error().decode()


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = last_error()
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("few_shot_seg_cwt_b200 ops run on CUDA tensors only (no CPU fallback); "
                               f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("no CUDA tensor given")
    return dev


def label_kind(t: torch.Tensor) -> int:
    if t.dtype == torch.uint8:
        return LABEL_U8
    if t.dtype == torch.int64:
        return LABEL_I64
    raise TypeError(f"labels must be uint8 or int64, got {t.dtype}")


def launch_count() -> int:
    return int(load().cwt_launch_count())


class Workspace:
    """Grow-only per-device scratch buffer handed to the library (the caller owns all memory)."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes: int, device: torch.device, tag: str = "ws") -> torch.Tensor:
        key = (device.index if device.index is not None else torch.cuda.current_device(), tag,
               torch.cuda.current_stream(device).cuda_stream)
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


WORKSPACE = Workspace()
