"""The reference's validation transform (Resize -> ToTensor -> Normalize, src/dataset/dataset.py:78-84) as one fused
CUDA kernel per image. Drop-in for ``transform.Compose([transform.Resize(image_size, padding), transform.ToTensor(),
transform.Normalize(mean, std)])`` (src/dataset/transform.py:109-163, 58-82, 85-107)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def find_new_hw(ori_h: int, ori_w: int, test_size: int) -> Tuple[int, int]:
    """``Resize.find_new_hw`` (src/dataset/transform.py:117-137): the larger side becomes ``test_size``, the other keeps
    the aspect ratio (truncated), both are then rounded DOWN to a multiple of 8."""
    if ori_h >= ori_w:
        ratio = test_size * 1.0 / ori_h
        new_h, new_w = test_size, int(ori_w * ratio)
    else:
        ratio = test_size * 1.0 / ori_w
        new_h, new_w = int(ori_h * ratio), test_size
    if new_h % 8 != 0:
        new_h = (int(new_h / 8)) * 8
    if new_w % 8 != 0:
        new_w = (int(new_w / 8)) * 8
    return new_h, new_w


def resize_pad_normalize(image: torch.Tensor, label: Optional[torch.Tensor], size: int, mean: Sequence[float],
                         std: Sequence[float], padding: Optional[Sequence[float]] = None, pad_label: int = 255,
                         label_dtype: torch.dtype = torch.int64):
    """image [h,w,3] float32 RGB in [0,255] (CUDA), label [h,w] uint8 (CUDA) or None ->
    (image [3,size,size] float32 normalised, label [size,size] ``label_dtype``) — or (image, new_h, new_w) without a
    label, like the reference's ``Resize`` (transform.py:162-163)."""
    dev = L.require_cuda(image) if label is None else L.require_cuda(image, label)
    if image.dim() != 3 or image.shape[2] != 3 or image.dtype != torch.float32:
        raise ValueError(f"image must be [h,w,3] float32, got {tuple(image.shape)} {image.dtype}")
    image = image.contiguous()
    oh, ow = int(image.shape[0]), int(image.shape[1])
    if label is not None:
        if tuple(label.shape) != (oh, ow) or label.dtype != torch.uint8:
            raise ValueError("label must be [h,w] uint8 with the image's size")
        label = label.contiguous()
    nh, nw = find_new_hw(oh, ow, size)
    out = torch.empty(3, size, size, dtype=torch.float32, device=dev)
    lab_out = torch.empty(size, size, dtype=label_dtype, device=dev) if label is not None else None
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    m, s_, p_ = f3(mean), f3(std), (f3(padding) if padding is not None else None)
    with torch.cuda.device(dev):
        rc = L.load().cwt_resize_pad_normalize_f32(L.ptr(image), L.ptr(label), oh, ow, nh, nw, int(size),
                                                   C.cast(m, C.c_void_p), C.cast(s_, C.c_void_p),
                                                   C.cast(p_, C.c_void_p) if p_ is not None else None, int(pad_label),
                                                   L.ptr(out), L.ptr(lab_out),
                                                   L.LABEL_I64 if label_dtype == torch.int64 else L.LABEL_U8, L.stream_ptr(dev))
    L.check(rc, "cwt_resize_pad_normalize_f32")
    return (out, lab_out) if label is not None else (out, nh, nw)


class ValTransform:
    """Callable with the signature of the reference's composed validation transform: ``(image, label) -> (image, label)``
    on numpy arrays as ``cv2.imread`` + ``np.float32`` deliver them (src/dataset/dataset.py:138-141); returns CUDA tensors
    ``[3,size,size]`` float32 and ``[size,size]`` int64."""

    def __init__(self, size: int, mean: Sequence[float], std: Sequence[float], padding: Optional[Sequence[float]] = None,
                 device="cuda"):
        self.size, self.mean, self.std, self.padding, self.device = size, list(mean), list(std), padding, torch.device(device)

    def __call__(self, image, label):
        img = torch.as_tensor(np.ascontiguousarray(image, dtype=np.float32)).to(self.device, non_blocking=True)
        lab = torch.as_tensor(np.ascontiguousarray(label, dtype=np.uint8)).to(self.device, non_blocking=True)
        return resize_pad_normalize(img, lab, self.size, self.mean, self.std, self.padding)
