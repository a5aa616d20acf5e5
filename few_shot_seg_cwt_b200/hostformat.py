"""Zero-compressed host format of an episode batch (transport only: what crosses PCIe in front of the head).

Post-ReLU features are about half zeros; an evaluation sweep from host memory is bound by the H2D copy (15.2 MB per 1-shot
episode; the 8 GPUs of a box share ~180 GB/s of host bandwidth). ``compress_batch`` keeps, per feature tensor, a bit mask,
the per-word prefix counts and the packed non-zero values (lossless on the bit pattern: -0.0 and NaN survive);
``HostPipeline`` copies those and rebuilds the dense tensors on the device with ``cwt_expand_zero_compressed_f32``.
The reference ships dense tensors (``.cuda()``, src/test.py:153-157); this is the B200-side replacement of that leg."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import _lib as L
from .synthetic import EpisodeBatch


@dataclass
class CompressedMap:
    """One fp32 tensor [E, ...] compressed per episode (so that sub-batches are slices)."""
    shape: tuple                 # dense shape
    mask: torch.Tensor           # uint32 as int32 [E, W]   W = words per episode
    woff: torch.Tensor           # uint32 as int32 [E, ceil(W / 32)]   set bits before each block of 32 words, over the whole batch
    vals: torch.Tensor           # float32 [nnz]
    val_start: List[int]         # [E + 1] start of every episode's values

    def nbytes(self, lo: int = 0, hi: int = None) -> int:
        hi = self.mask.shape[0] if hi is None else hi
        return (hi - lo) * (self.mask.shape[1] + self.woff.shape[1]) * 4 + (self.val_start[hi] - self.val_start[lo]) * 4

    def pin_memory(self) -> "CompressedMap":
        return CompressedMap(self.shape, self.mask.pin_memory(), self.woff.pin_memory(), self.vals.pin_memory(), self.val_start)


def compress_map(t: torch.Tensor) -> CompressedMap:
    """t: float32 CPU tensor [E, ...] with a multiple of 32 elements per episode."""
    if t.dtype != torch.float32 or t.is_cuda:
        raise TypeError("compress_map takes a float32 CPU tensor")
    E = t.shape[0]
    per = t[0].numel() if E else 0
    if per % 32:
        raise ValueError("elements per episode must be a multiple of 32")
    bits = t.contiguous().view(torch.int32).reshape(-1).numpy() != 0            # the BIT PATTERN decides (keeps -0.0, NaN)
    vals = torch.from_numpy(t.contiguous().reshape(-1).numpy()[bits].copy())
    mask = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(E, per // 32)
    W = per // 32
    cnt = bits.reshape(-1, 32).sum(axis=1, dtype=np.int64).reshape(E, W)                # set bits per word
    woff = (np.cumsum(cnt.reshape(-1)) - cnt.reshape(-1)).reshape(E, W)[:, ::32]         # ... before each block of 32 words of a row
    total = int(cnt.sum())
    if total >= 2 ** 32:
        raise ValueError("batch too large for 32-bit value offsets: compress fewer episodes per batch")
    starts = [0] + [int(x) for x in np.cumsum(cnt.sum(axis=1))] if E else [0]
    return CompressedMap(tuple(t.shape), torch.from_numpy(mask.view(np.int32).copy()),
                         torch.from_numpy(np.ascontiguousarray(woff).astype(np.uint32).view(np.int32).copy()), vals, starts)


def expand_map_reference(c: CompressedMap) -> torch.Tensor:
    """Host-side inverse (checker for the tests; the product expands on the device)."""
    bits = np.unpackbits(c.mask.numpy().view(np.uint8).reshape(-1), bitorder="little").astype(bool)
    out = np.zeros(bits.size, dtype=np.float32)
    out[bits] = c.vals.numpy()
    return torch.from_numpy(out).reshape(c.shape)


def expand_map(mask: torch.Tensor, woff: torch.Tensor, vals: torch.Tensor, out: torch.Tensor, base: int) -> torch.Tensor:
    """Device expansion of ``mask`` [n, W] / ``woff`` [n, ceil(W/32)] (device int32) + ``vals`` (device fp32) into ``out``."""
    dev = L.require_cuda(mask, woff, out)
    n_rows, W = mask.shape
    if out.numel() != 32 * n_rows * W or tuple(woff.shape) != (n_rows, (W + 31) // 32):
        raise ValueError("out must hold 32 elements per mask word and woff one entry per block of 32 words of a row")
    with torch.cuda.device(dev):
        rc = L.load().cwt_expand_zero_compressed_f32(L.ptr(mask), L.ptr(woff), L.ptr(vals), L.ptr(out), n_rows, W,
                                                     int(base) & 0xFFFFFFFF, L.stream_ptr(dev))
    L.check(rc, "cwt_expand_zero_compressed_f32")
    return out


@dataclass
class CompressedEpisodeBatch:
    """An :class:`EpisodeBatch` whose two feature tensors travel zero-compressed (labels, weights, indices stay dense)."""
    f_s: CompressedMap
    s_label: torch.Tensor
    f_q: CompressedMap
    q_label: torch.Tensor
    w0: torch.Tensor
    subcls: torch.Tensor
    idx: torch.Tensor

    @property
    def n_episodes(self) -> int:
        return self.s_label.shape[0]

    def pin_memory(self) -> "CompressedEpisodeBatch":
        pm = lambda t: t.pin_memory()
        return CompressedEpisodeBatch(self.f_s.pin_memory(), pm(self.s_label), self.f_q.pin_memory(), pm(self.q_label),
                                      pm(self.w0), pm(self.subcls), pm(self.idx))

    def nbytes(self, lo: int = 0, hi: int = None) -> int:
        hi = self.n_episodes if hi is None else hi
        dense = sum(t[lo:hi].numel() * t.element_size() for t in (self.s_label, self.q_label, self.w0))
        return dense + self.f_s.nbytes(lo, hi) + self.f_q.nbytes(lo, hi)

    def dense_nbytes(self) -> int:
        n = lambda shape: int(np.prod(shape)) * 4
        return n(self.f_s.shape) + n(self.f_q.shape) + sum(t.numel() * t.element_size() for t in (self.s_label, self.q_label, self.w0))


def compress_batch(b: EpisodeBatch) -> CompressedEpisodeBatch:
    return CompressedEpisodeBatch(compress_map(b.f_s), b.s_label, compress_map(b.f_q), b.q_label, b.w0, b.subcls, b.idx)
