"""Deterministic synthetic few-shot episodes.

Real PASCAL-5i / COCO-20i images are not available offline, so the benchmark and
the parity suites use feature-level synthetic episodes whose *shapes, dtypes and
label semantics* follow what the reference data pipeline feeds the head:

* features  ``f_s [S,C,h,w]``, ``f_q [C,h,w]`` fp32, post-ReLU (>= 0, ~half zeros),
  what ``PSPNet.extract_features`` returns (reference ``src/model/pspnet.py:172-181``);
* labels    ``s_label [S,H,W]``, ``q_label [H,W]`` with values {0, 1, 255}: 0 background,
  1 the episode's novel class, 255 on the bottom/right padding band that the
  reference ``Resize`` transform creates (``src/dataset/transform.py:140-163``: aspect
  preserving resize to a multiple-of-8 side, label padded with 255) plus a thin
  255 ring on the object boundary (VOC "void" pixels);
* ``W0 [2,C]`` ~ U(-1/sqrt(C), 1/sqrt(C)) — what ``nn.Conv2d(C, 2, 1, bias=False)``
  draws (kaiming-uniform, a=sqrt(5); reference ``src/test.py:164``);
* ``subcls = idx % num_classes_val + 1`` (reference ``src/dataset/dataset.py:265``).

Every episode owns one ``torch.Generator`` seeded ``seed + idx`` (manual_seed 2021 in
``config_files/pascal.yaml``) so shards are order independent: rank r of G simply
generates episodes ``idx = r, r+G, ...``.

Two feature styles:

``unit``      ``relu(N(0,1) + 0.5 * proto * mask)`` — O(1) scale, learnable fg/bg.
``backbone``  the same pattern times 20 — mean ~8, max ~100, ~50 % zeros, i.e. the
              statistics of a random-init PSPNet-R50 bottleneck measured in the survey;
              with ``cls_lr=0.1`` this is the regime where the fit loss spikes before
              recovering.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

IGNORE = 255


@dataclass
class Episode:
    f_s: torch.Tensor       # [S,C,h,w] fp32
    s_label: torch.Tensor   # [S,H,W] uint8 or int64 in {0,1,255}
    f_q: torch.Tensor       # [C,h,w] fp32
    q_label: torch.Tensor   # [H,W]
    w0: torch.Tensor        # [2,C] fp32
    subcls: int
    idx: int


def _object_mask(g: torch.Generator, H: int, W: int, vh: int, vw: int) -> torch.Tensor:
    """Binary object mask [H,W] inside the valid (un-padded) window vh x vw:
    a random rectangle or ellipse covering roughly 3-40 % of the valid area."""
    frac = 0.03 + 0.37 * torch.rand((), generator=g).item()
    aspect = math.exp((torch.rand((), generator=g).item() - 0.5) * 1.2)
    area = frac * vh * vw
    bh = min(max(int(math.sqrt(area * aspect)), 8), vh - 2)
    bw = min(max(int(area / max(bh, 1)), 8), vw - 2)
    y0 = int(torch.randint(0, vh - bh, (), generator=g).item())
    x0 = int(torch.randint(0, vw - bw, (), generator=g).item())
    ellipse = torch.rand((), generator=g).item() < 0.5
    ys = torch.arange(H, dtype=torch.float32).view(H, 1)
    xs = torch.arange(W, dtype=torch.float32).view(1, W)
    if ellipse:
        cy, cx = y0 + bh / 2.0, x0 + bw / 2.0
        m = ((ys - cy) / (bh / 2.0)) ** 2 + ((xs - cx) / (bw / 2.0)) ** 2 <= 1.0
    else:
        m = (ys >= y0) & (ys < y0 + bh) & (xs >= x0) & (xs < x0 + bw)
    return m


def _label_and_mask(g: torch.Generator, H: int, W: int, h: int, w: int, label_dtype: torch.dtype):
    # valid window: one side full, the other a multiple of 8 >= 60 % (Resize semantics)
    if torch.rand((), generator=g).item() < 0.5:
        vh = H
        vw = max(int((0.6 + 0.4 * torch.rand((), generator=g).item()) * W) // 8 * 8, 16)
    else:
        vw = W
        vh = max(int((0.6 + 0.4 * torch.rand((), generator=g).item()) * H) // 8 * 8, 16)
    vh, vw = min(vh, H), min(vw, W)
    obj = _object_mask(g, H, W, vh, vw)
    # thin boundary ring of "void" pixels: 1-px dilation minus the object
    p = torch.nn.functional.max_pool2d(obj[None, None].float(), 3, 1, 1)[0, 0] > 0
    ring = p & ~obj
    lab = torch.zeros(H, W, dtype=torch.int64)
    lab[obj] = 1
    lab[ring] = IGNORE
    lab[vh:, :] = IGNORE
    lab[:, vw:] = IGNORE
    # low-resolution object mask sampled at the align_corners grid (stride (H-1)/(h-1))
    iy = torch.linspace(0, H - 1, h).round().long()
    ix = torch.linspace(0, W - 1, w).round().long()
    valid = torch.zeros(H, W, dtype=torch.bool)
    valid[:vh, :vw] = True
    m_lo = (obj & valid)[iy][:, ix].float()
    return lab.to(label_dtype), m_lo


def make_episode(idx: int, shot: int = 1, C: int = 512, h: int = 60, w: int = 60,
                 H: int = 473, W: int = 473, seed: int = 2021, style: str = "unit",
                 num_classes_val: int = 5, label_dtype: torch.dtype = torch.uint8) -> Episode:
    """Episode ``idx`` of the synthetic stream. CPU tensors, deterministic."""
    g = torch.Generator()
    g.manual_seed(seed + idx)
    scale = {"unit": 1.0, "backbone": 20.0}[style]
    proto = torch.randn(C, generator=g)
    w0 = (torch.rand(2, C, generator=g) * 2.0 - 1.0) / math.sqrt(C)

    s_labels, f_s = [], []
    for _ in range(shot):
        lab, m_lo = _label_and_mask(g, H, W, h, w, label_dtype)
        f = torch.relu(torch.randn(C, h, w, generator=g) + 0.5 * proto.view(C, 1, 1) * m_lo) * scale
        s_labels.append(lab)
        f_s.append(f)
    q_label, mq_lo = _label_and_mask(g, H, W, h, w, label_dtype)
    f_q = torch.relu(torch.randn(C, h, w, generator=g) + 0.5 * proto.view(C, 1, 1) * mq_lo) * scale
    return Episode(f_s=torch.stack(f_s), s_label=torch.stack(s_labels), f_q=f_q, q_label=q_label,
                   w0=w0, subcls=idx % num_classes_val + 1, idx=idx)


@dataclass
class EpisodeBatch:
    """``E`` episodes stacked on a leading axis (the layout the batched ops take)."""
    f_s: torch.Tensor       # [E,S,C,h,w]
    s_label: torch.Tensor   # [E,S,H,W]
    f_q: torch.Tensor       # [E,C,h,w]
    q_label: torch.Tensor   # [E,H,W]
    w0: torch.Tensor        # [E,2,C]
    subcls: torch.Tensor    # [E] int64
    idx: torch.Tensor       # [E] int64

    def to(self, device, non_blocking: bool = False) -> "EpisodeBatch":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return EpisodeBatch(mv(self.f_s), mv(self.s_label), mv(self.f_q), mv(self.q_label),
                            mv(self.w0), mv(self.subcls), mv(self.idx))

    def pin_memory(self) -> "EpisodeBatch":
        pm = lambda t: t.pin_memory()
        return EpisodeBatch(pm(self.f_s), pm(self.s_label), pm(self.f_q), pm(self.q_label),
                            pm(self.w0), pm(self.subcls), pm(self.idx))

    @property
    def n_episodes(self) -> int:
        return self.f_s.shape[0]

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in
                   (self.f_s, self.s_label, self.f_q, self.q_label, self.w0))


def make_batch(indices: List[int], **kw) -> EpisodeBatch:
    eps = [make_episode(i, **kw) for i in indices]
    return EpisodeBatch(
        f_s=torch.stack([e.f_s for e in eps]), s_label=torch.stack([e.s_label for e in eps]),
        f_q=torch.stack([e.f_q for e in eps]), q_label=torch.stack([e.q_label for e in eps]),
        w0=torch.stack([e.w0 for e in eps]),
        subcls=torch.tensor([e.subcls for e in eps], dtype=torch.int64),
        idx=torch.tensor([e.idx for e in eps], dtype=torch.int64))


def shard_indices(n_episodes: int, rank: int, world: int, start: int = 0) -> List[int]:
    """Episode ``i`` goes to rank ``i mod world``."""
    return list(range(start + rank, start + n_episodes, world))


def make_transformer_params(n_head: int, C: int = 512, seed: int = 2021,
                            perturb_ln: bool = True) -> dict:
    """Random-init parameters with the reference's initialisers and state-dict names
    (``src/model/transformer.py:44-51``): w_qkvs ~ N(0, sqrt(2/(C+C))), fc xavier-normal,
    fc.bias ~ U(+-1/sqrt(nH*C)), LayerNorm affine (perturbed from 1/0 so parity tests see it)."""
    g = torch.Generator()
    g.manual_seed(seed * 7919 + n_head)
    p = {
        "w_qkvs.weight": torch.randn(n_head * C, C, generator=g) * math.sqrt(2.0 / (C + C)),
        "fc.weight": torch.randn(C, n_head * C, generator=g) * math.sqrt(2.0 / (C + n_head * C)),
        "fc.bias": (torch.rand(C, generator=g) * 2 - 1) / math.sqrt(n_head * C),
        "layer_norm.weight": torch.ones(C),
        "layer_norm.bias": torch.zeros(C),
    }
    if perturb_ln:
        p["layer_norm.weight"] = 1.0 + 0.1 * torch.randn(C, generator=g)
        p["layer_norm.bias"] = 0.05 * torch.randn(C, generator=g)
    return p
