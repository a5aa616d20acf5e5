"""Support-classifier fit behind the reference's call surface.

The reference has no classifier module — the fit is an inline ``nn.Conv2d`` + ``optim.SGD`` +
``nn.CrossEntropyLoss`` loop (src/test.py:164-187, src/train.py:206-231) and the method form
``PSPNet.inner_loop(f_s, s_label)`` (src/model/pspnet.py:189-205) which mutates
``self.classifier.weight``. This module offers both forms on top of one fused CUDA fit.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

import torch.nn.functional as F

from . import _lib as L
from .ops import fit_classifier, fit_classifier_bias, fit_classifier_dice, fit_coscls, fit_multiclass, label_counts, normalize_features


def get_classifier(bottleneck_dim: int = 512, num_classes: int = 2, device=None) -> nn.Conv2d:
    """``nn.Conv2d(bottleneck_dim, num_classes, 1, bias=False)`` — what the reference builds per episode
    (src/test.py:164; get_classifier, src/model/pspnet.py:326-334 with cls_type 'o')."""
    return nn.Conv2d(bottleneck_dim, num_classes, kernel_size=1, bias=False).to(device)


def draw_initial_weights(n_episodes: int, C: int = 512, generator: Optional[torch.Generator] = None,
                         device=None) -> torch.Tensor:
    """W0 [E,2,C] ~ U(-1/sqrt(C), 1/sqrt(C)): the kaiming-uniform(a=sqrt(5)) init of a fresh
    nn.Conv2d(C, 2, 1). Drawn on the CPU generator like the reference, then moved."""
    w = (torch.rand(n_episodes, 2, C, generator=generator) * 2.0 - 1.0) / math.sqrt(C)
    return w.to(device) if device is not None else w


class CosCls(nn.Module):
    """Cosine classifier with the reference's constructor and parameter names (``CosCls``,
    src/model/pspnet.py:290-315): ``cls`` is the 1x1 conv, ``scale_factor`` the temperature (2.0).
    ``cls_type`` = four flags [weight-norm reparametrisation 'r', weight normalisation 'n', bias 'b', learnable
    temperature 't']; '0' / 'o' switch a flag off. ``forward`` is plain torch (it is not on the episodic hot path);
    :func:`inner_loop` fits the flag-free form ('oooo' / '0000') with the fused CUDA fit, the form with a bias ('oobo') and
    every other flag combination ('r', 'n', 't') with the streaming kernels."""

    def __init__(self, in_dim: int = 512, n_classes: int = 2, cls_type: str = "0000"):
        super().__init__()
        if len(cls_type) != 4 or any(ch not in ok for ch, ok in zip(cls_type, ("r0o", "n0o", "b0o", "t0o"))):
            raise KeyError(f"cls_type {cls_type!r}")                 # parse_param_coscls raises KeyError on unknown flags
        self.WeightNormR, self.weight_norm, self.bias, self.temp = (cls_type[0] == "r", cls_type[1] == "n",
                                                                    cls_type[2] == "b", cls_type[3] == "t")
        self.cls = nn.Conv2d(in_dim, n_classes, kernel_size=1, bias=self.bias)
        if self.WeightNormR:
            nn.utils.weight_norm(self.cls, "weight", dim=0)
        self.scale_factor = nn.Parameter(torch.tensor(2.0)) if self.temp else 2.0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x_norm = F.normalize(x, p=2, dim=1, eps=0.00001)
        if self.weight_norm:
            self.cls.weight.data = F.normalize(self.cls.weight.data, p=2, dim=1, eps=0.00001)
        return self.scale_factor * self.cls(x_norm)

    def reset_parameters(self) -> None:
        self.cls.reset_parameters()

    @property
    def plain(self) -> bool:
        return not (self.WeightNormR or self.weight_norm or self.bias or self.temp)


def inner_loop(classifier, f_s: torch.Tensor, s_label: torch.Tensor, cls_lr: float,
               adapt_iter: int, reset: bool = True, check: bool = True, algo: int = L.FIT_AUTO,
               loss_type: str = "wt_ce") -> None:
    """``PSPNet.inner_loop`` drop-in (src/model/pspnet.py:189-205): re-initialise the classifier, fit it on
    (f_s [S,C,h,w], s_label [S,H,W]) and write the result into its weight in place.

    ``loss_type`` is ``args.inner_loss_type`` (SegLoss, src/model/model_util.py:9-24):
      * ``'wt_ce'`` — class-weighted CE; the weight bg_cnt/fg_cnt of model_util.py:27-37 equals n0/n1 for labels
        in {0,1,255};
      * ``'ce'``    — plain ``nn.CrossEntropyLoss(ignore_index=255)`` (class weight [1, 1]);
      * ``'wt_dc'`` / ``'dc'`` — the per-channel sigmoid dice loss (weighted_dice_loss, model_util.py:40-73; both
        strings select the same function, model_util.py:18-19): :func:`fit_classifier_dice` (streaming kernels).
    ``classifier`` is the reference's ``nn.Conv2d(C, 2, 1, bias=False)`` (``dist == 'dot'``) or a :class:`CosCls`
    (``dist == 'cos'``) with any cls_type. 'oooo' / 'oobo': ``scale_factor * conv(F.normalize(x, eps=1e-5))`` is the same
    classifier on the features ``2 * x_norm``, which one extra kernel prepares (a bias is then scaled by 2 as well);
    a classifier with a bias is fitted by :func:`fit_classifier_bias` (weights and bias updated by the same SGD). With
    'r' / 'n' / 't' every parameter of ``classifier.parameters()`` (weight or weight_g + weight_v, bias, scale_factor) is
    fitted by :func:`fit_coscls`."""
    if loss_type not in ("wt_ce", "ce", "wt_dc", "dc"):
        loss_type = "wt_ce"                                     # SegLoss falls through to weighted CE for any other string
    if isinstance(classifier, CosCls) and (classifier.WeightNormR or classifier.weight_norm or classifier.temp):
        return _inner_loop_coscls(classifier, f_s, s_label, cls_lr, adapt_iter, reset, check, loss_type)
    if isinstance(classifier, CosCls):
        scale = float(classifier.scale_factor)
        conv, feat = classifier.cls, normalize_features(f_s, eps=1e-5, scale=scale)
    else:
        conv, feat, scale = classifier, f_s, 1.0
    if conv.weight.shape[0] != 2 or conv.kernel_size != (1, 1):
        raise NotImplementedError("cwt_b200 inner_loop fits the reference's 2-class 1x1 classifier")
    if reset:
        classifier.reset_parameters()                      # CPU/GPU generator order as in the reference
    C = conv.weight.shape[1]
    w0 = conv.weight.detach().reshape(2, C).to(f_s.device, torch.float32)
    if conv.bias is not None:                              # CosCls cls_type[2] == 'b' (or any nn.Conv2d(C, 2, 1, bias=True))
        if loss_type in ("wt_dc", "dc"):
            raise NotImplementedError("cwt_b200 inner_loop: the dice losses are built for the bias-free classifier")
        cw = torch.ones(2, dtype=torch.float32, device=f_s.device) if loss_type == "ce" else None
        b0 = conv.bias.detach().to(f_s.device, torch.float32)
        w, b = fit_classifier_bias(feat, s_label, w0, b0, cls_lr, adapt_iter, class_weight=cw, bias_scale=scale, check=check)
        with torch.no_grad():
            conv.bias.copy_(b)
    elif loss_type in ("wt_dc", "dc"):
        w = fit_classifier_dice(feat, s_label, w0, cls_lr, adapt_iter, check=check)
    else:
        cw = torch.ones(2, dtype=torch.float32, device=f_s.device) if loss_type == "ce" else None
        w = fit_classifier(feat, s_label, w0, cls_lr, adapt_iter, class_weight=cw, check=check, algo=algo)
    with torch.no_grad():
        conv.weight.copy_(w.reshape(2, C, 1, 1))


def _inner_loop_coscls(classifier: CosCls, f_s, s_label, cls_lr, adapt_iter, reset, check, loss_type) -> None:
    """CosCls with weight-norm ('r'), per-forward weight normalisation ('n') or a learnable temperature ('t'), with or
    without a bias: every parameter of ``classifier.parameters()`` is updated by the same SGD (cwt_fit_coscls_f32)."""
    if loss_type in ("wt_dc", "dc"):
        raise NotImplementedError("cwt_b200 inner_loop: the dice losses are built for the flag-free classifier")
    conv = classifier.cls
    if conv.weight.shape[0] != 2 or conv.kernel_size != (1, 1):
        raise NotImplementedError("cwt_b200 inner_loop fits the reference's 2-class 1x1 classifier")
    if reset:
        classifier.reset_parameters()       # with 'r' this re-draws only the derived .weight, not weight_g / weight_v — as in the reference
    dev, C = f_s.device, conv.weight.shape[1]
    if f_s.dim() != 4:
        raise ValueError("inner_loop fits ONE classifier: pass f_s [S,C,h,w] and s_label [S,H,W]")
    x, lab, E = normalize_features(f_s, eps=1e-5, scale=1.0).unsqueeze(0), s_label.unsqueeze(0), 1
    rep = lambda t, shape: t.detach().to(dev, torch.float32).reshape(1, *shape).expand(E, *shape).contiguous()
    flags = (L.COSCLS_R if classifier.WeightNormR else 0) | (L.COSCLS_N if classifier.weight_norm else 0) | \
            (L.COSCLS_T if classifier.temp else 0)
    wsrc = conv.weight_v if classifier.WeightNormR else conv.weight
    out = fit_coscls(x, lab, rep(wsrc, (2, C)), torch.full((E,), float(classifier.scale_factor), device=dev), cls_lr, adapt_iter,
                     weight_g=rep(conv.weight_g, (2,)) if classifier.WeightNormR else None,
                     bias=rep(conv.bias, (2,)) if conv.bias is not None else None, flags=flags,
                     class_weight=torch.ones(2, dtype=torch.float32, device=dev) if loss_type == "ce" else None, check=check)
    with torch.no_grad():
        wsrc.copy_(out["weight"][0].reshape(wsrc.shape))
        if classifier.WeightNormR:
            conv.weight_g.copy_(out["weight_g"][0].reshape(conv.weight_g.shape))
        if conv.bias is not None:
            conv.bias.copy_(out["bias"][0])
        if classifier.temp:
            classifier.scale_factor.copy_(out["scale"][0])


def increment_inner_loop(classifier: nn.Conv2d, f_s: torch.Tensor, s_label: torch.Tensor, cls_idx: int, cls_lr: float,
                         adapt_iter: int, tp: float = 1.0, check: bool = True, algo: int = L.FIT_AUTO) -> None:
    """``PSPNet.increment_inner_loop`` drop-in (src/model/pspnet.py:207-221): continue fitting ``classifier`` (NO
    parameter reset, unlike :func:`inner_loop`) with ``Adapt_SegLoss(num_cls, fg_idx=cls_idx, tp)`` ->
    weighted_adpt_ce_loss (src/model/model_util.py:76-98): CE with ``weight[cls_idx] = (bg_cnt / fg_cnt) ** tp``,
    bg_cnt = every non-ignored pixel that is not ``cls_idx``. The weight is computed on the device (no host sync; a
    support mask without ``cls_idx`` pixels gives an infinite weight exactly as the reference's tensor division does).
    ``classifier`` = ``self.classifier`` (``meta_train``) or ``self.val_classifier``: 2 classes run on the fused binary
    fit; more classes (the multi-way setting of src/train_cca.py, 16 / 17 on PASCAL, 61 / 62 on COCO) on
    :func:`fit_multiclass`."""
    if not isinstance(classifier, nn.Conv2d) or classifier.bias is not None or classifier.kernel_size != (1, 1):
        raise NotImplementedError("cwt_b200 increment_inner_loop fits a bias-free 1x1 classifier")
    K, C = classifier.weight.shape[:2]
    if not 0 <= cls_idx < K:
        raise IndexError(f"index {cls_idx} is out of bounds for dimension 0 with size {K}")     # weight[fg_idx] in the reference
    dev = f_s.device
    if K == 2:
        n = label_counts(s_label).reshape(-1, 4).sum(0)              # (#0, #1, #ignored, #invalid) over all shots
        fg, bg = n[cls_idx].float(), n[1 - cls_idx].float()
        cw = torch.ones(2, dtype=torch.float32, device=dev)
        cw[cls_idx] = (bg / fg) ** tp
        w0 = classifier.weight.detach().reshape(2, C).to(dev, torch.float32)
        w = fit_classifier(f_s, s_label, w0, cls_lr, adapt_iter, class_weight=cw, check=check, algo=algo)
    else:
        lab = s_label.long()
        fg = (lab == cls_idx).sum()
        bg = (lab != 255).sum() - fg                                   # model_util.py:90: everything that is not cls_idx nor 255
        cw = torch.ones(K, dtype=torch.float32, device=dev)
        cw[cls_idx] = (bg / fg) ** tp
        w = fit_multiclass(f_s, s_label, classifier.weight.detach().to(dev, torch.float32), cw, cls_lr, adapt_iter, check=check)
    with torch.no_grad():
        classifier.weight.copy_(w.reshape(K, C, 1, 1))
