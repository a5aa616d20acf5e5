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

from . import _lib as L
from .ops import fit_classifier


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


def inner_loop(classifier: nn.Conv2d, f_s: torch.Tensor, s_label: torch.Tensor, cls_lr: float,
               adapt_iter: int, reset: bool = True, check: bool = True, algo: int = L.FIT_AUTO) -> None:
    """``PSPNet.inner_loop`` drop-in (src/model/pspnet.py:189-205, loss 'wt_ce'): re-initialise the
    classifier, fit it on (f_s [S,C,h,w], s_label [S,H,W]) and write the result into
    ``classifier.weight`` in place. The class weight bg_cnt/fg_cnt of model_util.py:27-37 equals
    n0/n1 for labels in {0,1,255}."""
    if classifier.bias is not None or classifier.weight.shape[0] != 2 or classifier.kernel_size != (1, 1):
        raise NotImplementedError("cwt_b200 inner_loop fits the reference's 2-class bias-free 1x1 classifier")
    if reset:
        classifier.reset_parameters()                      # CPU/GPU generator order as in the reference
    C = classifier.weight.shape[1]
    w0 = classifier.weight.detach().reshape(2, C).to(f_s.device, torch.float32)
    w = fit_classifier(f_s, s_label, w0, cls_lr, adapt_iter, check=check, algo=algo)
    with torch.no_grad():
        classifier.weight.copy_(w.reshape(2, C, 1, 1))
