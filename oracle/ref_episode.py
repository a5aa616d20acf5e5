"""TEST INFRASTRUCTURE (never imported by the product): the episode body of src/test.py:162-234 with the REFERENCE'S OWN
objects on the path — its ``MultiHeadAttentionOne`` module and its ``batch_intersectionAndUnionGPU`` function — loaded either
from the live checkout (/root/reference, this container only) or from oracle/_ref (the verbatim copy made by
oracle/make_ref.py, which travels to the GPU box). Used by oracle/pin_against_reference.py to pin the restatement
(oracle/head_ref.py), and by bench.py's CPU legs (`--impl reference`, `cpu_baseline`) when oracle/_ref exists."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(HERE, "_ref")


def load_reference_modules(prefer_live: bool = True):
    """-> (MultiHeadAttentionOne, batch_intersectionAndUnionGPU, intersectionAndUnionGPU, where) or None."""
    live = os.environ.get("CWT_REFERENCE", "/root/reference")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if prefer_live and os.path.isdir(os.path.join(live, "src")):
            sys.path.insert(0, live)
            try:
                from src.model.transformer import MultiHeadAttentionOne
                from src.util import batch_intersectionAndUnionGPU, intersectionAndUnionGPU
                return MultiHeadAttentionOne, batch_intersectionAndUnionGPU, intersectionAndUnionGPU, live
            finally:
                sys.path.remove(live)
        if os.path.isfile(os.path.join(_REF_DIR, "ref_src", "model", "transformer.py")):
            sys.path.insert(0, _REF_DIR)
            try:
                from ref_src.model.transformer import MultiHeadAttentionOne
                from ref_src.util import batch_intersectionAndUnionGPU, intersectionAndUnionGPU
                return MultiHeadAttentionOne, batch_intersectionAndUnionGPU, intersectionAndUnionGPU, "oracle/_ref"
            finally:
                sys.path.remove(_REF_DIR)
    return None


def build_reference_transformer(MHA, params, n_head, C, dropout=0.5):
    m = MHA(n_head, C, C, C, dropout=dropout)       # ctor as src/test.py:57 / src/train.py:96
    m.load_state_dict(params)                        # names pinned by SURVEY.md §5
    return m


def episode_via_reference(ep, params, n_head, lr, n_iter, MHA, batch_iou):
    """src/test.py:162-234 with the reference's own objects on the path (CPU, fp32)."""
    C, h, w = ep.f_q.shape
    H, W = ep.q_label.shape
    transformer = build_reference_transformer(MHA, params, n_head, C).eval()
    s_label = ep.s_label.long().unsqueeze(0)                                   # [1,S,H,W]
    binary_classifier = nn.Conv2d(C, 2, kernel_size=1, bias=False)
    with torch.no_grad():
        binary_classifier.weight.copy_(ep.w0.view(2, C, 1, 1))                # explicit W0 (a-1)
    optimizer = torch.optim.SGD(binary_classifier.parameters(), lr=lr)
    arr = s_label.numpy().copy()
    back_pix, target_pix = np.where(arr == 0), np.where(arr == 1)
    criterion = nn.CrossEntropyLoss(weight=torch.tensor([1.0, len(back_pix[0]) / len(target_pix[0])]),
                                    ignore_index=255)
    f_s = ep.f_s
    for _ in range(n_iter):
        out = binary_classifier(f_s)
        out = F.interpolate(out, size=s_label.size()[2:], mode="bilinear", align_corners=True)
        loss = criterion(out, s_label.squeeze(0))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    with torch.no_grad():
        f_q = ep.f_q.unsqueeze(0)
        pred_q0 = binary_classifier(f_q)
        f_q = F.normalize(f_q, dim=1)
        weights_cls = binary_classifier.weight.data
        wr = weights_cls.squeeze().unsqueeze(0).expand(f_q.shape[0], 2, C)
        updated = transformer(wr, f_q, f_q)
        pseudo = nn.Conv2d(C, 2, kernel_size=1, bias=False)
        pseudo.weight.data = torch.as_tensor(updated.squeeze(0).unsqueeze(2).unsqueeze(3))
        pred_q = pseudo(f_q)
    logits_q = pred_q.detach().unsqueeze(0)          # [1,1,2,h,w]
    logits_q0 = pred_q0.detach().unsqueeze(0)
    gt_q = ep.q_label.long().view(1, 1, H, W)
    logits = F.interpolate(logits_q.squeeze(1), size=(H, W), mode="bilinear", align_corners=True).detach()
    logits0 = F.interpolate(logits_q0.squeeze(1), size=(H, W), mode="bilinear", align_corners=True).detach()
    I, U, T = batch_iou(logits.unsqueeze(1), gt_q.clone(), 2)
    I0, U0, T0 = batch_iou(logits0.unsqueeze(1), gt_q.clone(), 2)
    loss = nn.CrossEntropyLoss(ignore_index=255)(logits, gt_q.squeeze(1))
    return {
        "W_fit": weights_cls.view(2, C).clone(), "W_adapted": updated[0].clone(),
        "logits60": pred_q[0].clone(), "logits60_0": pred_q0[0].clone(),
        "counts": torch.stack([I[0, 0], U[0, 0], T[0, 0]], 1).round().long(),
        "counts0": torch.stack([I0[0, 0], U0[0, 0], T0[0, 0]], 1).round().long(),
        "loss": loss,
        "tie_margin": (logits[0, 1] - logits[0, 0]).abs(),
        "tie_margin0": (logits0[0, 1] - logits0[0, 0]).abs(),
    }
