"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product package.

CPU restatement (PyTorch fp32/fp64, the same ATen ops the reference executes) of the
per-episode few-shot segmentation head of TeamOfProfGuo/Few_Shot_Seg_CWT:

    support-classifier fit  ->  Classifier Weight Transformer  ->  query logits /
    473x473 upsample / argmax / intersection-union

Every function cites the reference ``file:line`` it follows (paths relative to the
reference checkout). Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the
checker or as the timed CPU baseline — never as a product code path.

Parity pin: the reference repository has **no tests, fixtures or golden vectors** of its
own, so this restatement is pinned against the *live reference code*
imported from ``/root/reference`` (``oracle/pin_against_reference.py`` — the reference's
own ``MultiHeadAttentionOne``, ``batch_intersectionAndUnionGPU``, ``intersectionAndUnionGPU``
objects and a literal transcription of the ``src/test.py:162-234`` episode body that
calls them) and the outputs are committed under ``tests/golden/``. All arithmetic on
the path is third-party ATen (torch; the reference pins torch==1.6.0 in README prose
only, README.md:11-15; this container has torch 2.11.0).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

IGNORE = 255


# --------------------------------------------------------------------------------------
# (a-2) class weight                                              src/test.py:169-175
# --------------------------------------------------------------------------------------
def class_weight_ref(s_label: torch.Tensor) -> torch.Tensor:
    """``[1.0, n0 / n1]`` with n0/n1 counted over all shots jointly; python float division
    (raises ZeroDivisionError when the support mask is empty), stored as fp32."""
    arr = s_label.cpu().numpy()
    n0 = int((arr == 0).sum())
    n1 = int((arr == 1).sum())
    return torch.tensor([1.0, n0 / n1])


# --------------------------------------------------------------------------------------
# (a-1, a-3) support classifier fit                               src/test.py:164-187
#            (= src/train.py:206-231, src/model/pspnet.py:189-205 inner_loop)
# --------------------------------------------------------------------------------------
def fit_classifier_ref(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float,
                       n_iter: int, class_weight: Optional[torch.Tensor] = None,
                       return_losses: bool = False, dtype: torch.dtype = torch.float32):
    """Literal loop: 1x1 conv -> bilinear upsample (align_corners) to the label size ->
    class-weighted CE (ignore 255) -> backward -> plain SGD step.

    f_s [S,C,h,w]; s_label [S,H,W] (any integer dtype); w0 [2,C]. Returns W_fit [2,C]
    (and the per-step losses)."""
    S, C = f_s.shape[:2]
    f_s = f_s.to(dtype)
    tgt = s_label.long()
    clf = nn.Conv2d(C, 2, kernel_size=1, bias=False).to(dtype)
    with torch.no_grad():
        clf.weight.copy_(w0.reshape(2, C, 1, 1).to(dtype))
    opt = torch.optim.SGD(clf.parameters(), lr=lr)
    wt = class_weight_ref(s_label) if class_weight is None else class_weight
    crit = nn.CrossEntropyLoss(weight=wt.to(dtype), ignore_index=IGNORE)
    losses = []
    for _ in range(n_iter):
        out = clf(f_s)
        out = F.interpolate(out, size=tgt.shape[-2:], mode="bilinear", align_corners=True)
        loss = crit(out, tgt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if return_losses:
            losses.append(float(loss))
    w = clf.weight.detach().reshape(2, C).clone()
    return (w, losses) if return_losses else w


# --------------------------------------------------------------------------------------
# PSPNet.inner_loop variants                      src/model/pspnet.py:189-205, 290-334
#   SegLoss 'wt_ce' / 'ce' / 'wt_dc' / 'dc' (src/model/model_util.py:9-73); classifier 'dot' or CosCls 'oooo'
# --------------------------------------------------------------------------------------
def initial_bias(idx: int, C: int) -> torch.Tensor:
    """Deterministic initial bias [2] ~ U(-1/sqrt(C), 1/sqrt(C)) — the distribution nn.Conv2d.reset_parameters draws a
    bias from (fan_in = C) — for the golden cases of classifiers with a bias."""
    g = torch.Generator().manual_seed(977 + idx)
    return (torch.rand(2, generator=g) * 2.0 - 1.0) / math.sqrt(C)


def weighted_dice_loss_ref(prediction: torch.Tensor, target_seg: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """SegLoss 'wt_dc' / 'dc' (model_util.py:18-19 -> weighted_dice_loss, :40-73, weighted_val 1.0, reduction 'sum',
    input_type 'lg'): every (image, channel) row is a sigmoid dice, 1 - 2 sum(t p) / clamp(sum p^2 + sum t^2, eps), with
    targets [label == 0, label == 1] (255 is in neither, but its p^2 counts); the rows are summed and divided by the
    number of images. prediction [S,2,H,W] logits, target_seg [S,H,W]."""
    tgt = torch.stack([target_seg == 0, target_seg == 1], dim=1).float()
    n, _, h, w = tgt.shape
    p = torch.sigmoid(prediction.reshape(-1, h, w)).reshape(-1, h * w)
    t = tgt.reshape(-1, h * w)
    part = (p ** 2).sum(dim=-1) + (t ** 2).sum(dim=-1)
    loss = 1 - 2 * (t * p).sum(dim=-1) / torch.clamp(part, min=eps)
    return loss.sum() / n


def inner_loop_ref(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float, n_iter: int,
                   loss_type: str = "wt_ce", dist: str = "dot", fg_idx: int = 1, tp: float = 1.0,
                   bias0: Optional[torch.Tensor] = None):
    """The inner loop with its two switches: ``args.inner_loss_type`` ('wt_ce': weight[fg] = bg_cnt / fg_cnt from
    torch.bincount, model_util.py:27-37; 'ce': plain CE, ignore 255; 'wt_dc' / 'dc': per-channel sigmoid dice) and ``args.dist`` ('dot': nn.Conv2d;
    'cos': CosCls with cls_type 'oooo' — scores = 2.0 * conv(F.normalize(x, p=2, dim=1, eps=1e-5)),
    pspnet.py:302-310). f_s [S,C,h,w]; s_label [S,H,W]; w0 [2,C] -> fitted weight [2,C].
    ``loss_type == 'adapt_ce'`` is ``PSPNet.increment_inner_loop`` (pspnet.py:207-221) for a 2-class classifier: the same
    loop (no parameter reset) with Adapt_SegLoss -> weighted_adpt_ce_loss (model_util.py:76-98):
    weight[fg_idx] = (bg_cnt / fg_cnt) ** tp.
    ``bias0`` [2]: the classifier has a bias (CosCls cls_type[2] == 'b', pspnet.py:294; its bias sits inside ``cls`` and
    is therefore multiplied by scale_factor too, pspnet.py:307-308); returns (weight [2,C], bias [2]) then."""
    S, C = f_s.shape[:2]
    tgt = s_label.long()
    conv = nn.Conv2d(C, 2, kernel_size=1, bias=bias0 is not None)
    with torch.no_grad():
        conv.weight.copy_(w0.reshape(2, C, 1, 1))
        if bias0 is not None:
            conv.bias.copy_(bias0)
    opt = torch.optim.SGD(conv.parameters(), lr=lr)
    for _ in range(n_iter):
        if dist == "dot":
            out = conv(f_s)
        else:
            out = 2.0 * conv(F.normalize(f_s, p=2, dim=1, eps=0.00001))
        out = F.interpolate(out, size=tgt.shape[-2:], mode="bilinear", align_corners=True)
        if loss_type in ("wt_dc", "dc"):
            loss = weighted_dice_loss_ref(out, tgt)
        elif loss_type == "ce":
            loss = nn.CrossEntropyLoss(ignore_index=IGNORE)(out, tgt)
        else:
            count = torch.bincount(tgt.view(-1))
            fg = count[1]
            bg = (torch.sum(count) - fg) if len(count) <= 255 else (torch.sum(count) - count[255] - fg)
            weight = torch.tensor([1.0, 1.0])
            if loss_type == "adapt_ce":
                fg = count[fg_idx]
                bg = (torch.sum(count) - fg) if len(count) <= 255 else (torch.sum(count) - count[255] - fg)
                weight[fg_idx] = (bg / fg) ** tp
            else:
                weight[1] = bg / fg
            loss = nn.CrossEntropyLoss(weight=weight, ignore_index=IGNORE)(out, tgt)
        opt.zero_grad()
        loss.backward()
        opt.step()
    if bias0 is not None:
        return conv.weight.detach().reshape(2, C).clone(), conv.bias.detach().clone()
    return conv.weight.detach().reshape(2, C).clone()


def coscls_inner_loop_ref(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float, n_iter: int,
                          cls_type: str, loss_type: str = "wt_ce", bias0: Optional[torch.Tensor] = None,
                          g0: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """PSPNet.inner_loop (pspnet.py:189-205) on CosCls with any ``cls_type`` (pspnet.py:290-323), restated with explicit
    parameters: 'r' -> WeightNorm.apply(cls, 'weight', dim=0) (parameters weight_g [2,1,1,1] = g0, weight_v = w0; the
    pre-forward hook rebuilds cls.weight, which also undoes 'n'); 'n' -> cls.weight.data re-normalised (dim=1, eps 1e-5)
    at every forward; 'b' -> bias (= bias0); 't' -> scale_factor is a parameter (2.0). SGD over classifier.parameters()."""
    fr, fn, fb, ft = (cls_type[0] == "r", cls_type[1] == "n", cls_type[2] == "b", cls_type[3] == "t")
    C = f_s.shape[1]
    tgt = s_label.long()
    W = w0.detach().clone().reshape(2, C, 1, 1).requires_grad_(True)              # cls.weight, or cls.weight_v with 'r'
    g = g0.detach().clone().reshape(2, 1, 1, 1).requires_grad_(True) if fr else None
    b = bias0.detach().clone().requires_grad_(True) if fb else None
    s = torch.tensor(2.0, requires_grad=ft)
    params = [p for p in (g, W, b, s if ft else None) if p is not None]
    opt = torch.optim.SGD(params, lr=lr)
    for _ in range(n_iter):
        x_norm = F.normalize(f_s, p=2, dim=1, eps=0.00001)
        if fr:
            weight = torch._weight_norm(W, g, 0)
        else:
            if fn:
                W.data = F.normalize(W.data, p=2, dim=1, eps=0.00001)
            weight = W
        out = s * F.conv2d(x_norm, weight, b)
        out = F.interpolate(out, size=tgt.shape[-2:], mode="bilinear", align_corners=True)
        if loss_type == "ce":
            loss = nn.CrossEntropyLoss(ignore_index=IGNORE)(out, tgt)
        else:
            count = torch.bincount(tgt.view(-1))
            fg = count[1]
            bg = (torch.sum(count) - fg) if len(count) <= 255 else (torch.sum(count) - count[255] - fg)
            wt = torch.tensor([1.0, 1.0])
            wt[1] = bg / fg
            loss = nn.CrossEntropyLoss(weight=wt, ignore_index=IGNORE)(out, tgt)
        opt.zero_grad()
        loss.backward()
        opt.step()
    return {"weight": W.detach().reshape(2, C).clone(), "weight_g": g.detach().reshape(2).clone() if fr else None,
            "bias": b.detach().clone() if fb else None, "scale": s.detach().clone()}


def multiclass_labels(s_label: torch.Tensor, K: int, fg_idx: int) -> torch.Tensor:
    """Deterministic K-way labels for the multi-way inner loop from a {0,1,255} support mask: foreground -> fg_idx,
    255 stays, background -> a block pattern over the other classes (what reset_spt_label, model_util.py:120-128,
    produces from the base classifier's prediction: pseudo-labels of the base classes on the background)."""
    H, W = s_label.shape[-2:]
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pat = ((yy // 11) + 2 * (xx // 13)) % K
    pat = torch.where(pat == fg_idx, torch.zeros_like(pat) if fg_idx != 0 else torch.ones_like(pat), pat)
    lab = s_label.long()
    out = torch.where(lab == 1, torch.full_like(lab, fg_idx), torch.where(lab == 0, pat.expand_as(lab), torch.full_like(lab, IGNORE)))
    return out


def multiclass_w0(idx: int, K: int, C: int) -> torch.Tensor:
    """Deterministic initial K-class weights ~ U(-1/sqrt(C), 1/sqrt(C)) (nn.Conv2d's init) for the golden cases."""
    g = torch.Generator().manual_seed(4409 + idx)
    return (torch.rand(K, C, generator=g) * 2.0 - 1.0) / math.sqrt(C)


def increment_inner_loop_ref(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float, n_iter: int,
                             fg_idx: int, tp: float = 1.0) -> torch.Tensor:
    """PSPNet.increment_inner_loop (pspnet.py:207-221) for a K-class classifier: SGD on Adapt_SegLoss ->
    weighted_adpt_ce_loss (model_util.py:87-98), weight = ones(K) with weight[fg_idx] = (bg_cnt / fg_cnt) ** tp.
    f_s [S,C,h,w]; s_label [S,H,W] in [0,K) or 255; w0 [K,C] -> fitted weight [K,C]."""
    K, C = w0.shape
    tgt = s_label.long()
    conv = nn.Conv2d(C, K, kernel_size=1, bias=False)
    with torch.no_grad():
        conv.weight.copy_(w0.reshape(K, C, 1, 1))
    opt = torch.optim.SGD(conv.parameters(), lr=lr)
    for _ in range(n_iter):
        out = F.interpolate(conv(f_s), size=tgt.shape[-2:], mode="bilinear", align_corners=True)
        count = torch.bincount(tgt.view(-1))
        fg = count[fg_idx]
        bg = (torch.sum(count) - fg) if len(count) <= 255 else (torch.sum(count) - count[255] - fg)
        weight = torch.tensor([1.0] * K)
        weight[fg_idx] = (bg / fg) ** tp
        loss = nn.CrossEntropyLoss(weight=weight, ignore_index=IGNORE)(out, tgt)
        opt.zero_grad()
        loss.backward()
        opt.step()
    return conv.weight.detach().reshape(K, C).clone()


def bilinear_matrix(n_in: int, n_out: int, dtype=torch.float64) -> torch.Tensor:
    """Dense [n_out, n_in] align_corners=True interpolation matrix (2 non-zeros per row).
    ATen: scale = (n_in-1)/(n_out-1); src = scale*dst; i0 = floor(src); l1 = src - i0."""
    B = torch.zeros(n_out, n_in, dtype=dtype)
    scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
    for o in range(n_out):
        src = scale * o
        i0 = min(int(math.floor(src)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        l1 = src - i0
        B[o, i0] += 1.0 - l1
        B[o, i1] += l1
    return B


def fit_classifier_closed_form(f_s, s_label, w0, lr, n_iter, class_weight=None,
                               dtype=torch.float64):
    """Independent restatement of the same loop with explicit matrices — no autograd,
    no F.interpolate, no nn.CrossEntropyLoss (SURVEY.md §8 a-3 formula):
        L60 = W F ; U = Bh L60 Bw^T ; p = softmax_c(U) ; G = w_y (p - 1_y) / sum w_y
        G60 = Bh^T G Bw ; dW = G60 F^T ; W <- W - lr dW
    Used to cross-check fit_classifier_ref and to study fp32 vs fp64 drift."""
    S, C, h, w = f_s.shape
    H, W = s_label.shape[-2:]
    Fm = f_s.to(dtype).reshape(S, C, h * w)
    y = s_label.long()
    wt = (class_weight_ref(s_label) if class_weight is None else class_weight).to(dtype)
    Bh = bilinear_matrix(h, H, dtype)
    Bw = bilinear_matrix(w, W, dtype)
    valid = (y != IGNORE)
    ys = torch.where(valid, y, torch.zeros_like(y))
    wy = wt[ys] * valid.to(dtype)                       # [S,H,W]
    denom = wy.sum()
    onehot = torch.stack([(ys == 0), (ys == 1)], 1).to(dtype)  # [S,2,H,W]
    Wc = w0.to(dtype).clone()
    for _ in range(n_iter):
        L60 = torch.einsum("kc,scp->skp", Wc, Fm).reshape(S, 2, h, w)
        U = torch.einsum("Yy,skyx,Xx->skYX", Bh, L60, Bw)
        p = torch.softmax(U, dim=1)
        G = (p - onehot) * (wy / denom).unsqueeze(1)
        G60 = torch.einsum("Yy,skYX,Xx->skyx", Bh, G, Bw).reshape(S, 2, h * w)
        dW = torch.einsum("skp,scp->kc", G60, Fm)
        Wc = Wc - lr * dW
    return Wc


# --------------------------------------------------------------------------------------
# (a-6) MultiHeadAttentionOne                         src/model/transformer.py:12-83
# --------------------------------------------------------------------------------------
def mha_one_forward_ref(q: torch.Tensor, k: torch.Tensor, params: Dict[str, torch.Tensor],
                        n_head: int, keep_attn: Optional[torch.Tensor] = None,
                        keep_out: Optional[torch.Tensor] = None,
                        p_attn: float = 0.1, p_out: float = 0.5) -> torch.Tensor:
    """Functional restatement of ``MultiHeadAttentionOne.forward(q, k, v)`` with ``v is k``.

    q [B,Lq,C] (the classifier weights, Lq=2); k [B,C,h,w] (normalised query features).
    params: reference state-dict names. Eval mode when both keep masks are None; otherwise
    inverted dropout with explicit uint8/bool keep-masks: ``keep_attn [nH*B, Lq, HW]``
    (head-major, as the reference's permute(2,0,1,3) lays heads out; transformer.py:71-75)
    scaled by 1/(1-p_attn) and ``keep_out [B,Lq,C]`` scaled by 1/(1-p_out)
    (transformer.py:17-20,28,52,80)."""
    A = params["w_qkvs.weight"]
    B_, Lq, C = q.shape
    d_k = A.shape[0] // n_head
    X = k.reshape(B_, k.shape[1], -1).permute(0, 2, 1)           # [B,HW,C]   :55-59
    HW = X.shape[1]
    Q = (q @ A.t()).view(B_, Lq, n_head, d_k)                     # :67
    K = (X @ A.t()).view(B_, HW, n_head, d_k)                     # :68-69 (k is v, one weight)
    Qh = Q.permute(2, 0, 1, 3).reshape(-1, Lq, d_k)               # [(nH*B),Lq,dk]  :71
    Kh = K.permute(2, 0, 1, 3).reshape(-1, HW, d_k)               # :72-73
    attn = torch.bmm(Qh, Kh.transpose(1, 2)) / float(np.power(d_k, 0.5))   # :24-25,47
    attn = torch.softmax(attn, dim=2)                             # :27
    if keep_attn is not None:
        attn = attn * keep_attn.to(attn.dtype) / (1.0 - p_attn)   # :28
    out = torch.bmm(attn, Kh)                                     # :29
    out = out.view(n_head, B_, Lq, d_k).permute(1, 2, 0, 3).reshape(B_, Lq, -1)  # :77-78
    out = out @ params["fc.weight"].t() + params["fc.bias"]       # :80
    if keep_out is not None:
        out = out * keep_out.to(out.dtype) / (1.0 - p_out)        # :80
    return F.layer_norm(out + q, (C,), params["layer_norm.weight"],
                        params["layer_norm.bias"], 1e-5)          # :81


# --------------------------------------------------------------------------------------
# (a-9, a-10) intersection / union                                src/util.py:237-308
# --------------------------------------------------------------------------------------
def intersection_and_union_ref(preds: torch.Tensor, target: torch.Tensor, num_classes: int,
                               ignore_index: int = IGNORE):
    """preds, target integer tensors of identical shape. Returns float32 (I, U, T) [num_classes]
    computed with histc exactly as src/util.py:297-307 (ignored pixels fall outside the
    histc range). Unlike the reference it does not mutate the caller's ``preds``."""
    assert preds.dim() in [1, 2, 3]
    assert preds.shape == target.shape
    preds = preds.reshape(-1).clone()
    target = target.reshape(-1)
    preds[target == ignore_index] = ignore_index
    inter = preds[preds == target]
    a_i = torch.histc(inter.float(), bins=num_classes, min=0, max=num_classes - 1)
    a_o = torch.histc(preds.float(), bins=num_classes, min=0, max=num_classes - 1)
    a_t = torch.histc(target.float(), bins=num_classes, min=0, max=num_classes - 1)
    return a_i, a_o + a_t - a_i, a_t


def batch_intersection_and_union_ref(logits: torch.Tensor, target: torch.Tensor, num_classes: int,
                                     ignore_index: int = IGNORE):
    """logits [n_task,shot,C,h,w]; target [n_task,shot,H,W] -> (I,U,T) float32 [n_task,shot,C]
    (src/util.py:237-277: bilinear align_corners upsample, argmax over classes, per-item histc)."""
    n_task, shots, ncls, h, w = logits.shape
    H, W = target.shape[-2:]
    up = F.interpolate(logits.reshape(n_task * shots, ncls, h, w), size=(H, W), mode="bilinear",
                       align_corners=True).view(n_task, shots, ncls, H, W)
    preds = up.argmax(2)
    I = torch.zeros(n_task, shots, ncls)
    U = torch.zeros(n_task, shots, ncls)
    T = torch.zeros(n_task, shots, ncls)
    for t in range(n_task):
        for s in range(shots):
            i, u, tt = intersection_and_union_ref(preds[t][s], target[t][s].long(), ncls, ignore_index)
            I[t, s], U[t, s], T[t, s] = i, u, tt
    return I, U, T


# --------------------------------------------------------------------------------------
# the whole episode body                                          src/test.py:162-234
# --------------------------------------------------------------------------------------
def episode_ref(f_s, s_label, f_q, q_label, w0, params, n_head: int, lr: float, n_iter: int,
                dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """One evaluation episode on pre-computed features (the frozen backbone is outside
    the path). f_s [S,C,h,w]; s_label [S,H,W]; f_q [C,h,w]; q_label [H,W]; w0 [2,C].

    Returns W_fit, W_adapted [2,C]; logits60 / logits60_0 [2,h,w] (adapted / baseline);
    counts / counts0 int64 [2 classes, 3 = I,U,T]; loss (unweighted CE of the upsampled
    adapted logits, src/test.py:222-223); tie_margin [H,W] = |up(l1) - up(l0)| of the adapted
    logits (for the stated argmax-tie set)."""
    C, h, w = f_q.shape
    H, W = q_label.shape
    w_fit = fit_classifier_ref(f_s, s_label, w0, lr, n_iter, dtype=dtype)          # :164-187
    with torch.no_grad():
        fq = f_q.to(dtype).unsqueeze(0)
        pred_q0 = F.conv2d(fq, w_fit.view(2, C, 1, 1))                              # :192
        fqn = F.normalize(fq, dim=1)                                                # :194
        pd = {k: v.to(dtype) for k, v in params.items()}
        w_ad = mha_one_forward_ref(w_fit.view(1, 2, C), fqn, pd, n_head)            # :195-197
        pred_q = F.conv2d(fqn, w_ad.view(2, C, 1, 1))                               # :200-204
        tgt = q_label.long().view(1, 1, H, W)
        up = F.interpolate(pred_q.float(), size=(H, W), mode="bilinear", align_corners=True)   # :214
        up0 = F.interpolate(pred_q0.float(), size=(H, W), mode="bilinear", align_corners=True)  # :215
        I, U, T = batch_intersection_and_union_ref(up.unsqueeze(1), tgt, 2)         # :216
        I0, U0, T0 = batch_intersection_and_union_ref(up0.unsqueeze(1), tgt, 2)     # :218
        loss = F.cross_entropy(up, tgt.view(1, H, W), ignore_index=IGNORE)          # :222-223
    counts = torch.stack([I[0, 0], U[0, 0], T[0, 0]], 1).round().long()
    counts0 = torch.stack([I0[0, 0], U0[0, 0], T0[0, 0]], 1).round().long()
    return {
        "W_fit": w_fit, "W_adapted": w_ad[0], "logits60": pred_q[0], "logits60_0": pred_q0[0],
        "counts": counts, "counts0": counts0, "loss": loss,
        "tie_margin": (up[0, 1] - up[0, 0]).abs(), "tie_margin0": (up0[0, 1] - up0[0, 0]).abs(),
    }


def miou_from_counts(cls_I: Dict[int, float], cls_U: Dict[int, float]) -> float:
    """src/test.py:232-243: IoU_c = I_c / (U_c + 1e-10), mean over classes seen."""
    return float(np.mean([cls_I[c] / (cls_U[c] + 1e-10) for c in cls_U]))


# --------------------------------------------------------------------------------------
# (a-13) meta-training step of the transformer                    src/train.py:233-267
# --------------------------------------------------------------------------------------
def query_class_weight_ref(q_label: torch.Tensor) -> torch.Tensor:
    """src/train.py:237-243: [1, n0 / (n1 + 1e-12)]."""
    arr = q_label.cpu().numpy()
    n0 = int((arr == 0).sum())
    n1 = int((arr == 1).sum())
    return torch.tensor([1.0, n0 / (n1 + 1e-12)])


def meta_train_step_ref(w_fit, f_q, q_label, params, n_head, keep_attn=None, keep_out=None,
                        p_attn=0.1, p_out=0.5, dtype=torch.float32):
    """Forward a-5..a-7 in train mode, weighted CE at full resolution, backward to the five
    transformer parameters (autograd). No gradient flows to w_fit (it enters as ``.data``,
    src/train.py:253) nor to f_q (``no_grad``, :246-250).

    Returns dict(loss, W_adapted, grads{name: tensor})."""
    C, h, w = f_q.shape
    H, W = q_label.shape
    pd = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in params.items()}
    with torch.no_grad():
        fqn = F.normalize(f_q.to(dtype).unsqueeze(0), dim=1)                        # :250
    q = w_fit.detach().to(dtype).view(1, 2, C)
    w_ad = mha_one_forward_ref(q, fqn, pd, n_head, keep_attn, keep_out, p_attn, p_out)   # :257
    pred = torch.matmul(w_ad, fqn.view(1, C, -1)).view(1, 2, h, w)                  # :259-261
    pred = F.interpolate(pred, size=(H, W), mode="bilinear", align_corners=True)    # :262
    wt = query_class_weight_ref(q_label).to(dtype)
    loss = F.cross_entropy(pred, q_label.long().view(1, H, W), weight=wt, ignore_index=IGNORE)  # :264
    loss.backward()                                                                  # :266
    return {"loss": loss.detach(), "W_adapted": w_ad.detach()[0],
            "grads": {k: v.grad.detach() for k, v in pd.items()}}


def sgd_nesterov_step_ref(params, grads, bufs, lr, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD(momentum, nesterov=True, weight_decay) update (src/optimizer.py:11-15)."""
    new_p, new_b = {}, {}
    for k in params:
        g = grads[k] + weight_decay * params[k]
        b = g.clone() if bufs is None or k not in bufs else momentum * bufs[k] + g
        new_b[k] = b
        new_p[k] = params[k] - lr * (g + momentum * b)
    return new_p, new_b


# --------------------------------------------------------------------------------------
# (f-4) validation transform                      src/dataset/transform.py:58-163, dataset.py:78-84
# --------------------------------------------------------------------------------------
def synthetic_image(idx: int, h: int, w: int):
    """Deterministic test image [h,w,3] float32 in [0,255] (integers, as np.float32(cv2.imread(...)) delivers,
    dataset.py:138-140) and label [h,w] uint8 in {0,1,255}."""
    g = np.random.RandomState(7000 + idx)
    base = g.randint(0, 256, size=(h // 7 + 2, w // 7 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((7, 7, 1), np.float32))[:h, :w] + g.randint(-9, 10, size=(h, w, 3))
    img = np.clip(img, 0, 255).astype(np.float32)
    lab = np.zeros((h, w), np.uint8)
    lab[h // 4: h // 4 + h // 3, w // 5: w // 5 + w // 2] = 1
    lab[h // 4 - 2: h // 4, :] = 255
    return img, lab


def val_transform_ref(image: np.ndarray, label: np.ndarray, size: int, mean, std, padding=None, plain_cv2: bool = False):
    """Resize -> ToTensor -> Normalize exactly as the reference composes them (dataset.py:78-84), with cv2 doing the
    resampling as in transform.py:143-157 (cv2 = the reference's third-party dependency for this step).
    ``plain_cv2``: run cv2.resize with ``cv2.setUseOptimized(False)`` — OpenCV's portable C++ resampler instead of its
    IPP / SIMD one. The two differ by up to 0.0044 grey levels (of 255) on PASCAL-sized images [measured, cv2 4.13]; the
    CUDA kernel reproduces the portable path bit for bit."""
    import cv2
    was_optimized = cv2.useOptimized()
    if plain_cv2:
        cv2.setUseOptimized(False)
    try:
        return _val_transform_ref(cv2, image, label, size, mean, std, padding)
    finally:
        cv2.setUseOptimized(was_optimized)


def _val_transform_ref(cv2, image, label, size, mean, std, padding):

    def find_new_hw(ori_h, ori_w, test_size):                     # transform.py:117-137
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

    new_h, new_w = find_new_hw(image.shape[0], image.shape[1], size)
    image_crop = cv2.resize(image, dsize=(int(new_w), int(new_h)), interpolation=cv2.INTER_LINEAR)      # :143-144
    back_crop = np.zeros((size, size, 3))                                                                # :147 (float64)
    if padding:
        back_crop[:, :, 0], back_crop[:, :, 1], back_crop[:, :, 2] = padding[0], padding[1], padding[2]
    back_crop[:new_h, :new_w, :] = image_crop
    s_mask = cv2.resize(label.astype(np.float32), dsize=(int(new_w), int(new_h)), interpolation=cv2.INTER_NEAREST)  # :156-157
    back_mask = np.ones((size, size)) * 255
    back_mask[:new_h, :new_w] = s_mask
    img = torch.from_numpy(back_crop.transpose((2, 0, 1)))                                               # ToTensor :70-72
    if not isinstance(img, torch.FloatTensor):
        img = img.float().div(255)
    lab = torch.from_numpy(back_mask).long()
    for t, m, s in zip(img, mean, std):                                                                  # Normalize :101-102
        t.sub_(m).div_(s)
    return img, lab
