"""Functional API over the C ABI — batched over E episodes, CUDA tensors in / CUDA tensors out.

Each function is also registered as a PyTorch custom op (``torch.ops.cwt_b200.*``, see
``_register_custom_ops``) so the head can sit inside torch graphs / profilers like any ATen op.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L

IGNORE = 255


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _episode_inputs(f_s: torch.Tensor, s_label: torch.Tensor, name: str = "f_s"):
    """Common shape handling of the fit wrappers: ([E,]S,C,h,w) features + ([E,]S,H,W) labels -> contiguous 5-d / 4-d."""
    if f_s.dim() == 4:
        f_s, s_label = f_s.unsqueeze(0), s_label.unsqueeze(0)
    if f_s.dim() != 5 or s_label.dim() != 4:
        raise ValueError(f"{name} must be [E,S,C,h,w] and s_label [E,S,H,W]; got {tuple(f_s.shape)}, {tuple(s_label.shape)}")
    if s_label.shape[0] != f_s.shape[0] or s_label.shape[1] != f_s.shape[1]:
        raise ValueError(f"{name} and s_label disagree on E or S")
    return _f32c(f_s, name), s_label.contiguous()


def _class_weight(class_weight: Optional[torch.Tensor], E: int, dev) -> Optional[torch.Tensor]:
    if class_weight is None:
        return None
    cw = _f32c(class_weight.to(dev), "class_weight").reshape(-1, 2)
    if cw.shape[0] == 1 and E > 1:
        cw = cw.expand(E, 2).contiguous()
    if cw.shape[0] != E:
        raise ValueError("class_weight must be [2] or [E,2]")
    return cw


def _raise_on_label_counts(counts: torch.Tensor, has_class_weight: bool) -> None:
    """The synchronous ``check=True`` path: one host sync, the reference's own exceptions."""
    c = counts.cpu()
    if int(c[:, 3].sum()) > 0:
        raise ValueError("s_label holds values outside {0, 1, ignore_index}")
    if not has_class_weight and bool((c[:, 1] == 0).any()):
        raise ZeroDivisionError("division by zero: an episode's support mask has no foreground pixel "
                                "(reference: len(back_pix[0]) / len(target_pix[0]), src/test.py:174)")


def fit_status(label_counts_: torch.Tensor, w_fit: torch.Tensor, has_class_weight: bool = False) -> torch.Tensor:
    """Deferred error word per episode (int32 [E], device, no host sync): bits L.FIT_BAD_LABEL / L.FIT_NO_FG /
    L.FIT_NONFINITE. Copy it to the host with the episode's results and hand it to :func:`raise_for_status`."""
    dev = L.require_cuda(label_counts_, w_fit)
    E = label_counts_.shape[0]
    Cc = w_fit.shape[-1]
    status = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().cwt_fit_status(L.ptr(label_counts_.contiguous()), L.ptr(_f32c(w_fit, "w_fit")), int(has_class_weight),
                                     L.ptr(status), E, Cc, L.stream_ptr(dev))
    L.check(rc, "cwt_fit_status")
    return status


def raise_for_status(status_host: torch.Tensor, first_episode: int = 0) -> None:
    """Raise what the reference would have raised for the first bad episode of a (host) status vector."""
    bad = torch.nonzero(status_host.reshape(-1)).reshape(-1)
    if bad.numel() == 0:
        return
    i = int(bad[0]); code = int(status_host.reshape(-1)[i]); where = f"episode {first_episode + i}"
    if code & L.FIT_BAD_LABEL:
        raise ValueError(f"{where}: s_label holds values outside {{0, 1, ignore_index}}")
    if code & L.FIT_NO_FG:
        raise ZeroDivisionError(f"division by zero: {where}: the support mask has no foreground pixel "
                                "(reference: len(back_pix[0]) / len(target_pix[0]), src/test.py:174)")
    raise FloatingPointError(f"{where}: the fitted classifier is not finite (non-finite features, or the on-chip fit aborted)")


# ----------------------------------------------------------------------------------------
# (a-2) label statistics
# ----------------------------------------------------------------------------------------
def label_counts(labels: torch.Tensor, ignore_index: int = IGNORE) -> torch.Tensor:
    """labels [..., H, W] uint8/int64 -> int32 [..., 4] = (#0, #1, #ignored, #invalid) per image.
    Device-side replacement of the ``np.where`` counting in src/test.py:169-171."""
    dev = L.require_cuda(labels)
    lab = labels.contiguous()
    lead = lab.shape[:-2]
    n_img = int(torch.Size(lead).numel()) if len(lead) else 1
    npix = lab.shape[-2] * lab.shape[-1]
    counts = torch.empty(n_img, 4, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().cwt_prep_labels(L.ptr(lab), L.label_kind(lab), n_img, npix, ignore_index, None,
                                      L.ptr(counts), L.stream_ptr(dev))
    L.check(rc, "cwt_prep_labels")
    return counts.view(*lead, 4)


# ----------------------------------------------------------------------------------------
# (a-1..a-3) classifier fit
# ----------------------------------------------------------------------------------------
def fit_classifier(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float, n_iter: int,
                   class_weight: Optional[torch.Tensor] = None, ignore_index: int = IGNORE,
                   return_losses: bool = False, check: bool = True, algo: int = L.FIT_AUTO, return_status: bool = False):
    """Fit a fresh 2-class 1x1-conv classifier per episode with ``n_iter`` plain-SGD steps
    (reference: src/test.py:164-187, src/train.py:206-231, PSPNet.inner_loop pspnet.py:189-205).

    f_s      [E,S,C,h,w] (or [S,C,h,w] for one episode) float32 support features
    s_label  [E,S,H,W]   (or [S,H,W]) uint8 / int64 in {0,1,ignore_index}; H = 8(h-1)+1
    w0       [E,2,C]     (or [2,C] / [2,C,1,1]) initial weights
    class_weight [E,2] or [2] or None (None => [1, n0/n1] counted on the device, all shots pooled)
    check    True: one host sync to raise ZeroDivisionError when an episode has no foreground
             pixel (what the reference's python division does) and ValueError on labels outside
             {0,1,ignore}. False: fully asynchronous — pass ``return_status=True`` to get the deferred error
             word of every episode (int32 [E] on the device, :func:`fit_status`) and raise later with
             :func:`raise_for_status` (what ``validate_transformer`` / ``HostPipeline`` do, one batch late).

    Returns W [E,2,C] (same leading shape as given), and losses [n_iter,E] if requested."""
    dev = L.require_cuda(f_s, s_label, w0)
    f_s, lab = _episode_inputs(f_s, s_label)
    E, S, Cc, h, w = f_s.shape
    H, W = lab.shape[-2:]
    w0_shape = w0.shape
    w0v = _f32c(w0, "w0").reshape(E, 2, Cc)
    cw = _class_weight(class_weight, E, dev)
    w_out = torch.empty(E, 2, Cc, dtype=torch.float32, device=dev)
    losses = torch.empty(n_iter, E, dtype=torch.float32, device=dev) if return_losses else None
    counts = torch.empty(E, 4, dtype=torch.int32, device=dev)
    lib = L.load()
    nbytes = lib.cwt_fit_workspace_bytes(E, S, Cc, h, w, H, W)
    ws = L.WORKSPACE.get(nbytes, dev, "fit")
    with torch.cuda.device(dev):
        rc = lib.cwt_fit_classifier_f32(L.ptr(f_s), L.ptr(lab), L.label_kind(lab), L.ptr(w0v), L.ptr(cw),
                                        L.ptr(w_out), L.ptr(losses), L.ptr(counts), E, S, Cc, h, w, H, W,
                                        int(n_iter), float(lr), int(ignore_index), int(algo),
                                        L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_fit_classifier_f32")
    if check:
        _raise_on_label_counts(counts, cw is not None)
    if return_status:
        status = fit_status(counts, w_out, cw is not None)
    out = w_out.reshape(w0_shape)          # same leading shape as the w0 that was given
    res = (out,) + ((losses,) if return_losses else ()) + ((status,) if return_status else ())
    return res if len(res) > 1 else out


def fit_classifier_bias(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, b0: torch.Tensor, lr: float,
                        n_iter: int, class_weight: Optional[torch.Tensor] = None, bias_scale: float = 1.0,
                        ignore_index: int = IGNORE, return_losses: bool = False, check: bool = True):
    """:func:`fit_classifier` for a classifier with a bias (``nn.Conv2d(C, 2, 1, bias=True)``, built by the reference's
    CosCls with cls_type[2] == 'b', src/model/pspnet.py:294,319): logits = W . f + bias_scale * b, plain SGD on both.
    b0 [E,2] (or [2]). Returns (W [E,2,C], b [E,2]) with the leading shapes given (+ losses [n_iter,E] if requested)."""
    dev = L.require_cuda(f_s, s_label, w0, b0)
    f_s, lab = _episode_inputs(f_s, s_label)
    E, S, Cc, h, w = f_s.shape
    H, W = lab.shape[-2:]
    w0_shape, b0_shape = w0.shape, b0.shape
    w0v = _f32c(w0, "w0").reshape(E, 2, Cc)
    b0v = _f32c(b0, "b0").reshape(E, 2)
    cw = _class_weight(class_weight, E, dev)
    w_out = torch.empty(E, 2, Cc, dtype=torch.float32, device=dev)
    b_out = torch.empty(E, 2, dtype=torch.float32, device=dev)
    losses = torch.empty(n_iter, E, dtype=torch.float32, device=dev) if return_losses else None
    counts = torch.empty(E, 4, dtype=torch.int32, device=dev)
    lib = L.load()
    nbytes = lib.cwt_fit_bias_workspace_bytes(E, S, Cc, h, w, H, W)
    ws = L.WORKSPACE.get(nbytes, dev, "fit_bias")
    with torch.cuda.device(dev):
        rc = lib.cwt_fit_classifier_bias_f32(L.ptr(f_s), L.ptr(lab), L.label_kind(lab), L.ptr(w0v), L.ptr(b0v), L.ptr(cw),
                                             L.ptr(w_out), L.ptr(b_out), L.ptr(losses), L.ptr(counts),
                                             E, S, Cc, h, w, H, W, int(n_iter), float(lr), float(bias_scale),
                                             int(ignore_index), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_fit_classifier_bias_f32")
    if check:
        _raise_on_label_counts(counts, cw is not None)
    out = (w_out.reshape(w0_shape), b_out.reshape(b0_shape))
    return out + (losses,) if return_losses else out


def fit_coscls(x_norm: torch.Tensor, s_label: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor, lr: float,
               n_iter: int, weight_g: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None, flags: int = 0,
               class_weight: Optional[torch.Tensor] = None, ignore_index: int = IGNORE, return_losses: bool = False,
               check: bool = True):
    """The inner loop on the reference's cosine classifier ``CosCls`` with any ``cls_type`` (src/model/pspnet.py:290-323):
    scores = scale * (weight . x_norm + bias), SGD over every parameter the flags switch on.

    x_norm [E,S,C,h,w] = F.normalize(f_s, dim=1, eps=1e-5) (:func:`normalize_features`); weight [E,2,C] (``cls.weight``, or
    ``cls.weight_v`` with flag 'r'); weight_g [E,2] (flag 'r'); bias [E,2] or None; scale [E]; flags = OR of L.COSCLS_R /
    L.COSCLS_N / L.COSCLS_T. Returns a dict of the fitted parameters (new tensors; the inputs are not modified)."""
    dev = L.require_cuda(x_norm, s_label, weight, scale)
    if x_norm.dim() != 5:
        raise ValueError(f"x_norm must be [E,S,C,h,w] and s_label [E,S,H,W]; got {tuple(x_norm.shape)}, {tuple(s_label.shape)}")
    x_norm, lab = _episode_inputs(x_norm, s_label, "x_norm")
    E, S, Cc, h, w = x_norm.shape
    H, W = lab.shape[-2:]
    wv = _f32c(weight, "weight").reshape(E, 2, Cc).clone()
    sc = _f32c(scale, "scale").reshape(E).clone()
    gv = _f32c(weight_g, "weight_g").reshape(E, 2).clone() if weight_g is not None else None
    bv = _f32c(bias, "bias").reshape(E, 2).clone() if bias is not None else None
    if (flags & L.COSCLS_R) and gv is None:
        raise ValueError("flag 'r' needs weight_g")
    cw = _class_weight(class_weight, E, dev)
    losses = torch.empty(n_iter, E, dtype=torch.float32, device=dev) if return_losses else None
    counts = torch.empty(E, 4, dtype=torch.int32, device=dev)
    lib = L.load()
    nbytes = lib.cwt_fit_coscls_workspace_bytes(E, S, Cc, h, w, H, W)
    ws = L.WORKSPACE.get(nbytes, dev, "fit_coscls")
    with torch.cuda.device(dev):
        rc = lib.cwt_fit_coscls_f32(L.ptr(x_norm), L.ptr(lab), L.label_kind(lab), L.ptr(wv), L.ptr(gv), L.ptr(bv), L.ptr(sc),
                                    L.ptr(cw), L.ptr(losses), L.ptr(counts), int(flags), E, S, Cc, h, w, H, W,
                                    int(n_iter), float(lr), int(ignore_index), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_fit_coscls_f32")
    if check:
        _raise_on_label_counts(counts, cw is not None)
    out = {"weight": wv, "weight_g": gv, "bias": bv, "scale": sc}
    if return_losses:
        out["losses"] = losses
    return out


def fit_multiclass(f_s: torch.Tensor, s_label: torch.Tensor, weight: torch.Tensor, class_weight: torch.Tensor, lr: float,
                   n_iter: int, ignore_index: int = IGNORE, check: bool = True) -> torch.Tensor:
    """``n_iter`` plain-SGD steps on ONE K-class bias-free 1x1 classifier with ``CrossEntropyLoss(class_weight,
    ignore_index)`` of the logits up-sampled to the label size — the loop of ``PSPNet.increment_inner_loop``
    (src/model/pspnet.py:207-221) for K > 2. f_s [S,C,h,w], s_label [S,H,W] in [0,K) or ignore_index, weight [K,C] (or
    [K,C,1,1]), class_weight [K]. Returns the fitted weight (new tensor, same shape as given)."""
    dev = L.require_cuda(f_s, s_label, weight, class_weight)
    if f_s.dim() != 4 or s_label.dim() != 3:
        raise ValueError(f"f_s must be [S,C,h,w] and s_label [S,H,W]; got {tuple(f_s.shape)}, {tuple(s_label.shape)}")
    S, Cc, h, w = f_s.shape
    H, W = s_label.shape[-2:]
    K = weight.shape[0]
    f_s = _f32c(f_s, "f_s")
    lab = s_label.contiguous()
    wv = _f32c(weight, "weight").reshape(K, Cc).clone()
    cw = _f32c(class_weight, "class_weight").reshape(K)
    valid = lab != ignore_index
    if check and bool(((lab.long() >= K) & valid).any()):
        raise ValueError(f"s_label holds values outside [0, {K}) and ignore_index")
    # sum_i w[y_i] over the non-ignored pixels, on the device (no host sync)
    sumw = cw[lab.long().clamp(0, K - 1)][valid].sum(dtype=torch.float32)
    inv_sumw = (1.0 / sumw).reshape(1).contiguous()
    lib = L.load()
    nbytes = lib.cwt_fit_multiclass_workspace_bytes(K, S, Cc, h, w, H, W)
    ws = L.WORKSPACE.get(nbytes, dev, "fit_multiclass")
    with torch.cuda.device(dev):
        rc = lib.cwt_fit_multiclass_f32(L.ptr(f_s), L.ptr(lab), L.label_kind(lab), L.ptr(wv), L.ptr(cw), L.ptr(inv_sumw),
                                        K, S, Cc, h, w, H, W, int(n_iter), float(lr), int(ignore_index),
                                        L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_fit_multiclass_f32")
    return wv.reshape(weight.shape)


def fit_classifier_dice(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float, n_iter: int,
                        ignore_index: int = IGNORE, return_losses: bool = False, check: bool = True):
    """The inner loop of :func:`fit_classifier` with the reference's dice loss, ``SegLoss('wt_dc' | 'dc')``
    (src/model/model_util.py:18-19 -> weighted_dice_loss, :40-73): a sigmoid per logit channel, one dice term per
    (support image, channel), summed and divided by the number of support images. Same shapes as
    :func:`fit_classifier`; there is no class weight and an episode without foreground is legal (the dice term of an
    empty target is 1). ``check``: one host sync to raise ValueError on labels outside {0, 1, ignore_index}."""
    dev = L.require_cuda(f_s, s_label, w0)
    f_s, lab = _episode_inputs(f_s, s_label)
    E, S, Cc, h, w = f_s.shape
    H, W = lab.shape[-2:]
    w0_shape = w0.shape
    w0v = _f32c(w0, "w0").reshape(E, 2, Cc)
    w_out = torch.empty(E, 2, Cc, dtype=torch.float32, device=dev)
    losses = torch.empty(n_iter, E, dtype=torch.float32, device=dev) if return_losses else None
    lib = L.load()
    nbytes = lib.cwt_fit_dice_workspace_bytes(E, S, Cc, h, w, H, W)
    ws = L.WORKSPACE.get(nbytes, dev, "fit_dice")
    with torch.cuda.device(dev):
        rc = lib.cwt_fit_classifier_dice_f32(L.ptr(f_s), L.ptr(lab), L.label_kind(lab), L.ptr(w0v), L.ptr(w_out),
                                             L.ptr(losses), E, S, Cc, h, w, H, W, int(n_iter), float(lr),
                                             int(ignore_index), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_fit_classifier_dice_f32")
    if check and int(label_counts(lab.reshape(E * S, H, W), ignore_index)[:, 3].sum()) > 0:
        raise ValueError("s_label holds values outside {0, 1, ignore_index}")
    out = w_out.reshape(w0_shape)
    return (out, losses) if return_losses else out


# ----------------------------------------------------------------------------------------
# (a-5, a-6, a-13) transformer forward / backward
# ----------------------------------------------------------------------------------------
def _tdims(q, k):
    if q.dim() != 3:
        raise ValueError(f"q must be [B,Lq,C], got {tuple(q.shape)}")
    B, Lq, Cc = q.shape
    if k.dim() == 4:
        HW = k.shape[2] * k.shape[3]
    elif k.dim() == 3:
        HW = k.shape[2]
    else:
        raise ValueError(f"k must be [B,C,h,w] or [B,C,HW], got {tuple(k.shape)}")
    if k.shape[0] != B or k.shape[1] != Cc:
        raise ValueError("q and k disagree on batch or channel size")
    return B, Lq, Cc, HW


def transformer_forward(q, k, w_qkvs, fc_w, fc_b, ln_g, ln_b, n_head: int, normalize_k: bool = False,
                        keep_attn=None, keep_out=None, p_attn: float = 0.1, p_out: float = 0.5,
                        need_saved: bool = False, algo: int = L.ATTN_REASSOC):
    """MultiHeadAttentionOne.forward(q, k, k) (src/model/transformer.py:54-83), batched.
    Returns out [B,Lq,C] (and the opaque saved-activation buffer when ``need_saved``)."""
    dev = L.require_cuda(q, k, w_qkvs, fc_w, fc_b, ln_g, ln_b)
    B, Lq, Cc, HW = _tdims(q, k)
    if tuple(w_qkvs.shape) != (n_head * Cc, Cc) or tuple(fc_w.shape) != (Cc, n_head * Cc):
        raise ValueError("only d_model == d_k == d_v is supported (the reference constructs 512/512/512): "
                         f"w_qkvs {tuple(w_qkvs.shape)}, fc {tuple(fc_w.shape)}, n_head {n_head}, C {Cc}")
    q_, k_ = _f32c(q, "q"), _f32c(k, "k")
    ka = ko = None
    if keep_attn is not None:
        ka = keep_attn.to(torch.uint8).contiguous()
        if ka.numel() != n_head * B * Lq * HW:
            raise ValueError("keep_attn must be [n_head*B, Lq, HW]")
    if keep_out is not None:
        ko = keep_out.to(torch.uint8).contiguous()
        if ko.numel() != B * Lq * Cc:
            raise ValueError("keep_out must be [B, Lq, C]")
    out = torch.empty(B, Lq, Cc, dtype=torch.float32, device=dev)
    lib = L.load()
    saved = None
    if need_saved:
        saved = torch.empty(lib.cwt_transformer_saved_bytes(B, Lq, n_head, Cc, HW), dtype=torch.uint8, device=dev)
    ws = L.WORKSPACE.get(lib.cwt_transformer_workspace_bytes(B, Lq, n_head, Cc, HW, algo), dev, "tr")
    with torch.cuda.device(dev):
        rc = lib.cwt_transformer_fwd_f32(L.ptr(q_), L.ptr(k_), int(bool(normalize_k)), L.ptr(_f32c(w_qkvs, "w_qkvs")),
                                         L.ptr(_f32c(fc_w, "fc_w")), L.ptr(_f32c(fc_b, "fc_b")),
                                         L.ptr(_f32c(ln_g, "ln_g")), L.ptr(_f32c(ln_b, "ln_b")),
                                         L.ptr(ka), L.ptr(ko), float(p_attn), float(p_out), L.ptr(out), L.ptr(saved),
                                         B, Lq, n_head, Cc, HW, int(algo), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_transformer_fwd_f32")
    return (out, saved) if need_saved else out


def transformer_backward(d_out, q, k, w_qkvs, fc_w, ln_g, n_head: int, saved, normalize_k: bool = False,
                         keep_attn=None, keep_out=None, p_attn: float = 0.1, p_out: float = 0.5):
    """Gradients of the block w.r.t. (w_qkvs, fc.weight, fc.bias, ln.weight, ln.bias), summed over the batch."""
    dev = L.require_cuda(d_out, q, k, w_qkvs, fc_w, ln_g, saved)
    B, Lq, Cc, HW = _tdims(q, k)
    ka = None if keep_attn is None else keep_attn.to(torch.uint8).contiguous()
    ko = None if keep_out is None else keep_out.to(torch.uint8).contiguous()
    d_wqkvs = torch.empty_like(w_qkvs, dtype=torch.float32).contiguous()
    d_fcw = torch.empty_like(fc_w, dtype=torch.float32).contiguous()
    d_fcb = torch.empty(Cc, dtype=torch.float32, device=dev)
    d_g = torch.empty(Cc, dtype=torch.float32, device=dev)
    d_b = torch.empty(Cc, dtype=torch.float32, device=dev)
    lib = L.load()
    ws = L.WORKSPACE.get(lib.cwt_transformer_workspace_bytes(B, Lq, n_head, Cc, HW, L.ATTN_REASSOC), dev, "tr")
    with torch.cuda.device(dev):
        rc = lib.cwt_transformer_bwd_f32(L.ptr(_f32c(d_out, "d_out")), L.ptr(_f32c(q, "q")), L.ptr(_f32c(k, "k")),
                                         int(bool(normalize_k)), L.ptr(_f32c(w_qkvs, "w_qkvs")), L.ptr(_f32c(fc_w, "fc_w")),
                                         L.ptr(_f32c(ln_g, "ln_g")), L.ptr(ka), L.ptr(ko), float(p_attn), float(p_out),
                                         L.ptr(saved), L.ptr(d_wqkvs), L.ptr(d_fcw), L.ptr(d_fcb), L.ptr(d_g), L.ptr(d_b),
                                         B, Lq, n_head, Cc, HW, L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_transformer_bwd_f32")
    return d_wqkvs, d_fcw, d_fcb, d_g, d_b


# ----------------------------------------------------------------------------------------
# (a-4, a-7..a-12) logits / upsample / argmax / IoU
# ----------------------------------------------------------------------------------------
def logits_iou(weights: torch.Tensor, f_q: torch.Tensor, q_label: torch.Tensor, normalize_mask: int = 0,
               ignore_index: int = IGNORE, return_logits: bool = True):
    """weights [E,V,2,C]; f_q [E,C,h,w]; q_label [E,H,W].
    Returns counts int64 [E,V,2,3] (class x (I,U,T)), ce float64 [E,V,2] (sum of -log p[y], #valid),
    logits60 float32 [E,V,2,h,w] or None."""
    dev = L.require_cuda(weights, f_q, q_label)
    E, V, two, Cc = weights.shape
    if two != 2:
        raise ValueError("weights must be [E,V,2,C]")
    _, C2, h, w = f_q.shape
    H, W = q_label.shape[-2:]
    if C2 != Cc or f_q.shape[0] != E or q_label.shape[0] != E:
        raise ValueError("shape mismatch between weights, f_q and q_label")
    wt, fq, lab = _f32c(weights, "weights"), _f32c(f_q, "f_q"), q_label.contiguous()
    counts = torch.empty(E, V, 2, 3, dtype=torch.int64, device=dev)
    ce = torch.empty(E, V, 2, dtype=torch.float64, device=dev)
    logits = torch.empty(E, V, 2, h, w, dtype=torch.float32, device=dev) if return_logits else None
    lib = L.load()
    ws = L.WORKSPACE.get(lib.cwt_logits_iou_workspace_bytes(E, V, Cc, h, w, H, W), dev, "iou")
    with torch.cuda.device(dev):
        rc = lib.cwt_logits_iou(L.ptr(wt), L.ptr(fq), L.ptr(lab), L.label_kind(lab), int(normalize_mask), L.ptr(counts),
                                L.ptr(logits), L.ptr(ce), E, V, Cc, h, w, H, W, int(ignore_index),
                                L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_logits_iou")
    return counts, ce, logits


def upsample_argmax_iou(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = IGNORE):
    """logits [n,2,h,w] float32; target [n,H,W] -> counts int64 [n,2,3], ce float64 [n,2]."""
    dev = L.require_cuda(logits, target)
    n, two, h, w = logits.shape
    if two != 2:
        raise NotImplementedError("the fused upsample/argmax/IoU kernel handles the head's 2 classes "
                                  f"(num_classes_tr = 2); got {two}")
    H, W = target.shape[-2:]
    lg, tg = _f32c(logits, "logits"), target.contiguous()
    counts = torch.empty(n, 2, 3, dtype=torch.int64, device=dev)
    ce = torch.empty(n, 2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().cwt_upsample_argmax_iou(L.ptr(lg), L.ptr(tg), L.label_kind(tg), L.ptr(counts), L.ptr(ce),
                                              n, h, w, H, W, int(ignore_index), L.stream_ptr(dev))
    L.check(rc, "cwt_upsample_argmax_iou")
    return counts, ce


def intersection_union(preds: torch.Tensor, target: torch.Tensor, num_classes: int, ignore_index: int = IGNORE):
    """preds, target [n, ...] same shape and dtype (uint8 / int64) -> counts int64 [n, num_classes, 3] (I,U,T)."""
    dev = L.require_cuda(preds, target)
    if preds.shape != target.shape:
        raise AssertionError("preds.shape == target.shape")          # src/util.py:298
    if preds.dtype != target.dtype:
        target = target.to(preds.dtype)
    n = preds.shape[0]
    npix = preds[0].numel() if n else 0
    p, t = preds.contiguous(), target.contiguous()
    counts = torch.empty(n, num_classes, 3, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = L.load().cwt_intersection_union(L.ptr(p), L.ptr(t), L.label_kind(p), L.ptr(counts), n, npix,
                                             int(num_classes), int(ignore_index), L.stream_ptr(dev))
    L.check(rc, "cwt_intersection_union")
    return counts


def query_loss_grad(logits60: torch.Tensor, q_label: torch.Tensor, ignore_index: int = IGNORE):
    """Weighted CE at full resolution of up(logits60) with weight [1, n0/(n1+1e-12)]
    (src/train.py:237-243,262-264). logits60 [E,2,h,w]; q_label [E,H,W].
    Returns loss [E] and d loss / d logits60 [E,2,h,w]."""
    dev = L.require_cuda(logits60, q_label)
    E, two, h, w = logits60.shape
    H, W = q_label.shape[-2:]
    lg, lab = _f32c(logits60, "logits60"), q_label.contiguous()
    loss = torch.empty(E, dtype=torch.float32, device=dev)
    dl = torch.empty_like(lg)
    lib = L.load()
    ws = L.WORKSPACE.get(lib.cwt_query_loss_workspace_bytes(E, h, w, H, W), dev, "ql")
    with torch.cuda.device(dev):
        rc = lib.cwt_query_loss_grad(L.ptr(lg), L.ptr(lab), L.label_kind(lab), L.ptr(loss), L.ptr(dl), E, h, w, H, W,
                                     int(ignore_index), L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_query_loss_grad")
    return loss, dl



def rows_times_feat(M: torch.Tensor, f: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """out[e,r,p] = sum_c M[e,r,c] * fn[e,c,p] with fn = f or F.normalize(f, dim=1).
    M [E,R,C]; f [E,C,h,w] or [E,C,HW] -> [E,R,HW]  (logits = W' X^T, src/train.py:259-261)."""
    dev = L.require_cuda(M, f)
    E, R, Cc = M.shape
    HW = f[0, 0].numel()
    out = torch.empty(E, R, HW, dtype=torch.float32, device=dev)
    lib = L.load()
    ws = L.WORKSPACE.get(lib.cwt_skinny_workspace_bytes(E, R, Cc, HW), dev, "sk")
    with torch.cuda.device(dev):
        rc = lib.cwt_rows_times_feat(L.ptr(_f32c(M, "M")), L.ptr(_f32c(f, "f")), int(bool(normalize)), L.ptr(out),
                                     E, R, Cc, HW, L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_rows_times_feat")
    return out


def feat_times_rows(P: torch.Tensor, f: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """out[e,r,c] = sum_p P[e,r,p] * fn[e,c,p].  P [E,R,HW]; f [E,C,h,w] -> [E,R,C]."""
    dev = L.require_cuda(P, f)
    E, R, HW = P.shape
    Cc = f.shape[1]
    out = torch.empty(E, R, Cc, dtype=torch.float32, device=dev)
    lib = L.load()
    ws = L.WORKSPACE.get(lib.cwt_skinny_workspace_bytes(E, R, Cc, HW), dev, "sk")
    with torch.cuda.device(dev):
        rc = lib.cwt_feat_times_rows(L.ptr(_f32c(P, "P")), L.ptr(_f32c(f, "f")), int(bool(normalize)), L.ptr(out),
                                     E, R, Cc, HW, L.ptr(ws), ws.numel(), L.stream_ptr(dev))
    L.check(rc, "cwt_feat_times_rows")
    return out


def normalize_features(f: torch.Tensor, eps: float = 1e-5, scale: float = 1.0) -> torch.Tensor:
    """scale * F.normalize(f, p=2, dim=-3, eps=eps) for f [..., C, h, w]: the feature side of the cosine classifier
    (CosCls.forward, src/model/pspnet.py:302-310, with scale = scale_factor = 2.0)."""
    dev = L.require_cuda(f)
    fc = _f32c(f, "f")
    Cc, h, w = fc.shape[-3:]
    n_img = fc.numel() // max(Cc * h * w, 1)
    out = torch.empty_like(fc)
    lib = L.load()
    with torch.cuda.device(dev):
        rc = lib.cwt_normalize_features_f32(L.ptr(fc), L.ptr(out), n_img, Cc, h * w, float(eps), float(scale), L.stream_ptr(dev))
    L.check(rc, "cwt_normalize_features_f32")
    return out


# ----------------------------------------------------------------------------------------
# torch.ops.cwt_b200.* registration (PyTorch custom ops in front of the C ABI)
# ----------------------------------------------------------------------------------------
_registered = False


def _register_custom_ops() -> None:
    global _registered
    if _registered:
        return
    _registered = True
    from torch.library import custom_op

    @custom_op("cwt_b200::fit_classifier", mutates_args=(), device_types="cuda")
    def _fit(f_s: torch.Tensor, s_label: torch.Tensor, w0: torch.Tensor, lr: float, n_iter: int) -> torch.Tensor:
        return fit_classifier(f_s, s_label, w0, lr, n_iter, check=False)

    @_fit.register_fake
    def _(f_s, s_label, w0, lr, n_iter):
        return torch.empty_like(w0)

    @custom_op("cwt_b200::transformer_forward", mutates_args=(), device_types="cuda")
    def _tf(q: torch.Tensor, k: torch.Tensor, w_qkvs: torch.Tensor, fc_w: torch.Tensor, fc_b: torch.Tensor,
            ln_g: torch.Tensor, ln_b: torch.Tensor, n_head: int, normalize_k: bool) -> torch.Tensor:
        return transformer_forward(q, k, w_qkvs, fc_w, fc_b, ln_g, ln_b, n_head, normalize_k)

    @_tf.register_fake
    def _(q, k, w_qkvs, fc_w, fc_b, ln_g, ln_b, n_head, normalize_k):
        return torch.empty_like(q)

    @custom_op("cwt_b200::logits_iou", mutates_args=(), device_types="cuda")
    def _li(weights: torch.Tensor, f_q: torch.Tensor, q_label: torch.Tensor, normalize_mask: int) -> torch.Tensor:
        return logits_iou(weights, f_q, q_label, normalize_mask, return_logits=False)[0]

    @_li.register_fake
    def _(weights, f_q, q_label, normalize_mask):
        return weights.new_empty(weights.shape[0], weights.shape[1], 2, 3, dtype=torch.int64)

    @custom_op("cwt_b200::upsample_argmax_iou", mutates_args=(), device_types="cuda")
    def _ui(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return upsample_argmax_iou(logits, target)[0]

    @_ui.register_fake
    def _(logits, target):
        return logits.new_empty(logits.shape[0], 2, 3, dtype=torch.int64)
