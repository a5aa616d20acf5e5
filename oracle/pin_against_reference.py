"""ORACLE PIN — TEST INFRASTRUCTURE ONLY (run in the build container, where /root/reference exists).

    python -m oracle.pin_against_reference [--write]

1. Imports the *live reference code* (``/root/reference/src``) and checks the restatement in
   ``oracle/head_ref.py`` against it on seeded inputs:
     * ``MultiHeadAttentionOne`` (src/model/transformer.py:33-83) in eval mode, and in train
       mode with the dropout keep-masks captured from the reference's own ``nn.Dropout``
       modules by forward hooks (so the explicit-mask convention is pinned too), forward
       and parameter gradients;
     * ``batch_intersectionAndUnionGPU`` / ``intersectionAndUnionGPU`` (src/util.py:237-308).
2. Runs the episode body of ``src/test.py:162-234`` with the reference's own module objects
   doing the work (``episode_via_reference``) and the meta-training step of
   ``src/train.py:233-267``, and with ``--write`` stores the outputs as golden vectors under
   ``tests/golden/`` (inputs are regenerated from the episode seed by
   ``few_shot_seg_cwt_b200.synthetic``; an input checksum is stored beside the outputs).

The reference ships no tests or fixtures of its own, so these files *are* the pin.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("CWT_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import head_ref as O  # noqa: E402
from few_shot_seg_cwt_b200 import synthetic as syn  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> generator / hyper-parameter settings.  "full" = BASELINE.json shapes.
CASES = {
    # small shapes (CPU suite, seconds)
    "small_1shot_h1":  dict(idx=3, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", n_head=1, lr=0.0025, n_iter=200),
    "small_1shot_h4":  dict(idx=4, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", n_head=4, lr=0.1, n_iter=200),
    "small_5shot_h4":  dict(idx=5, shot=5, C=64, h=12, w=12, H=89, W=89, style="unit", n_head=4, lr=0.1, n_iter=50),
    "small_rect":      dict(idx=6, shot=2, C=128, h=9, w=14, H=65, W=105, style="backbone", n_head=2, lr=0.01, n_iter=30),
    # BASELINE.json configs[0] (pascal.yaml defaults) and the scripts/test.sh overrides
    "full_1shot_h1_yaml":    dict(idx=0, shot=1, C=512, h=60, w=60, H=473, W=473, style="unit", n_head=1, lr=0.0025, n_iter=200),
    "full_1shot_h4_script":  dict(idx=1, shot=1, C=512, h=60, w=60, H=473, W=473, style="unit", n_head=4, lr=0.1, n_iter=200),
    "full_1shot_h4_backbone": dict(idx=2, shot=1, C=512, h=60, w=60, H=473, W=473, style="backbone", n_head=4, lr=0.1, n_iter=200),
    # configs[1]: 5-shot, pooled class weight and pooled mean
    "full_5shot_h4":         dict(idx=7, shot=5, C=512, h=60, w=60, H=473, W=473, style="unit", n_head=4, lr=0.1, n_iter=200),
}
# PSPNet.inner_loop variants (src/model/pspnet.py:189-205): SegLoss type x classifier kind
INNER_CASES = {
    "inner_small_dot_ce":   dict(idx=21, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="ce", dist="dot"),
    "inner_small_cos_wtce": dict(idx=22, shot=2, C=64, h=12, w=12, H=89, W=89, style="backbone", lr=0.1, n_iter=50, loss_type="wt_ce", dist="cos"),
    "inner_small_dot_wtdc": dict(idx=24, shot=2, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="wt_dc", dist="dot"),
    "inner_small_cos_dc":   dict(idx=25, shot=1, C=64, h=12, w=12, H=89, W=89, style="backbone", lr=0.1, n_iter=50, loss_type="dc", dist="cos"),
    "inner_full_dot_dc":    dict(idx=26, shot=1, C=512, h=60, w=60, H=473, W=473, style="unit", lr=0.1, n_iter=100, loss_type="dc", dist="dot"),
    # PSPNet.increment_inner_loop (pspnet.py:207-221) on a 2-class classifier: Adapt_SegLoss(num_cls=2, fg_idx, tp)
    "inner_incr_small_tp05": dict(idx=27, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="adapt_ce", dist="dot", fg_idx=1, tp=0.5),
    "inner_incr_small_fg0":  dict(idx=28, shot=2, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="adapt_ce", dist="dot", fg_idx=0, tp=1.0),
    # classifier with a bias: CosCls cls_type 'oobo' (pspnet.py:294) and nn.Conv2d(C, 2, 1, bias=True); b0 ~ U(+-1/sqrt(C)) from seed idx
    "inner_small_cosb_wtce": dict(idx=29, shot=2, C=64, h=12, w=12, H=89, W=89, style="backbone", lr=0.1, n_iter=50, loss_type="wt_ce", dist="cos", bias=True),
    "inner_small_dotb_ce":   dict(idx=30, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="ce", dist="dot", bias=True),
    "inner_full_cosb_wtce":  dict(idx=31, shot=1, C=512, h=60, w=60, H=473, W=473, style="backbone", lr=0.1, n_iter=100, loss_type="wt_ce", dist="cos", bias=True),
    "inner_full_cos_ce":    dict(idx=23, shot=1, C=512, h=60, w=60, H=473, W=473, style="backbone", lr=0.1, n_iter=100, loss_type="ce", dist="cos"),
}
# CosCls with the remaining cls_type flags (pspnet.py:290-323): 'r' weight-norm reparametrisation, 'n' per-forward weight
# normalisation, 'b' bias, 't' learnable temperature. weight_v / weight = w0, weight_g = [0.9, 1.2] * row norms of w0,
# bias = O.initial_bias(idx, C), scale_factor = 2.0.
COSCLS_CASES = {
    "coscls_small_onoo": dict(idx=32, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="wt_ce", cls_type="onoo"),
    "coscls_small_ooot": dict(idx=33, shot=2, C=64, h=12, w=12, H=89, W=89, style="backbone", lr=0.1, n_iter=50, loss_type="wt_ce", cls_type="ooot"),
    "coscls_small_rooo": dict(idx=34, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="ce", cls_type="rooo"),
    "coscls_small_robt": dict(idx=35, shot=1, C=64, h=12, w=12, H=89, W=89, style="backbone", lr=0.1, n_iter=50, loss_type="wt_ce", cls_type="robt"),
    "coscls_small_onbt": dict(idx=36, shot=2, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=50, loss_type="wt_ce", cls_type="onbt"),
    "coscls_full_rnbt":  dict(idx=37, shot=1, C=512, h=60, w=60, H=473, W=473, style="backbone", lr=0.1, n_iter=100, loss_type="wt_ce", cls_type="rnbt"),
}

# PSPNet.increment_inner_loop with more than two classes (pspnet.py:207-221; train_cca.py's multi-way setting)
INCRMC_CASES = {
    "incrmc_small_k5":  dict(idx=38, shot=2, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=40, K=5, fg_idx=3, tp=1.1),
    # ("backbone"-scale features at lr 0.1 make the 17-way trajectory chaotic — a 2e-7 difference after one step grows
    #  tenfold every five steps on ANY implementation, measured — so the multi-way vectors use the O(1)-scale features)
    "incrmc_small_k17": dict(idx=39, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", lr=0.1, n_iter=40, K=17, fg_idx=16, tp=1.0),
    "incrmc_full_k16":  dict(idx=40, shot=1, C=512, h=60, w=60, H=473, W=473, style="unit", lr=0.1, n_iter=30, K=16, fg_idx=7, tp=1.1),
}

# validation transform (src/dataset/transform.py Resize -> ToTensor -> Normalize): source sizes x target size
TRANSFORM_CASES = {
    "transform_land_97":   dict(idx=1, h=75, w=120, size=97, padding=None),
    "transform_port_97":   dict(idx=2, h=131, w=88, size=97, padding="avg"),
    "transform_up_97":     dict(idx=3, h=40, w=52, size=97, padding=None),
    "transform_half_96":   dict(idx=4, h=192, w=192, size=96, padding=None),        # exact 2x down-scale (cv2's area fast path)
}
PASCAL_MEAN, PASCAL_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]               # config_files/pascal.yaml:15-16

TRAIN_CASES = {
    "train_small_h2": dict(idx=11, shot=1, C=64, h=12, w=12, H=89, W=89, style="unit", n_head=2, lr=0.1, n_iter=20, p_attn=0.1, p_out=0.5),
    "train_full_h1":  dict(idx=12, shot=1, C=512, h=60, w=60, H=473, W=473, style="unit", n_head=1, lr=0.1, n_iter=20, p_attn=0.1, p_out=0.5),
}


def _import_reference():
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.transformer import MultiHeadAttentionOne
        from src.util import batch_intersectionAndUnionGPU, intersectionAndUnionGPU
    return MultiHeadAttentionOne, batch_intersectionAndUnionGPU, intersectionAndUnionGPU


def input_checksum(ep: syn.Episode) -> str:
    hsh = hashlib.sha256()
    for t in (ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0):
        hsh.update(t.contiguous().numpy().tobytes())
    return hsh.hexdigest()


def gen_kwargs(case):
    return {k: case[k] for k in ("shot", "C", "h", "w", "H", "W", "style")}


from oracle.ref_episode import build_reference_transformer, episode_via_reference  # noqa: E402  (shared with bench.py's CPU legs)


def inner_loop_via_reference(ep, case):
    """src/model/pspnet.py:189-205 with the reference's own ``CosCls`` / ``nn.Conv2d`` classifier (get_classifier,
    pspnet.py:326-334) and its own ``SegLoss`` (model_util.py:9-24) doing the work."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.pspnet import CosCls
        from src.model.model_util import SegLoss, Adapt_SegLoss
    C = ep.f_s.shape[1]
    with_bias = bool(case.get("bias"))
    if case["dist"] == "dot":
        classifier = nn.Conv2d(C, 2, kernel_size=1, bias=with_bias)
        conv = classifier
    else:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):          # parse_param_coscls prints its flags
            classifier = CosCls(in_dim=C, n_classes=2, cls_type="oobo" if with_bias else "oooo")
        conv = classifier.cls
    with torch.no_grad():
        conv.weight.copy_(ep.w0.reshape(2, C, 1, 1))
        if with_bias:
            conv.bias.copy_(O.initial_bias(case["idx"], C))
    optimizer = torch.optim.SGD(classifier.parameters(), lr=case["lr"])
    if case["loss_type"] == "adapt_ce":                           # increment_inner_loop's criterion (pspnet.py:213)
        criterion = Adapt_SegLoss(num_cls=2, fg_idx=case["fg_idx"], tp=case["tp"])
    else:
        criterion = SegLoss(loss_type=case["loss_type"])
    f_s, s_label = ep.f_s, ep.s_label.long()
    cuda_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False                       # weighted_ce_loss moves its weight with .cuda() when it can
    try:
        for _ in range(case["n_iter"]):
            pred_s_label = classifier(f_s)
            pred_s_label = F.interpolate(pred_s_label, size=s_label.size()[1:], mode="bilinear", align_corners=True)
            s_loss = criterion(pred_s_label, s_label)
            optimizer.zero_grad()
            s_loss.backward()
            optimizer.step()
    finally:
        torch.cuda.is_available = cuda_avail
    if with_bias:
        return (conv.weight.detach().reshape(2, C).clone(), conv.bias.detach().clone()), float(s_loss)
    return conv.weight.detach().reshape(2, C).clone(), float(s_loss)


def coscls_g0(w0):
    return w0.reshape(2, -1).norm(dim=1) * torch.tensor([0.9, 1.2])


def coscls_via_reference(ep, case):
    """inner_loop (pspnet.py:189-205) with the reference's own CosCls(cls_type) and SegLoss objects."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.pspnet import CosCls
        from src.model.model_util import SegLoss
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            classifier = CosCls(in_dim=ep.f_s.shape[1], n_classes=2, cls_type=case["cls_type"])
    C = ep.f_s.shape[1]
    ct = case["cls_type"]
    with torch.no_grad():
        if ct[0] == "r":
            classifier.cls.weight_v.copy_(ep.w0.reshape(2, C, 1, 1))
            classifier.cls.weight_g.copy_(coscls_g0(ep.w0).reshape(2, 1, 1, 1))
        else:
            classifier.cls.weight.copy_(ep.w0.reshape(2, C, 1, 1))
        if ct[2] == "b":
            classifier.cls.bias.copy_(O.initial_bias(case["idx"], C))
    optimizer = torch.optim.SGD(classifier.parameters(), lr=case["lr"])
    criterion = SegLoss(loss_type=case["loss_type"])
    f_s, s_label = ep.f_s, ep.s_label.long()
    cuda_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(case["n_iter"]):
                pred = classifier(f_s)
                pred = F.interpolate(pred, size=s_label.size()[1:], mode="bilinear", align_corners=True)
                s_loss = criterion(pred, s_label)
                optimizer.zero_grad()
                s_loss.backward()
                optimizer.step()
    finally:
        torch.cuda.is_available = cuda_avail
    cls = classifier.cls
    return {"weight": (cls.weight_v if ct[0] == "r" else cls.weight).detach().reshape(2, C).clone(),
            "weight_g": cls.weight_g.detach().reshape(2).clone() if ct[0] == "r" else None,
            "bias": cls.bias.detach().clone() if ct[2] == "b" else None,
            "scale": torch.as_tensor(classifier.scale_factor).detach().clone().float()}, float(s_loss)


def incr_multiclass_via_reference(ep, case):
    """The loop of PSPNet.increment_inner_loop (pspnet.py:207-221) with the reference's own Adapt_SegLoss doing the work."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.model_util import Adapt_SegLoss
    K, C = case["K"], ep.f_s.shape[1]
    classifier = nn.Conv2d(C, K, kernel_size=1, bias=False)
    with torch.no_grad():
        classifier.weight.copy_(O.multiclass_w0(case["idx"], K, C).reshape(K, C, 1, 1))
    optimizer = torch.optim.SGD(classifier.parameters(), lr=case["lr"])
    criterion = Adapt_SegLoss(num_cls=K, fg_idx=case["fg_idx"], tp=case["tp"])
    f_s, s_label = ep.f_s, O.multiclass_labels(ep.s_label, K, case["fg_idx"])
    cuda_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        for _ in range(case["n_iter"]):
            pred_s_label = classifier(f_s)
            pred_s_label = F.interpolate(pred_s_label, size=s_label.size()[1:], mode="bilinear", align_corners=True)
            s_loss = criterion(pred_s_label, s_label)
            optimizer.zero_grad()
            s_loss.backward()
            optimizer.step()
    finally:
        torch.cuda.is_available = cuda_avail
    return classifier.weight.detach().reshape(K, C).clone(), float(s_loss)


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_transformer(MHA):
    """Eval + train-mode (captured masks) forward and parameter grads vs the reference module."""
    worst = 0.0
    for n_head, C, hw in ((1, 64, (12, 12)), (4, 64, (12, 12)), (2, 128, (9, 14))):
        params = syn.make_transformer_params(n_head, C)
        g = torch.Generator().manual_seed(n_head * 100 + C)
        q = torch.randn(1, 2, C, generator=g) * 0.3
        k = F.normalize(torch.relu(torch.randn(1, C, *hw, generator=g)), dim=1)
        m = build_reference_transformer(MHA, params, n_head, C, dropout=0.5)
        m.eval()
        with torch.no_grad():
            y_ref = m(q, k, k)
        y = O.mha_one_forward_ref(q, k, params, n_head)
        worst = max(worst, relerr(y, y_ref))
        # train mode: capture the keep masks the reference drew
        m.train()
        cap = {}
        h1 = m.attention.dropout.register_forward_hook(
            lambda mod, inp, out: cap.__setitem__("attn", ((out != 0) | (inp[0] == 0))))
        h2 = m.dropout.register_forward_hook(
            lambda mod, inp, out: cap.__setitem__("out", ((out != 0) | (inp[0] == 0))))
        torch.manual_seed(1234 + n_head)
        y_ref = m(q, k, k)
        tgt = torch.randn(y_ref.shape, generator=g)
        (y_ref * tgt).sum().backward()
        h1.remove(); h2.remove()
        pd = {kk: v.clone().requires_grad_(True) for kk, v in params.items()}
        y = O.mha_one_forward_ref(q, k, pd, n_head, cap["attn"], cap["out"],
                                  p_attn=m.attention.dropout.p, p_out=m.dropout.p)
        (y * tgt).sum().backward()
        worst = max(worst, relerr(y.detach(), y_ref.detach()))
        for name, prm in m.named_parameters():
            worst = max(worst, relerr(pd[name].grad, prm.grad))
    return worst


def check_iou(batch_iou, single_iou):
    g = torch.Generator().manual_seed(7)
    ok = True
    for (h, w, H, W) in ((12, 12, 89, 89), (60, 60, 473, 473), (9, 14, 65, 105)):
        logits = torch.randn(3, 1, 2, h, w, generator=g)
        tgt = torch.randint(0, 3, (3, 1, H, W), generator=g)
        tgt[tgt == 2] = 255
        a = batch_iou(logits, tgt.clone(), 2)
        b = O.batch_intersection_and_union_ref(logits, tgt, 2)
        ok &= all(torch.equal(x, y) for x, y in zip(a, b))
        p = torch.randint(0, 2, (H, W), generator=g)
        a = single_iou(p.clone(), tgt[0, 0].clone(), 2)
        b = O.intersection_and_union_ref(p, tgt[0, 0], 2)
        ok &= all(torch.equal(x, y) for x, y in zip(a, b))
    return ok


def train_step_via_reference(ep, params, case, MHA):
    """src/train.py:233-267 with the reference module in train mode; returns outputs and the
    keep masks it drew (so CUDA / oracle can be fed the same masks)."""
    C, h, w = ep.f_q.shape
    H, W = ep.q_label.shape
    n_head = case["n_head"]
    w_fit = O.fit_classifier_ref(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"])
    m = build_reference_transformer(MHA, params, n_head, C, dropout=case["p_out"]).train()
    cap = {}
    h1 = m.attention.dropout.register_forward_hook(
        lambda mod, inp, out: cap.__setitem__("attn", ((out != 0) | (inp[0] == 0))))
    h2 = m.dropout.register_forward_hook(
        lambda mod, inp, out: cap.__setitem__("out", ((out != 0) | (inp[0] == 0))))
    torch.manual_seed(2021 + ep.idx)
    q_label = ep.q_label.long().unsqueeze(0)
    arr = q_label.numpy().copy()
    q_back, q_tgt = np.where(arr == 0), np.where(arr == 1)
    criterion = nn.CrossEntropyLoss(weight=torch.tensor([1.0, len(q_back[0]) / (len(q_tgt[0]) + 1e-12)]),
                                    ignore_index=255)
    with torch.no_grad():
        f_q = F.normalize(ep.f_q.unsqueeze(0), dim=1)
    wr = w_fit.view(2, C, 1, 1).squeeze().unsqueeze(0).expand(1, 2, C)
    updated = m(wr, f_q, f_q)
    pred_q = torch.matmul(updated, f_q.view(1, C, -1)).view(1, 2, h, w)
    pred_q = F.interpolate(pred_q, size=q_label.shape[1:], mode="bilinear", align_corners=True)
    loss_q = criterion(pred_q, q_label)
    loss_q.backward()
    h1.remove(); h2.remove()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    return {"W_fit": w_fit, "W_adapted": updated.detach()[0], "loss": loss_q.detach(),
            "grads": grads, "keep_attn": cap["attn"].to(torch.uint8), "keep_out": cap["out"].to(torch.uint8)}


SUB = 5  # stride of the stored sub-sample of large gradient tensors


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write", action="store_true", help="write tests/golden/*.npz")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    MHA, batch_iou, single_iou = _import_reference()
    report = {"torch": torch.__version__, "numpy": np.__version__, "cases": {}}

    e = check_transformer(MHA)
    print(f"[pin] MultiHeadAttentionOne restatement vs reference module (fwd+grads, eval+train masks): max rel {e:.2e}")
    assert e < 2e-6, e
    ok = check_iou(batch_iou, single_iou)
    print(f"[pin] intersection/union restatement vs reference functions: bit-exact={ok}")
    assert ok
    report["transformer_max_rel"] = e

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, case in CASES.items():
        if a.only and a.only not in name:
            continue
        ep = syn.make_episode(case["idx"], **gen_kwargs(case))
        params = syn.make_transformer_params(case["n_head"], case["C"])
        ref = episode_via_reference(ep, params, case["n_head"], case["lr"], case["n_iter"], MHA, batch_iou)
        ora = O.episode_ref(ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0, params,
                            case["n_head"], case["lr"], case["n_iter"])
        errs = {k: relerr(ora[k], ref[k]) for k in ("W_fit", "W_adapted", "logits60", "logits60_0")}
        cnt_ok = bool(torch.equal(ora["counts"], ref["counts"]) and torch.equal(ora["counts0"], ref["counts0"]))
        print(f"[pin] {name}: oracle vs reference rel {errs} counts_equal={cnt_ok} "
              f"loss {float(ref['loss']):.6f} counts {ref['counts'].tolist()}")
        assert max(errs.values()) < 1e-6 and cnt_ok
        report["cases"][name] = {"errs": errs, "loss": float(ref["loss"])}
        if a.write:
            tm, tm0 = ref["tie_margin"], ref["tie_margin0"]
            np.savez_compressed(
                os.path.join(GOLDEN_DIR, name + ".npz"),
                case=json.dumps(case), checksum=input_checksum(ep), torch_version=torch.__version__,
                W_fit=ref["W_fit"].numpy(), W_adapted=ref["W_adapted"].numpy(),
                logits60=ref["logits60"].numpy(), logits60_0=ref["logits60_0"].numpy(),
                counts=ref["counts"].numpy(), counts0=ref["counts0"].numpy(), loss=float(ref["loss"]),
                # size of the stated argmax-tie set at a few thresholds (adapted / baseline)
                tie_thresholds=np.array([1e-6, 1e-5, 1e-4]),
                tie_counts=np.array([[int((tm <= t).sum()) for t in (1e-6, 1e-5, 1e-4)],
                                     [int((tm0 <= t).sum()) for t in (1e-6, 1e-5, 1e-4)]]))

    for name, case in INNER_CASES.items():
        if a.only and a.only not in name:
            continue
        ep = syn.make_episode(case["idx"], **gen_kwargs(case))
        w_ref, last_loss = inner_loop_via_reference(ep, case)
        b0 = O.initial_bias(case["idx"], case["C"]) if case.get("bias") else None
        w_ora = O.inner_loop_ref(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"], case["loss_type"], case["dist"],
                                 case.get("fg_idx", 1), case.get("tp", 1.0), b0)
        extra = {}
        if b0 is not None:
            (w_ref, b_ref), (w_ora, b_ora) = w_ref, w_ora
            assert relerr(b_ora, b_ref) < 1e-6
            extra = dict(b_fit=b_ref.numpy(), b0=b0.numpy())
        err = relerr(w_ora, w_ref)
        print(f"[pin] {name}: oracle vs reference inner_loop rel {err:.2e} last loss {last_loss:.6f}")
        assert err < 1e-6, err
        report["cases"][name] = {"errs": {"W_fit": err}, "loss": last_loss}
        if a.write:
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), case=json.dumps(case), checksum=input_checksum(ep),
                                torch_version=torch.__version__, W_fit=w_ref.numpy(), loss=last_loss, **extra)

    for name, case in COSCLS_CASES.items():
        if a.only and a.only not in name:
            continue
        ep = syn.make_episode(case["idx"], **gen_kwargs(case))
        ref, last_loss = coscls_via_reference(ep, case)
        ct = case["cls_type"]
        ora = O.coscls_inner_loop_ref(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"], ct, case["loss_type"],
                                      O.initial_bias(case["idx"], case["C"]) if ct[2] == "b" else None,
                                      coscls_g0(ep.w0) if ct[0] == "r" else None)
        errs = {k: relerr(ora[k], ref[k]) for k in ref if ref[k] is not None}
        print(f"[pin] {name}: oracle vs reference CosCls('{ct}') inner_loop " + " ".join(f"{k} {v:.1e}" for k, v in errs.items())
              + f" last loss {last_loss:.6f} scale {float(ref['scale']):.4f}")
        assert max(errs.values()) < 1e-6, errs
        report["cases"][name] = {"errs": errs, "loss": last_loss}
        if a.write:
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), case=json.dumps(case), checksum=input_checksum(ep),
                                torch_version=torch.__version__, loss=last_loss, g0=coscls_g0(ep.w0).numpy(),
                                **{k: v.numpy() for k, v in ref.items() if v is not None})

    for name, case in INCRMC_CASES.items():
        if a.only and a.only not in name:
            continue
        ep = syn.make_episode(case["idx"], **gen_kwargs(case))
        w_ref, last_loss = incr_multiclass_via_reference(ep, case)
        w_ora = O.increment_inner_loop_ref(ep.f_s, O.multiclass_labels(ep.s_label, case["K"], case["fg_idx"]),
                                           O.multiclass_w0(case["idx"], case["K"], case["C"]), case["lr"], case["n_iter"],
                                           case["fg_idx"], case["tp"])
        err = relerr(w_ora, w_ref)
        print(f"[pin] {name}: oracle vs reference increment_inner_loop (K={case['K']}) rel {err:.2e} last loss {last_loss:.6f}")
        assert err < 1e-6, err
        report["cases"][name] = {"errs": {"W_fit": err}, "loss": last_loss}
        if a.write:
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), case=json.dumps(case), checksum=input_checksum(ep),
                                torch_version=torch.__version__, W_fit=w_ref.numpy(), loss=last_loss)

    for name, case in TRANSFORM_CASES.items():
        if a.only and a.only not in name:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from src.dataset import transform as RT
        img, lab = O.synthetic_image(case["idx"], case["h"], case["w"])
        padding = [v * 255 for v in PASCAL_MEAN] if case["padding"] == "avg" else None                    # dataset.py:82
        compose = RT.Compose([RT.Resize(case["size"], padding=padding), RT.ToTensor(),
                              RT.Normalize(mean=PASCAL_MEAN, std=PASCAL_STD)])                            # dataset.py:78-84
        ref_img, ref_lab = compose(img.copy(), lab.copy())
        ora_img, ora_lab = O.val_transform_ref(img, lab, case["size"], PASCAL_MEAN, PASCAL_STD, padding)
        err = float((ora_img - ref_img).abs().max())
        print(f"[pin] {name}: oracle vs reference Compose(Resize, ToTensor, Normalize) max abs {err:.1e} labels equal {bool(torch.equal(ora_lab, ref_lab))}")
        assert err == 0.0 and torch.equal(ora_lab, ref_lab)
        report["cases"][name] = {"errs": {"image": err}}
        if a.write:
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), case=json.dumps(case), torch_version=torch.__version__,
                                image=ref_img.numpy().astype(np.float32), label=ref_lab.numpy().astype(np.uint8))

    for name, case in TRAIN_CASES.items():
        if a.only and a.only not in name:
            continue
        ep = syn.make_episode(case["idx"], **gen_kwargs(case))
        params = syn.make_transformer_params(case["n_head"], case["C"])
        ref = train_step_via_reference(ep, params, case, MHA)
        ora = O.meta_train_step_ref(ref["W_fit"], ep.f_q, ep.q_label, params, case["n_head"],
                                    ref["keep_attn"], ref["keep_out"], case["p_attn"], case["p_out"])
        errs = {k: relerr(ora["grads"][k], ref["grads"][k]) for k in ref["grads"] if k != "layer_norm.bias"}
        errs["loss"] = abs(float(ora["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
        errs["W_adapted"] = relerr(ora["W_adapted"], ref["W_adapted"])
        print(f"[pin] {name}: oracle vs reference train step rel {errs}")
        assert max(errs.values()) < 5e-5, errs
        report["cases"][name] = {"errs": errs, "loss": float(ref["loss"])}
        if a.write:
            out = dict(case=json.dumps(case), checksum=input_checksum(ep), torch_version=torch.__version__,
                       W_fit=ref["W_fit"].numpy(), W_adapted=ref["W_adapted"].numpy(), loss=float(ref["loss"]),
                       keep_attn=np.packbits(ref["keep_attn"].numpy().reshape(-1)),
                       keep_attn_shape=np.array(ref["keep_attn"].shape),
                       keep_out=ref["keep_out"].numpy(), sub=SUB)
            for k, gten in ref["grads"].items():
                flat = gten.reshape(-1)
                out["grad_" + k] = (flat if flat.numel() <= 70000 else flat[::SUB]).numpy()
                out["gradnorm_" + k] = float(gten.double().norm())
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)

    if a.write:
        rp = os.path.join(GOLDEN_DIR, "PIN_REPORT.json")
        if a.only and os.path.exists(rp):                        # partial run: keep the other cases of the stored report
            old = json.load(open(rp))
            old["cases"].update(report["cases"])
            report["cases"] = old["cases"]
        with open(rp, "w") as f:
            json.dump(report, f, indent=1)
    print("[pin] done")


if __name__ == "__main__":
    main()
