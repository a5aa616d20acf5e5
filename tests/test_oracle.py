"""CPU suite, part 1: the oracle against the golden vectors recorded from the live reference
(tests/golden/*.npz, written by oracle/pin_against_reference.py), and against an independent
closed-form restatement."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import gen_kwargs, golden_names, load_golden, rel_err
from few_shot_seg_cwt_b200 import synthetic as syn
from oracle import head_ref as O


def _checksum(ep):
    h = hashlib.sha256()
    for t in (ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0):
        h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


EVAL_CASES = [n for n in golden_names() if not n.startswith(("train_", "inner_", "coscls_", "incrmc_", "transform_"))]
FAST = [n for n in EVAL_CASES if n.startswith("small")] + ["full_1shot_h1_yaml"]


@pytest.mark.parametrize("name", FAST)
def test_oracle_reproduces_reference_golden(name):
    g = load_golden(name)
    case = g["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    assert _checksum(ep) == str(g["checksum"]), "synthetic generator drifted from the pinned inputs"
    params = syn.make_transformer_params(case["n_head"], case["C"])
    out = O.episode_ref(ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0, params, case["n_head"], case["lr"], case["n_iter"])
    for k in ("W_fit", "W_adapted", "logits60", "logits60_0"):
        assert rel_err(out[k], g[k]) < 1e-6, k
    assert np.array_equal(out["counts"].numpy(), g["counts"])
    assert np.array_equal(out["counts0"].numpy(), g["counts0"])
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-6


def _dice_fit_closed_form(f_s, s_label, w0, lr, n_iter):
    """weighted_dice_loss (model_util.py:40-73) + SGD without autograd, in float64: explicit bilinear matrices and the
    hand-derived gradient dl/du = (-2 t / D + 4 N p / D^2) p (1 - p) / S the CUDA kernels implement."""
    S, C, h, w = f_s.shape
    H, W = s_label.shape[-2:]
    By, Bx = O.bilinear_matrix(h, H), O.bilinear_matrix(w, W)
    Fm = f_s.double().reshape(S, C, h * w)
    t = torch.stack([s_label == 0, s_label == 1], dim=1).double()                     # [S,2,H,W]
    Wt = w0.double().clone()
    for _ in range(n_iter):
        lg = torch.einsum("rc,scp->srp", Wt, Fm).reshape(S, 2, h, w)
        p = torch.sigmoid(torch.einsum("Yy,sryx,Xx->srYX", By, lg, Bx))
        N, D = (t * p).sum((-1, -2), keepdim=True), ((p * p).sum((-1, -2), keepdim=True) + t.sum((-1, -2), keepdim=True)).clamp_min(1e-8)
        du = (-2 * t / D + 4 * N * p / D ** 2) * p * (1 - p) / S
        g60 = torch.einsum("Yy,srYX,Xx->sryx", By, du, Bx).reshape(S, 2, h * w)
        Wt -= lr * torch.einsum("srp,scp->rc", g60, Fm)
    return Wt.float()


@pytest.mark.parametrize("name", ["inner_small_dot_ce", "inner_small_cos_wtce", "inner_small_dot_wtdc", "inner_small_cos_dc",
                                  "inner_incr_small_tp05", "inner_incr_small_fg0", "inner_small_cosb_wtce", "inner_small_dotb_ce"])
def test_oracle_inner_loop_variants_reproduce_reference_golden(name):
    """PSPNet.inner_loop switches (SegLoss 'ce' / 'wt_ce', dot / cosine classifier) against the vectors recorded from the
    live reference's CosCls and SegLoss objects; also: 'ce' is the weighted fit with class weight [1, 1], and the cosine
    classifier is the dot classifier on 2 * F.normalize(f, eps=1e-5) — the two identities the CUDA path relies on."""
    g = load_golden(name)
    case = g["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    assert _checksum(ep) == str(g["checksum"])
    b0 = O.initial_bias(case["idx"], case["C"]) if case.get("bias") else None
    w = O.inner_loop_ref(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"], case["loss_type"], case["dist"],
                         case.get("fg_idx", 1), case.get("tp", 1.0), b0)
    if b0 is not None:                                        # classifier with a bias (CosCls 'oobo'): weight and bias
        assert rel_err(w[0], g["W_fit"]) < 1e-6 and rel_err(w[1], g["b_fit"]) < 1e-6
        return
    assert rel_err(w, g["W_fit"]) < 1e-6
    if case["loss_type"] in ("wt_dc", "dc"):                  # dice: no CE identity to check; an independent restatement instead
        feat = 2.0 * torch.nn.functional.normalize(ep.f_s, p=2, dim=1, eps=1e-5) if case["dist"] == "cos" else ep.f_s
        assert rel_err(_dice_fit_closed_form(feat, ep.s_label, ep.w0, case["lr"], case["n_iter"]), g["W_fit"]) < 2e-5
        return
    feat = 2.0 * torch.nn.functional.normalize(ep.f_s, p=2, dim=1, eps=1e-5) if case["dist"] == "cos" else ep.f_s
    cw = torch.ones(2) if case["loss_type"] == "ce" else None
    if case["loss_type"] == "adapt_ce":                       # increment_inner_loop == the weighted fit with weight[fg] = (bg/fg)^tp
        n = [int((ep.s_label == k).sum()) for k in (0, 1)]
        cw = torch.ones(2)
        cw[case["fg_idx"]] = (torch.tensor(n[1 - case["fg_idx"]]) / torch.tensor(n[case["fg_idx"]])) ** case["tp"]
    w2 = O.fit_classifier_ref(feat, ep.s_label, ep.w0, case["lr"], case["n_iter"], class_weight=cw)
    assert rel_err(w2, g["W_fit"]) < 2e-6


@pytest.mark.parametrize("name", ["small_1shot_h4", "small_5shot_h4", "small_rect"])
def test_closed_form_fit_matches_literal_loop(name):
    """The explicit-matrix restatement (fp64) agrees with the autograd loop (fp32) far inside the
    1e-4 tolerance of the north-star: this bounds fp32 drift of the 200-step recursion."""
    case = load_golden(name)["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    w_lit = O.fit_classifier_ref(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"])
    w_cf = O.fit_classifier_closed_form(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"])
    assert rel_err(w_lit, w_cf) < 2e-5
    # the two softmax-gradient rows cancel: W0 + W1 is invariant under the fit
    assert torch.allclose((w_cf[0] + w_cf[1]).float(), ep.w0[0] + ep.w0[1], atol=1e-5)


def test_train_step_golden_small():
    g = load_golden("train_small_h2")
    case = g["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    params = syn.make_transformer_params(case["n_head"], case["C"])
    ka = torch.from_numpy(np.unpackbits(g["keep_attn"])[: int(np.prod(g["keep_attn_shape"]))].reshape(tuple(g["keep_attn_shape"])))
    ko = torch.from_numpy(g["keep_out"])
    out = O.meta_train_step_ref(torch.from_numpy(g["W_fit"]), ep.f_q, ep.q_label, params, case["n_head"], ka, ko,
                                case["p_attn"], case["p_out"])
    assert abs(float(out["loss"]) - float(g["loss"])) < 1e-5
    assert rel_err(out["W_adapted"], g["W_adapted"]) < 1e-5
    for k, v in out["grads"].items():
        ref = g["grad_" + k]
        flat = v.reshape(-1)
        got = flat if flat.numel() == ref.size else flat[:: int(g["sub"])]
        if k == "layer_norm.bias":      # analytically zero: absolute tolerance (SURVEY.md §8 math block)
            assert float(got.abs().max()) < 1e-5
        else:
            assert rel_err(got, ref) < 5e-5, k


def test_iou_edge_cases():
    # all ignored -> all zero; ties -> class 0
    tgt = torch.full((1, 1, 17, 17), 255)
    lg = torch.zeros(1, 1, 2, 3, 3)
    I, U, T = O.batch_intersection_and_union_ref(lg, tgt, 2)
    assert I.sum() == 0 and U.sum() == 0 and T.sum() == 0
    tgt = torch.zeros(1, 1, 17, 17, dtype=torch.long)
    I, U, T = O.batch_intersection_and_union_ref(lg, tgt, 2)      # exact tie everywhere -> pred 0
    assert I[0, 0].tolist() == [289.0, 0.0] and U[0, 0].tolist() == [289.0, 0.0]


def test_class_weight_zero_foreground_raises():
    with pytest.raises(ZeroDivisionError):
        O.class_weight_ref(torch.zeros(1, 9, 9, dtype=torch.long))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present (GPU box)")
def test_pin_against_live_reference_small():
    """Where the reference is mounted: the restatement still matches the reference's own objects."""
    from oracle import pin_against_reference as P
    MHA, biou, siou = P._import_reference()
    assert P.check_transformer(MHA) < 2e-6
    assert P.check_iou(biou, siou)
    case = P.CASES["small_1shot_h4"]
    ep = syn.make_episode(case["idx"], **P.gen_kwargs(case))
    params = syn.make_transformer_params(case["n_head"], case["C"])
    ref = P.episode_via_reference(ep, params, case["n_head"], case["lr"], case["n_iter"], MHA, biou)
    g = load_golden("small_1shot_h4")
    assert rel_err(ref["W_adapted"], g["W_adapted"]) < 1e-6
    assert np.array_equal(ref["counts"].numpy(), g["counts"])


def test_oracle_ref_copy_matches_restatement_and_golden():
    """oracle/_ref (the reference's own transformer.py / util.py, copied verbatim by oracle/make_ref.py; what bench.py's CPU
    legs run on the GPU box): the episode body around those modules equals the restatement and the golden vector."""
    from oracle import ref_episode as R
    mods = R.load_reference_modules(prefer_live=False)
    if mods is None:
        pytest.skip("oracle/_ref has not been made (python oracle/make_ref.py needs the reference checkout)")
    MHA, biou, _siou, where = mods
    assert where == "oracle/_ref"
    g = load_golden("small_1shot_h4")
    case = g["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    params = syn.make_transformer_params(case["n_head"], case["C"])
    ref = R.episode_via_reference(ep, params, case["n_head"], case["lr"], case["n_iter"], MHA, biou)
    ora = O.episode_ref(ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0, params, case["n_head"], case["lr"], case["n_iter"])
    for k in ("W_fit", "W_adapted", "logits60", "logits60_0"):
        assert rel_err(ora[k], ref[k]) < 1e-6, k
    assert torch.equal(ora["counts"], ref["counts"]) and torch.equal(ora["counts0"], ref["counts0"])
    assert rel_err(ref["W_adapted"], g["W_adapted"]) < 1e-6 and np.array_equal(ref["counts"].numpy(), g["counts"])


@pytest.mark.parametrize("name", ["coscls_small_onoo", "coscls_small_robt", "coscls_small_onbt"])
def test_oracle_coscls_flag_variants_reproduce_reference_golden(name):
    """CosCls with cls_type flags 'r' / 'n' / 'b' / 't' (pspnet.py:290-323): the explicit-parameter restatement against the
    parameters recorded from the live reference's CosCls object after inner_loop."""
    g = load_golden(name)
    case = g["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    assert _checksum(ep) == str(g["checksum"])
    ct = case["cls_type"]
    out = O.coscls_inner_loop_ref(ep.f_s, ep.s_label, ep.w0, case["lr"], case["n_iter"], ct, case["loss_type"],
                                  O.initial_bias(case["idx"], case["C"]) if ct[2] == "b" else None,
                                  torch.from_numpy(g["g0"]) if ct[0] == "r" else None)
    for k in ("weight", "weight_g", "bias", "scale"):
        if out[k] is not None:
            assert rel_err(out[k], g[k]) < 1e-6


@pytest.mark.parametrize("name", ["incrmc_small_k5", "incrmc_small_k17"])
def test_oracle_increment_inner_loop_multiclass_reproduces_reference_golden(name):
    """PSPNet.increment_inner_loop with K > 2 (pspnet.py:207-221) against weights recorded with the reference's Adapt_SegLoss."""
    g = load_golden(name)
    case = g["case"]
    ep = syn.make_episode(case["idx"], **gen_kwargs(case))
    assert _checksum(ep) == str(g["checksum"])
    lab = O.multiclass_labels(ep.s_label, case["K"], case["fg_idx"])
    assert int(lab[lab != 255].max()) < case["K"] and bool((lab == case["fg_idx"]).any())
    w = O.increment_inner_loop_ref(ep.f_s, lab, O.multiclass_w0(case["idx"], case["K"], case["C"]), case["lr"],
                                   case["n_iter"], case["fg_idx"], case["tp"])
    assert rel_err(w, g["W_fit"]) < 1e-6


PASCAL_MEAN, PASCAL_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


@pytest.mark.parametrize("name", golden_names("transform_"))
def test_oracle_val_transform_reproduces_reference_golden(name):
    """Resize -> ToTensor -> Normalize (src/dataset/transform.py:58-163 as composed in dataset.py:78-84): the restatement
    (cv2 does the resampling, as in the reference) against tensors recorded from the reference's own transform classes."""
    g = load_golden(name)
    case = g["case"]
    img, lab = O.synthetic_image(case["idx"], case["h"], case["w"])
    padding = [v * 255 for v in PASCAL_MEAN] if case["padding"] == "avg" else None
    o_img, o_lab = O.val_transform_ref(img, lab, case["size"], PASCAL_MEAN, PASCAL_STD, padding)
    assert float((o_img - torch.from_numpy(g["image"])).abs().max()) <= 2e-6          # same cv2 build: identical in practice
    assert torch.equal(o_lab, torch.from_numpy(g["label"]).long())
    from few_shot_seg_cwt_b200.transforms import find_new_hw
    nh, nw = find_new_hw(case["h"], case["w"], case["size"])
    assert nh % 8 == 0 and nw % 8 == 0 and max(nh, nw) <= case["size"]
    assert bool((o_lab[nh:, :] == 255).all()) and bool((o_lab[:, nw:] == 255).all())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_iou_counts_are_additive_and_consistent(seed):
    """Size-independent properties of src/util.py:279-308 that the full-size GPU checks rely on:
    the counts of a pixel set are the sum of the counts of any split of it, I <= min(pred area,
    T), U = pred area + T - I, ignored pixels count nowhere, and a relabelling of the ignored
    pixels' predictions changes nothing."""
    g = torch.Generator().manual_seed(seed)
    n = 473 * 7
    preds = torch.randint(0, 2, (n,), generator=g)
    target = torch.randint(0, 2, (n,), generator=g)
    target[torch.rand(n, generator=g) < 0.1] = O.IGNORE
    before = preds.clone()
    I, U, T = O.intersection_and_union_ref(preds, target, 2)
    assert torch.equal(preds, before)                                    # the oracle leaves its input alone
    cut = int(torch.randint(1, n - 1, (1,), generator=g))
    Ia, Ua, Ta = O.intersection_and_union_ref(preds[:cut], target[:cut], 2)
    Ib, Ub, Tb = O.intersection_and_union_ref(preds[cut:], target[cut:], 2)
    assert torch.equal(I, Ia + Ib) and torch.equal(U, Ua + Ub) and torch.equal(T, Ta + Tb)
    valid = target != O.IGNORE
    area = torch.stack([(preds[valid] == c).sum() for c in range(2)]).float()
    assert torch.equal(U, area + T - I)
    assert bool((I <= torch.minimum(area, T)).all())
    assert float(T.sum()) == float(valid.sum())
    flipped = preds.clone()
    flipped[~valid] = 1 - flipped[~valid]
    I2, U2, T2 = O.intersection_and_union_ref(flipped, target, 2)
    assert torch.equal(I, I2) and torch.equal(U, U2) and torch.equal(T, T2)
