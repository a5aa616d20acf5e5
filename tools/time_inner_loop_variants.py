"""Developer tool: throughput of the `inner_loop` variants at the head geometry (512 x 60 x 60 -> 473 x 473,
1 shot, 200 steps), beside the default weighted-CE fit. The variants run on the streaming kernels (several launches per SGD
step over the whole batch), so they are timed at E = 1 (one classifier, the reference's call pattern) and at a batch of E.
    python tools/time_inner_loop_variants.py [--episodes 16] [--iters 200]      (prints one JSON object)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, ops, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=16)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--distinct", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda:0")
T, lr = a.iters, 0.1


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.launch_count()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
        launches = L.launch_count() - n0
    return best, launches


out = {"geometry": "f_s [E,1,512,60,60] fp32, labels 473x473, %d SGD steps" % T, "rows": {}}
for E in (1, a.episodes):
    base = syn.make_batch(list(range(900, 900 + min(a.distinct, E)))).to(dev)
    idx = [i % base.f_s.shape[0] for i in range(E)]
    f_s, lab, w0 = base.f_s[idx].contiguous(), base.s_label[idx].contiguous(), base.w0[idx].contiguous()
    b0 = torch.zeros(E, 2, device=dev)
    x_norm = ops.normalize_features(f_s.reshape(E, 512, 60, 60), eps=1e-5, scale=1.0).reshape(E, 1, 512, 60, 60)
    ones = torch.ones(2, device=dev)
    scale = torch.full((E,), 2.0, device=dev)
    g0 = w0.norm(dim=2)
    variants = {
        "wt_ce (default fit, on chip)": lambda: ops.fit_classifier(f_s, lab, w0, lr, T, check=False),
        "wt_ce (streaming kernels)": lambda: ops.fit_classifier(f_s, lab, w0, lr, T, check=False, algo=L.FIT_STREAM),
        "ce (class weight [1,1])": lambda: ops.fit_classifier(f_s, lab, w0, lr, T, class_weight=ones, check=False),
        "wt_dc / dc (dice)": lambda: ops.fit_classifier_dice(f_s, lab, w0, lr, T, check=False),
        "classifier with bias": lambda: ops.fit_classifier_bias(f_s, lab, w0, b0, lr, T, check=False),
        "CosCls 'r' (weight norm)": lambda: ops.fit_coscls(x_norm, lab, w0, scale, lr, T, weight_g=g0, flags=L.COSCLS_R, check=False),
        "CosCls 'n' + 't' + bias": lambda: ops.fit_coscls(x_norm, lab, w0, scale, lr, T, bias=b0, flags=L.COSCLS_N | L.COSCLS_T, check=False),
    }
    for name, fn in variants.items():
        ms, launches = timed(fn)
        out["rows"].setdefault(name, {})[f"E={E}"] = {"ms": round(ms, 3), "episodes_per_s": round(E / ms * 1e3, 1),
                                                      "launches": launches, "us_per_step": round(ms * 1e3 / T, 2)}
# increment_inner_loop with K > 2 classes (one classifier per call)
for K in (17, 62):
    base = syn.make_batch([950]).to(dev)
    f1 = base.f_s[0].contiguous()                         # [1,512,60,60]
    lab1 = base.s_label[0].clone()
    lab1[lab1 == 1] = K - 1
    wk = torch.randn(K, 512, device=dev) * 0.05
    cwk = torch.ones(K, device=dev)
    ms, launches = timed(lambda: ops.fit_multiclass(f1, lab1, wk, cwk, lr, T, check=False))
    out["rows"][f"increment_inner_loop K={K}"] = {"E=1": {"ms": round(ms, 3), "episodes_per_s": round(1e3 / ms, 1), "launches": launches,
                                                          "us_per_step": round(ms * 1e3 / T, 2)}}
print(json.dumps(out, indent=1))
