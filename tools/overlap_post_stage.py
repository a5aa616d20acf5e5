"""Experiment: run the post-fit stage (transformer + logits/IoU, 0.41 ms per 64 episodes) of step i on a second stream
while the resident fit of step i+1 (144 of the 148 SMs, one cooperative launch) runs. Prints ms per step for the serial
loop and the software-pipelined one, and checks that the counts are identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import ops, synthetic as syn

dev = torch.device("cuda:0")
E, heads, lr, T, steps = 64, 4, 0.1, 200, 10
hb = syn.make_batch(list(range(8)))
t = lambda x: x.repeat(E // 8, *([1] * (x.dim() - 1))).contiguous().to(dev)
b = syn.EpisodeBatch(*(t(x) for x in (hb.f_s, hb.s_label, hb.f_q, hb.q_label, hb.w0, hb.subcls, hb.idx)))
P = {k: v.to(dev) for k, v in syn.make_transformer_params(heads, 512).items()}


def post(w_fit):
    w_ad = ops.transformer_forward(w_fit, b.f_q, P["w_qkvs.weight"], P["fc.weight"], P["fc.bias"], P["layer_norm.weight"],
                                   P["layer_norm.bias"], heads, normalize_k=True)
    counts, ce, _ = ops.logits_iou(torch.stack([w_ad, w_fit], dim=1), b.f_q, b.q_label, normalize_mask=0b01, return_logits=False)
    return counts


def serial(n):
    out = None
    for _ in range(n):
        out = post(ops.fit_classifier(b.f_s, b.s_label, b.w0, lr, T, check=False))
    return out


def pipelined(n, prio, main_prio=None):
    """main_prio: run the fits on a stream of that priority instead of the caller's stream (-1 = above the side stream's 0:
    pending CTAs of the cooperative fit are then placed before pending CTAs of the post stage)."""
    caller = torch.cuda.current_stream(dev)
    main = caller if main_prio is None else torch.cuda.Stream(dev, priority=main_prio)
    side = torch.cuda.Stream(dev, priority=prio)
    out = None
    main.wait_stream(caller)
    with torch.cuda.stream(main):
        for _ in range(n):
            w_fit = ops.fit_classifier(b.f_s, b.s_label, b.w0, lr, T, check=False)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                w_fit.record_stream(side)
                out = post(w_fit)
        main.wait_stream(side)
    caller.wait_stream(main)
    return out


def timed(fn, *a):
    fn(3, *a[1:]) if a else fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn(*a) if a else fn(steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


ms_s, c_s = timed(serial)
print(f"serial     : {ms_s:.3f} ms per step ({E / ms_s * 1e3:.0f} episodes/s)", flush=True)
for prio, main_prio in ((0, None), (-1, None), (0, -1), (0, -2), (0, None), (0, -1)):
    ms_p, c_p = timed(pipelined, steps, prio, main_prio)
    print(f"pipelined  : {ms_p:.3f} ms per step ({E / ms_p * 1e3:.0f} episodes/s), side-stream priority {prio}, fit stream "
          f"{'= caller' if main_prio is None else 'priority %d' % main_prio}, counts identical {bool(torch.equal(c_s, c_p))}", flush=True)
print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a")
