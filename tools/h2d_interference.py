"""Does a concurrent host->device copy slow the resident fit down (or the other way round)?  Times the fit (E = 64) and a
1 GB pinned H2D copy, each alone and both together."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn
dev = torch.device("cuda:0")
E = 64
hb = syn.make_batch(list(range(4)))
rep = lambda t: t.repeat(E // 4, *([1] * (t.dim() - 1))).contiguous()
f_s, s_label, w0 = rep(hb.f_s).to(dev), rep(hb.s_label).to(dev), rep(hb.w0).to(dev)
n = 972 * 1024 * 1024 // 4
src = torch.empty(n).pin_memory()
dst = torch.empty(n, device=dev)
side = torch.cuda.Stream()
def fit(): return cwt.fit_classifier(f_s, s_label, w0, 0.1, 200, check=False)
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(2): fit()
torch.cuda.synchronize()
def run(do_fit, do_copy, reps=5):
    tf = tc = 0.0
    for _ in range(reps):
        a0, a1, b0, b1 = ev(), ev(), ev(), ev()
        torch.cuda.synchronize()
        if do_copy:
            with torch.cuda.stream(side):
                b0.record(side); dst.copy_(src, non_blocking=True); b1.record(side)
        if do_fit:
            a0.record(); fit(); a1.record()
        torch.cuda.synchronize()
        if do_fit: tf += a0.elapsed_time(a1)
        if do_copy: tc += b0.elapsed_time(b1)
    return tf / reps, tc / reps
print("fit alone  : %.2f ms" % run(True, False)[0])
print("copy alone : %.2f ms (%.1f GB/s)" % (run(False, True)[1], n * 4 / run(False, True)[1] / 1e6))
f, c = run(True, True)
print("together   : fit %.2f ms, copy %.2f ms (%.1f GB/s)" % (f, c, n * 4 / c / 1e6))
