import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
b = syn.make_batch([0]).to(dev)
def run(E, S, algo, iters=200):
    f_s = torch.relu(torch.randn(E, S, 512, 60, 60, device=dev, generator=g))
    s_label = b.s_label[:1, :1].expand(E, S, 473, 473).contiguous()
    w0 = (torch.rand(E, 2, 512, device=dev, generator=g) * 2 - 1) / 512 ** 0.5
    for _ in range(2): cwt.fit_classifier(f_s, s_label, w0, 0.1, iters, check=False, algo=algo)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): cwt.fit_classifier(f_s, s_label, w0, 0.1, iters, check=False, algo=algo)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return ms, E / ms * 1e3
for E in (2, 4, 6, 8, 10, 12, 16, 32):
    ms, r = run(E, 1, L.FIT_STREAM)
    print(f"stream  E={E:3d} S=1: {ms:8.2f} ms  {r:7.0f} ep/s  ({ms/200*1e3:.1f} us/step)")
for E in (1, 2, 4):
    ms, r = run(E, 5, L.FIT_STREAM)
    print(f"stream  E={E:3d} S=5: {ms:8.2f} ms  {r:7.0f} ep/s  ({ms/200*1e3:.1f} us/step)")
