"""Soak test of the on-chip fit: many batch sizes / step counts against the streaming algorithm (full geometry)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn

dev = torch.device("cuda:0")
base = syn.make_batch(list(range(400, 416)), shot=1, C=512, h=60, w=60, H=473, W=473).to(dev)
worst = 0.0
for E, T in [(1, 1), (1, 200), (2, 7), (3, 50), (4, 200), (5, 33), (7, 200), (16, 200), (63, 60), (65, 60), (130, 40), (257, 10)]:
    idx = [i % 16 for i in range(E)]
    f_s, s_label, w0 = base.f_s[idx], base.s_label[idx], base.w0[idx]
    ws = cwt.fit_classifier(base.f_s, base.s_label, base.w0, 0.1, T, algo=L.FIT_STREAM)
    wr = cwt.fit_classifier(f_s, s_label, w0, 0.1, T, algo=L.FIT_RESIDENT)
    wr2 = cwt.fit_classifier(f_s, s_label, w0, 0.1, T, algo=L.FIT_RESIDENT)
    err = max(float((wr[i] - ws[idx[i]]).norm() / ws[idx[i]].norm()) for i in range(E))
    worst = max(worst, err)
    print(f"E={E:4d} T={T:4d}: max rel diff vs streaming {err:.2e}  bit-reproducible {bool(torch.equal(wr, wr2))}  finite {bool(torch.isfinite(wr).all())}")
    assert err < 2e-5 and torch.equal(wr, wr2)
print("worst", worst)
