"""Soak test of the TMA-ring kernels (streaming logits/IoU, rows_times_feat, feat_times_rows): many back-to-back launches on
fixed inputs at several batch sizes must return bit-identical results (no race in the mbarrier protocol, no hang)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import ops, synthetic as syn
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 500
g = torch.Generator(device=dev).manual_seed(1)
lab1 = syn.make_batch([0, 1, 2, 3]).q_label.to(dev)
t0 = time.time()
for E in (1, 3, 17, 64):
    f_q = torch.relu(torch.randn(E, 512, 60, 60, device=dev, generator=g))
    lab = lab1.repeat((E + 3) // 4, 1, 1)[:E].contiguous()
    wts = torch.randn(E, 2, 2, 512, device=dev, generator=g) * 0.05
    M = torch.randn(E, 8, 512, device=dev, generator=g)
    P = torch.randn(E, 8, 3600, device=dev, generator=g)
    ref = None
    for it in range(iters):
        c, ce, _ = ops.logits_iou(wts, f_q, lab, 0b01, return_logits=False)
        a = ops.rows_times_feat(M, f_q, True)
        b = ops.feat_times_rows(P, f_q, False)
        if it % 50 == 0 or it == iters - 1:
            cur = (c.clone(), ce.clone(), a.clone(), b.clone())
            if ref is None:
                ref = cur
            else:
                assert all(torch.equal(x, y) for x, y in zip(ref, cur)), f"E={E} iteration {it}: results changed"
    torch.cuda.synchronize()
    print(f"E={E}: {iters} iterations, bit-identical")
print(f"ok in {time.time() - t0:.1f} s")
