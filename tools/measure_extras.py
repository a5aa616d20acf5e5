"""Round-1 side measurements (not the headline bench): L2 bandwidth, 5-shot head, COCO-20i sizing, training step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn

dev = torch.device("cuda:0")
out = {}

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# --- L2 bandwidth: STREAM-style copy of a working set that stays in the 126 MB L2 (2 x 24 MB), vs an HBM-sized one
for name, mb in (("l2_copy_2x24MB", 24), ("hbm_copy_2x1GB", 1024)):
    a = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    ms = timed(lambda: b.copy_(a), reps=20, warm=5)
    out[name + "_GBps"] = 2 * a.numel() * 4 / (ms * 1e-3) / 1e9
    del a, b

# --- L2 read bandwidth with a dedicated kernel (the torch copy above is launch-bound at 24 MB)
import ctypes
lib = L.load()
sink = torch.zeros(1, device=dev)
for mb in (16, 32, 64, 96):
    buf = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32, device=dev).normal_()
    iters = 40
    f = lambda: lib.cwt_debug_l2_read(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 4, iters, 148 * 4, ctypes.c_void_p(sink.data_ptr()),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    ms = timed(f, reps=3, warm=1)
    out[f"l2_read_{mb}MB_GBps"] = buf.numel() * 4 * iters / (ms * 1e-3) / 1e9
    del buf
buf = torch.empty(2 * 1024 * 1024 * 1024 // 4, dtype=torch.float32, device=dev).normal_()
ms = timed(lambda: lib.cwt_debug_l2_read(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 4, 2, 148 * 4, ctypes.c_void_p(sink.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), reps=3, warm=1)
out["hbm_read_2GB_GBps"] = buf.numel() * 4 * 2 / (ms * 1e-3) / 1e9
del buf

GEOM = dict(C=512, h=60, w=60, H=473, W=473)
def head_rate(E, shot, heads, n_cls=5, distinct=8):
    hb = syn.make_batch(list(range(distinct)), shot=shot, num_classes_val=n_cls, **GEOM)
    rep = (E + distinct - 1) // distinct
    t = lambda x: x.repeat(rep, *([1] * (x.dim() - 1)))[:E].contiguous().to(dev)
    b = syn.EpisodeBatch(*(t(x) for x in (hb.f_s, hb.s_label, hb.f_q, hb.q_label, hb.w0, hb.subcls, hb.idx)))
    params = {k: v.to(dev) for k, v in syn.make_transformer_params(heads, 512).items()}
    table = cwt.IoUTable(n_cls, dev)
    def step():
        o = cwt.episode_head(b.f_s, b.s_label, b.f_q, b.q_label, b.w0, params, heads, 0.1, 200)
        table.update(o.counts, b.subcls, o.ce)
    ms = timed(step, reps=3, warm=2)
    return E / (ms * 1e-3), ms

r, ms = head_rate(32, 5, 4); out["head_5shot_E32_eps_per_s"] = r; out["head_5shot_ms_per_32"] = ms
out["head_5shot_fit_GBps_algorithmic"] = 32 * (401 * 5 * 7372800) / (ms * 1e-3) / 1e9
r, ms = head_rate(64, 1, 4, n_cls=20); out["head_coco20i_1shot_E64_eps_per_s"] = r
r, ms = head_rate(64, 1, 1); out["head_1shot_heads1_E64_eps_per_s"] = r

# --- meta-training step (config 5): fit + transformer fwd (train mode) + query CE + backward + SGD-nesterov
for heads in (1, 4):
    tr = cwt.MultiHeadAttentionOne(heads, 512, 512, 512, dropout=0.5).to(dev)
    opt = torch.optim.SGD(tr.parameters(), lr=0.0025, momentum=0.9, weight_decay=1e-4, nesterov=True)
    b = syn.make_batch([0], **GEOM).to(dev)
    ms_all = timed(lambda: cwt.meta_train_step(tr, opt, b.f_s, b.s_label, b.f_q, b.q_label, b.w0, 0.1, 200), reps=5, warm=2)
    w_fit = cwt.fit_classifier(b.f_s, b.s_label, b.w0, 0.1, 200, check=False)
    def tstep():
        tr.train(); tr.normalize_k = True
        upd = tr(w_fit, b.f_q, b.f_q)
        loss, _ = cwt.query_loss(upd, b.f_q, b.q_label)
        opt.zero_grad(); loss.backward(); opt.step()
    ms_t = timed(tstep, reps=10, warm=3)
    out[f"train_step_heads{heads}_ms_total_E1"] = ms_all
    out[f"train_step_heads{heads}_ms_transformer_fwd_bwd_opt_E1"] = ms_t
print(json.dumps(out, indent=1))
