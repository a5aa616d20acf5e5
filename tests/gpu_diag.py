"""GPU diagnostic: per-episode error report of every stage against the oracle (writes to stdout)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn
from oracle import head_ref as O

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

kw = dict(shot=1, C=64, h=12, w=12, H=89, W=89, style="unit")
n_head, lr, n_iter = 2, 0.1, 50
params = syn.make_transformer_params(n_head, 64)
pd = {k: v.to(dev) for k, v in params.items()}
idxs = [int(x) for x in sys.argv[1:]] or [100, 101, 102, 103, 104, 105]
for ldt in (torch.uint8, torch.int64):
    eps = [syn.make_episode(i, label_dtype=ldt, **kw) for i in idxs]
    torch.manual_seed(5)
    w0s = [torch.nn.Conv2d(64, 2, 1, bias=False).weight.detach().view(2, 64).clone() for _ in eps]
    b = syn.make_batch(idxs, label_dtype=ldt, **kw).to(dev)
    w0 = torch.stack(w0s).to(dev)
    out = cwt.episode_head(b.f_s, b.s_label, b.f_q, b.q_label, w0, pd, n_head, lr, n_iter, return_logits=True)
    for j, ep in enumerate(eps):
        o = O.episode_ref(ep.f_s, ep.s_label, ep.f_q, ep.q_label, w0s[j], params, n_head, lr, n_iter)
        single = cwt.episode_head(b.f_s[j:j+1], b.s_label[j:j+1], b.f_q[j:j+1], b.q_label[j:j+1], w0[j:j+1], pd, n_head, lr, n_iter, return_logits=True)
        print(f"{ldt} idx {ep.idx}: W_fit {rel(out.w_fit[j], o['W_fit']):.2e} W_ad {rel(out.w_adapted[j], o['W_adapted']):.2e} "
              f"lg {rel(out.logits60[j,0], o['logits60']):.2e} lg0 {rel(out.logits60[j,1], o['logits60_0']):.2e} "
              f"counts gpu {out.counts[j,0].tolist()} ref {o['counts'].tolist()} | single-vs-batch W_ad {rel(single.w_adapted[0], out.w_adapted[j]):.2e} "
              f"counts_single {single.counts[0,0].tolist()}")
