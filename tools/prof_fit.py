"""Small fixed workload for ncu: E episodes of the full-size fit (n_iter steps) + transformer + logits/IoU."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=16)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--shot", type=int, default=1)
ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--fit-algo", type=int, default=0)
ap.add_argument("--attn-algo", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
E, S = a.episodes, a.shot
f_s = torch.relu(torch.randn(E, S, 512, 60, 60, device=dev, generator=g))
f_q = torch.relu(torch.randn(E, 512, 60, 60, device=dev, generator=g))
b = syn.make_batch([0, 1], shot=S).to(dev)
s_label = b.s_label[:1].expand(E, S, 473, 473).contiguous()
q_label = b.q_label[:1].expand(E, 473, 473).contiguous()
w0 = (torch.rand(E, 2, 512, device=dev, generator=g) * 2 - 1) / 512 ** 0.5
params = {k: v.to(dev) for k, v in syn.make_transformer_params(a.heads, 512).items()}
for _ in range(a.reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = cwt.episode_head(f_s, s_label, f_q, q_label, w0, params, a.heads, 0.1, a.iters, fit_algo=a.fit_algo, attn_algo=a.attn_algo)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"E={E} S={S} iters={a.iters}: {dt*1e3:.2f} ms  counts[0]={out.counts[0,0].tolist()}")
