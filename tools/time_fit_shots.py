"""Developer tool: the multi-shot fit, streaming algorithm vs the L2-resident persistent kernel (CWT_FIT_L2), full head geometry.
    python tools/time_fit_shots.py --shot 5 --episodes 32       (CWT_FIT_L2_NT / CWT_FIT_L2_MB select the plan)"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--shot", type=int, default=5)
ap.add_argument("--episodes", type=int, default=32)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--distinct", type=int, default=4)
ap.add_argument("--algos", default="1,3")
a = ap.parse_args()
dev = torch.device("cuda:0")
E = a.episodes
base = syn.make_batch(list(range(800, 800 + a.distinct)), shot=a.shot).to(dev)
idx = [i % a.distinct for i in range(E)]
f_s, s_label, w0 = base.f_s[idx].contiguous(), base.s_label[idx].contiguous(), base.w0[idx].contiguous()
res = {}
for algo in [int(x) for x in a.algos.split(",")]:
    ts = []
    for r in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w = cwt.fit_classifier(f_s, s_label, w0, 0.1, a.iters, check=False, algo=algo)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[algo] = (min(ts[1:]), w)
    F = a.shot * 3600 * 512 * 4
    print(f"shot={a.shot} E={E} T={a.iters} algo={algo} NT={os.environ.get('CWT_FIT_L2_NT','auto')} MB={os.environ.get('CWT_FIT_L2_MB','100')}: "
          f"{min(ts[1:]):8.2f} ms  {E / min(ts[1:]) * 1e3:7.1f} episodes/s  {min(ts[1:]) / E * 1e3 / a.iters:6.2f} us/episode-step  "
          f"algorithmic {(2 * a.iters + 1) * F * E / min(ts[1:]) / 1e6:7.0f} GB/s", flush=True)
if 1 in res and 3 in res:
    print("   rel diff L2 vs stream", float((res[3][1] - res[1][1]).norm() / res[1][1].norm()))
