"""Developer tool: time the resident fit (E episodes x T steps, PSPNet head geometry) with the library named by CWT_LIB_PATH
(product build by default), print ms / clk per step, and compare W against a reference file written by the product build.
    python tools/ablate_resident.py --save /tmp/w.pt          # product build
    CWT_LIB_PATH=tools/variants/libcwt_v0x40.so python tools/ablate_resident.py --ref /tmp/w.pt"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=64)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--distinct", type=int, default=16)
ap.add_argument("--save"); ap.add_argument("--ref"); ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
E = a.episodes
base = syn.make_batch(list(range(400, 400 + a.distinct))).to(dev)
idx = [i % a.distinct for i in range(E)]
f_s, s_label, w0 = base.f_s[idx].contiguous(), base.s_label[idx].contiguous(), base.w0[idx].contiguous()
ts = []
for r in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w = cwt.fit_classifier(f_s, s_label, w0, 0.1, a.iters, check=False, algo=L.FIT_RESIDENT)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
w2 = cwt.fit_classifier(f_s, s_label, w0, 0.1, a.iters, check=False, algo=L.FIT_RESIDENT)
t = min(ts[1:])
G = min(4, E)
steps = a.iters * ((E + G - 1) // G)
msg = f"{a.tag or os.environ.get('CWT_LIB_PATH', 'product'):40s} E={E} T={a.iters}: {t:7.3f} ms  {E / t * 1e3:7.0f} ep/s  {t * 1e-3 * 1.965e9 / steps:7.0f} clk/step@1965  reproducible={bool(torch.equal(w, w2))} finite={bool(torch.isfinite(w).all())}"
if a.save:
    torch.save(w.cpu(), a.save)
if a.ref:
    ref = torch.load(a.ref).to(dev)
    msg += f"  rel diff vs product {float((w - ref).norm() / ref.norm()):.2e} bit-equal={bool(torch.equal(w, ref))}"
print(msg, flush=True)
