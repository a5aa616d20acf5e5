"""Developer tool: end-to-end rate of HostPipeline.run (zero-compressed transport, 64 distinct episodes per step, 20 steps) for
several policies in ONE process — where the expansion kernel runs (1 head's stream, 0 copy stream, 2 its own stream), number of
staging slots, piece size — beside the device-resident HeadPipeline rate. Counts are compared with the device-resident run.
    python tools/e2e_policies.py [--steps 20]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--episodes", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
E = a.episodes
host = syn.make_batch(list(range(E))).pin_memory()
comp = cwt.compress_batch(host).pin_memory()
devb = host.to(dev)
params = syn.make_transformer_params(4, 512)
dparams = {k: v.to(dev) for k, v in params.items()}


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps, r


head = cwt.HeadPipeline(dev, dparams, 4, 0.1, 200, table=cwt.IoUTable(5, dev))
def resident():
    for _ in range(a.steps):
        out, _ = head.submit(devb.f_s, devb.s_label, devb.f_q, devb.q_label, devb.w0, devb.subcls)
    head.finish()
    return out
resident()
ms0, out0 = timed(resident)
print(f"device-resident HeadPipeline: {ms0:.3f} ms per step ({E / ms0 * 1e3:.0f} episodes/s)", flush=True)
for expand, slots, sub, ramp in ((1, 3, 32, (8, 8, 16)), (2, 3, 32, (8, 8, 16)), (2, 4, 32, (8, 8, 16)), (2, 4, 16, (8, 8)), (2, 5, 16, (8, 8)),
                                 (2, 3, 64, (8, 8, 16, 32)), (1, 3, 32, (8, 8, 16)), (2, 4, 32, (8, 8, 16))):
    pipe = cwt.HostPipeline(dev, params, 4, 0.1, 200, n_slots=slots, sub_batch=sub, expand_on_main=expand, ramp=ramp)
    pipe.run([comp] * 3)
    best = None
    for r in range(2):
        ms, res = timed(lambda: pipe.run([comp] * a.steps))
        best = ms if best is None else min(best, ms)
    same = bool(torch.equal(res[0].to(dev), out0.counts))
    print(f"expand {expand} slots {slots} pieces of {sub} ramp {ramp}: {best:.3f} ms per step ({E / best * 1e3:.0f} episodes/s, "
          f"{ms0 / best:.3f} of the device-resident rate), counts identical {same}", flush=True)
    del pipe
    torch.cuda.empty_cache()
