"""Time the resident fit kernel and dump its per-phase cycle counters (P1, X1, HR, P3, X2)."""
import argparse, ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=16)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--prof", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
E = a.episodes
f_s = torch.relu(torch.randn(E, 1, 512, 60, 60, device=dev, generator=g))
b = syn.make_batch([0]).to(dev)
s_label = b.s_label[:1].expand(E, 1, 473, 473).contiguous()
w0 = (torch.rand(E, 2, 512, device=dev, generator=g) * 2 - 1) / 512 ** 0.5
lib = L.load()
def run(algo):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w = cwt.fit_classifier(f_s, s_label, w0, 0.1, a.iters, check=False, algo=algo)
    e1.record(); torch.cuda.synchronize()
    return w, e0.elapsed_time(e1)
ws, ts = run(L.FIT_STREAM); ws, ts = run(L.FIT_STREAM)
wr, tr = run(L.FIT_RESIDENT); wr, tr = run(L.FIT_RESIDENT)
err = float((wr - ws).norm() / ws.norm())
print(f"E={E} iters={a.iters}: stream {ts:.2f} ms ({E/ts*1e3:.0f} ep/s)  resident {tr:.2f} ms ({E/tr*1e3:.0f} ep/s)  rel diff {err:.2e} finite={bool(torch.isfinite(wr).all())}")
if a.prof:
    buf = torch.zeros(320, 12, dtype=torch.int64, device=dev)
    # developer entry point (include/cwt_b200_debug.h): the same fit on the instrumented kernel, profile buffer per call
    nbytes = lib.cwt_fit_workspace_bytes(E, 1, 512, 60, 60, 473, 473)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    wr = torch.empty(E, 2, 512, device=dev)
    def run_prof():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.cwt_debug_fit_classifier_prof_f32(L.ptr(f_s), L.ptr(s_label), L.label_kind(s_label), L.ptr(w0), None, L.ptr(wr),
                                                   E, 512, 60, 60, 473, 473, a.iters, 0.1, 255, L.ptr(ws), ws.numel(),
                                                   L.ptr(buf), L.stream_ptr(dev))
        L.check(rc, "cwt_debug_fit_classifier_prof_f32")
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    run_prof(); buf.zero_(); tp = run_prof()
    c = buf.cpu().double()
    used = c[:, 0] > 0
    G = 3 if os.environ.get("CWT_RESIDENT_BPS") == "2" else 4          # episode groups of the plan (60x60x512)
    n_steps = a.iters * ((E + G - 1) // G)
    names = ["P1 (incl. all-reduce waits)", "halo wait (inside HR)", "HR (incl. halo wait)", "P3 (quad 0 only)", "n_waited_chunks", "sum(store->applied latency when waited)"]
    print(f"instrumented run {tp:.2f} ms; per-step cycles (mean over {int(used.sum())} CTAs, ~{n_steps} steps each; min/max over CTAs):")
    for i, n in enumerate(names):
        col = c[used, i] / n_steps
        print(f"  {n}: mean {col.mean():8.0f}  min {col.min():8.0f}  max {col.max():8.0f}")
    print(f"  total/step {c[used, 10].mean() / n_steps:8.0f} cycles (thread 0: P1 start -> its P3 store)")
    print(f"  applier lane 0: {c[used,6].sum()/c[used,9].sum():.2f} poll rounds per chunk, {c[used,7].sum()/c[used,9].sum():.0f} cycles polling per chunk "
          f"({c[used,7].sum()/c[used,6].sum():.0f} per round), own store -> chunk complete {c[used,8].sum()/c[used,9].sum():.0f} cycles")
    if c[used, 11].sum() > 0:
        print(f"  tensor-memory kernel, thread 0: {c[used, 11].mean() / n_steps:.2f} poll rounds per step; RED -> applied {c[used,5].sum()/c[used,4].sum():.0f} cycles "
              f"(includes the initial delay); min / max over CTAs {(c[used,5]/c[used,4]).min():.0f} / {(c[used,5]/c[used,4]).max():.0f}")
    print(f"  all-reduce latency seen by waiting compute warps: {(c[used,5].sum()/c[used,4].sum()):.0f} cycles (waited on {c[used,4].mean()/n_steps:.2f} chunks/step)")
