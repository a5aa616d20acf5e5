"""GPU timeline of HostPipeline.run (zero-compressed transport): per piece, when its H2D copy starts / ends on the copy stream
(the copy may first wait for its staging slot; the expansion kernel runs there only with expand_on_main=False — by default it
runs on the head's stream in front of the fit), when the fit starts / ends (main stream) and when the post stage ends (side
stream) — ms since the start. SUB = sub-batch size, RAMP = comma-separated sizes at the start of the run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn, episodic, ops
dev = torch.device("cuda:0")
E, SUB, STEPS = 64, int(os.environ.get("SUB", "32")), 4
hb = syn.make_batch(list(range(8)))
rep = lambda t: t.repeat(E // 8, *([1] * (t.dim() - 1))).contiguous()
host = syn.EpisodeBatch(*(rep(t) for t in (hb.f_s, hb.s_label, hb.f_q, hb.q_label, hb.w0, hb.subcls, hb.idx)))
host = cwt.compress_batch(host).pin_memory() if os.environ.get("DENSE") != "1" else host.pin_memory()
params = syn.make_transformer_params(4, 512)
pipe = cwt.HostPipeline(dev, params, 4, 0.1, 200, sub_batch=SUB,
                        ramp=tuple(int(x) for x in os.environ.get("RAMP", "8,8,16").split(",") if x))
marks = []
def ev(stream, tag):
    e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append((tag, e))
pipe.run([host] * 2)
orig_stage, orig_fit, orig_submit = pipe._stage, ops.fit_classifier, episodic.HeadPipeline.submit
def stage(*a, **k):
    ev(pipe.copy_stream, "copy start"); r = orig_stage(*a, **k); ev(pipe.copy_stream, "copy done "); return r
def fit(*a, **k):
    m = torch.cuda.current_stream(dev); ev(m, "  fit start"); r = orig_fit(*a, **k); ev(m, "  fit done "); return r
def submit(self, *a, **k):
    out, done = orig_submit(self, *a, **k); marks.append(("    post done", done_timing(self))); return out, done
def done_timing(hp):
    e = torch.cuda.Event(enable_timing=True); e.record(hp.side); return e
pipe._stage, ops.fit_classifier, episodic.HeadPipeline.submit = stage, fit, submit
t0 = torch.cuda.Event(enable_timing=True); t0.record()
import time; w0 = time.perf_counter()
pipe.run([host] * STEPS)
torch.cuda.synchronize(); w1 = time.perf_counter()
print(f"sub_batch {SUB}: {STEPS} steps of {E} episodes, wall {1e3 * (w1 - w0):.2f} ms = {1e3 * (w1 - w0) / STEPS:.2f} ms per step")
for tag, e in marks: print(f"{tag} {t0.elapsed_time(e):8.2f} ms")
