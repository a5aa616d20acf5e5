"""GPU timeline of HostPipeline.run: when does each H2D copy and each head start / end (ms since the first event)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn, episodic
dev = torch.device("cuda:0")
E = 64
hb = syn.make_batch(list(range(4)))
rep = lambda t: t.repeat(E // 4, *([1] * (t.dim() - 1))).contiguous()
host = syn.EpisodeBatch(*(rep(t) for t in (hb.f_s, hb.s_label, hb.f_q, hb.q_label, hb.w0, hb.subcls, hb.idx))).pin_memory()
params = syn.make_transformer_params(4, 512)
pipe = cwt.HostPipeline(dev, params, 4, 0.1, 200)
marks = []
def ev(stream, tag):
    e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append((tag, e))
orig_stage, orig_head = pipe._stage, episodic.episode_head
def stage(hb_, slot, cap=0):
    ev(pipe.copy_stream, "copy issue"); r = orig_stage(hb_, slot, cap); ev(pipe.copy_stream, "copy done "); return r
def head(*a, **k):
    m = torch.cuda.current_stream(dev); ev(m, "head start"); r = orig_head(*a, **k); ev(m, "head done "); return r
pipe.run([host] * 3)
pipe._stage, episodic.episode_head = stage, head
t0 = torch.cuda.Event(enable_timing=True); t0.record()
pipe.run([host] * 6)
torch.cuda.synchronize()
for tag, e in marks: print(f"{tag} {t0.elapsed_time(e):8.2f} ms")
