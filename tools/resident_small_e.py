"""Latency of the resident fit at small batch sizes for every tile the planner could take (CWT_RESIDENT_TILE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import _lib as L, synthetic as syn

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
b = syn.make_batch([0]).to(dev)


def run(E, tile, iters=200):
    if tile:
        os.environ["CWT_RESIDENT_TILE"] = tile
    else:
        os.environ.pop("CWT_RESIDENT_TILE", None)
    f_s = torch.relu(torch.randn(E, 1, 512, 60, 60, device=dev, generator=g))
    s_label = b.s_label[:1, :1].expand(E, 1, 473, 473).contiguous()
    w0 = (torch.rand(E, 2, 512, device=dev, generator=g) * 2 - 1) / 512 ** 0.5
    ref = cwt.fit_classifier(f_s, s_label, w0, 0.1, iters, check=False, algo=L.FIT_STREAM)
    for _ in range(2):
        out = cwt.fit_classifier(f_s, s_label, w0, 0.1, iters, check=False, algo=L.FIT_RESIDENT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        cwt.fit_classifier(f_s, s_label, w0, 0.1, iters, check=False, algo=L.FIT_RESIDENT)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    rel = float((out - ref).norm() / ref.norm())
    return ms, rel


tiles = sys.argv[1:] or ["", "20x5", "20x4", "20x3", "20x2", "12x6", "12x5", "12x4", "12x3", "4x10", "4x12"]
for E in (1, 2, 3, 4, 8):
    for t in tiles:
        try:
            ms, rel = run(E, t)
            print(f"E={E:2d} tile={t or 'auto':>5s}: {ms:7.3f} ms  ({ms / 200 * 1e3:5.2f} us/step/batch)  rel vs stream {rel:.1e}", flush=True)
        except Exception as ex:  # unsupported tile
            print(f"E={E:2d} tile={t:>5s}: {type(ex).__name__}: {str(ex)[:90]}", flush=True)
