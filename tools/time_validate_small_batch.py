"""validate_transformer at the reference scripts' batch_size_val 1 (scripts/test.sh:12) on a fake loader / stand-in backbone at
the PSPNet head geometry: episodes/s of the whole driver loop with the loader batches fused into head launches of
``head_batch`` episodes (1 = one launch per loader batch, as the reference iterates)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import few_shot_seg_cwt_b200 as cwt
from few_shot_seg_cwt_b200 import synthetic as syn

dev = torch.device("cuda:0")
N = int(os.environ.get("N", "128"))
eps = [syn.make_episode(500 + i % 16, label_dtype=torch.uint8) for i in range(16)]          # 16 distinct full-size episodes


class Backbone(torch.nn.Module):
    """extract_features looks device-resident features up by image id (the real backbone is timed by tools/time_backbone.py)"""
    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.feats = {}
        for i, ep in enumerate(eps):
            self.feats[2 * i], self.feats[2 * i + 1] = ep.f_s[0].to(dev), ep.f_q.to(dev)

    def extract_features(self, x):
        ids = x[:, 0, 0, 0].long().tolist()
        return torch.stack([self.feats[i] for i in ids]), None


items = []
for n in range(N):
    i = n % 16
    ep = eps[i]
    simg = torch.full((1, 1, 3, 8, 8), float(2 * i)); qimg = torch.full((1, 3, 8, 8), float(2 * i + 1))
    items.append((qimg, ep.q_label.unsqueeze(0), simg, ep.s_label.unsqueeze(0), [torch.tensor([ep.subcls])], None, None))


class A: pass
a = A()
a.test_num, a.batch_size_val, a.image_size, a.n_runs = N, 1, 473, 1
a.bottleneck_dim, a.num_classes_tr, a.cls_lr, a.adapt_iter = 512, 2, 0.1, 200
tr = cwt.MultiHeadAttentionOne(4, 512, 512, 512, dropout=0.5).to(dev)
tr.load_state_dict(syn.make_transformer_params(4, 512))
bb = Backbone().to(dev)
ref = None
for hb in (1, 4, 16, 64):
    cwt.validate_transformer(a, items[:16], bb, tr, verbose=False, head_batch=hb) if hb == 1 else None     # warm-up
    torch.manual_seed(0); torch.cuda.synchronize(); t0 = time.perf_counter()
    miou, loss = cwt.validate_transformer(a, items, bb, tr, verbose=False, head_batch=hb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ref = ref or (miou, loss)
    print(f"batch_size_val 1, head_batch {hb:3d}: {N / dt:7.0f} episodes/s ({1e3 * dt / N:6.3f} ms per episode)  mIoU {miou:.6f} loss {loss:.6f}"
          f"  same numbers as head_batch 1: {abs(miou - ref[0]) < 1e-9 and abs(loss - ref[1]) < 1e-6}", flush=True)
