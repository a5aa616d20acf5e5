"""Sharded evaluation sweep over synthetic episodes (BASELINE.json configs[3]).

    python -m few_shot_seg_cwt_b200.sweep --episodes 10000 --batch 32 --workers 8
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m few_shot_seg_cwt_b200.sweep --episodes 10000

Episode i goes to rank i mod world; every rank generates its own episodes with DataLoader
workers (the reference uses ``DataLoader(workers=2)``, pascal.yaml:12), stages them through pinned memory and
``HostPipeline``; the only collective is the int64 all-reduce of the IoU table. Rank 0 prints one JSON line
with mIoU (adapted / baseline classifier), FB-IoU, mean loss and episodes/s. ``--dump-counts PATH --dump-sample K``
saves the per-episode counts of K episodes; the checker that re-runs them through the CPU oracle lives with the
tests (``python tests/sweep_oracle_check.py``) — this package never imports ``oracle/``.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch

from . import synthetic as syn
from .episodic import HostPipeline


class _Episodes(torch.utils.data.Dataset):
    def __init__(self, indices, kw):
        self.indices, self.kw = indices, kw

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        e = syn.make_episode(self.indices[i], **self.kw)
        return e.f_s, e.s_label, e.f_q, e.q_label, e.w0, e.subcls, e.idx


def _collate(items):
    f_s, s_l, f_q, q_l, w0, sub, idx = zip(*items)
    return syn.EpisodeBatch(torch.stack(f_s), torch.stack(s_l), torch.stack(f_q), torch.stack(q_l), torch.stack(w0),
                            torch.tensor(sub, dtype=torch.int64), torch.tensor(idx, dtype=torch.int64))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--shot", type=int, default=1)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--cls-lr", type=float, default=0.1)
    ap.add_argument("--adapt-iter", type=int, default=200)
    ap.add_argument("--style", default="unit")
    ap.add_argument("--num-classes-val", type=int, default=5)
    ap.add_argument("--dump-counts", default="", help="torch.save the per-episode counts of the first --dump-sample episodes here")
    ap.add_argument("--dump-sample", type=int, default=0)
    ap.add_argument("--small", action="store_true", help="64-channel 12x12 -> 89x89 episodes (quick checks)")
    a = ap.parse_args(argv)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("the sweep needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    geom = dict(C=64, h=12, w=12, H=89, W=89) if a.small else dict(C=512, h=60, w=60, H=473, W=473)
    kw = dict(shot=a.shot, style=a.style, num_classes_val=a.num_classes_val, **geom)
    mine = syn.shard_indices(a.episodes, rank, world)
    loader = torch.utils.data.DataLoader(_Episodes(mine, kw), batch_size=a.batch, shuffle=False, num_workers=a.workers,
                                         collate_fn=_collate, pin_memory=False, persistent_workers=False)
    params = syn.make_transformer_params(a.heads, geom["C"])
    pipe = HostPipeline(dev, params, a.heads, a.cls_lr, a.adapt_iter, num_classes_val=a.num_classes_val)

    kept = {}

    def batches():
        for b in loader:
            if a.dump_counts and rank == 0 and len(kept) < a.dump_sample:
                for j in range(b.n_episodes):
                    if len(kept) < a.dump_sample:
                        kept[int(b.idx[j])] = None
            yield b.pin_memory()

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    results = pipe.run(batches())                 # one integer all-reduce of the table at the end (every rank calls it once)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0

    import hashlib
    tb = pipe.table
    digest = hashlib.sha256(tb.cls.cpu().numpy().tobytes() + tb.fb.cpu().numpy().tobytes()).hexdigest()
    out = {"episodes": int(pipe.table.n_episodes), "world": world, "bad_episodes": int(tb.n_bad),
           # the whole int64 table: equal bit for bit for every world size (src/test.py:225-251 accumulates the same sums)
           "table_sha256": digest, "table_cls_I_U": tb.cls.cpu().tolist(), "table_fb_I_U": tb.fb.cpu().tolist(), "seconds": dt, "episodes_per_s_incl_generation": a.episodes / dt,
           "mIoU_adapted": pipe.table.miou(0), "mIoU_baseline": pipe.table.miou(1), "FBIoU_adapted": pipe.table.fb_iou(0),
           "FBIoU_baseline": pipe.table.fb_iou(1), "loss_adapted": pipe.table.mean_loss(0),
           "class_iou_adapted": pipe.table.class_iou(0), "config": {**kw, "heads": a.heads, "cls_lr": a.cls_lr, "adapt_iter": a.adapt_iter}}

    if a.dump_counts and rank == 0:
        # per-episode counts of the first --dump-sample episodes of rank 0, for tests/sweep_oracle_check.py
        counts_by_idx, pos = {}, 0
        for r in results:
            for j in range(r.shape[0]):
                if mine[pos] in kept:
                    counts_by_idx[mine[pos]] = r[j].clone()
                pos += 1
        torch.save({"counts": counts_by_idx, "kw": kw, "heads": a.heads, "cls_lr": a.cls_lr, "adapt_iter": a.adapt_iter,
                    "num_classes_val": a.num_classes_val}, a.dump_counts)
        out["dumped_counts"] = {"path": a.dump_counts, "episodes": len(counts_by_idx)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
