"""Sweep checker (test infrastructure, not collected by pytest): run the sharded sweep of
``few_shot_seg_cwt_b200.sweep`` on the GPU, then re-run a sample of its episodes through the CPU oracle and
report the per-episode count differences and the mIoU / FB-IoU gaps (BASELINE.json configs[3]:
"matching mIoU on 10k synthetic episodes").

    python tests/sweep_oracle_check.py --sample 8 -- --episodes 10000 --batch 64 --workers 14

Everything after ``--`` goes to the sweep.  The oracle is imported HERE only; the product package never does.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def check(dump_path: str) -> dict:
    from few_shot_seg_cwt_b200 import synthetic as syn
    from few_shot_seg_cwt_b200.episodic import IoUTable
    from oracle import head_ref as O

    d = torch.load(dump_path)
    params = syn.make_transformer_params(d["heads"], d["kw"]["C"])
    tg, to = IoUTable(d["num_classes_val"], "cpu"), IoUTable(d["num_classes_val"], "cpu")
    max_diff, n_tie = 0, 0
    for idx, got in sorted(d["counts"].items()):
        ep = syn.make_episode(idx, **d["kw"])
        o = O.episode_ref(ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0, params, d["heads"], d["cls_lr"], d["adapt_iter"])
        ref = torch.stack([o["counts"], o["counts0"]]).unsqueeze(0)
        got = got.cpu().unsqueeze(0)
        max_diff = max(max_diff, int((got - ref).abs().max()))
        n_tie += int((o["tie_margin"] <= 1e-5 * max(1.0, float(o["logits60"].abs().max()))).sum())
        sub = torch.tensor([ep.subcls])
        tg.update(got, sub)
        to.update(ref, sub)
    return {"episodes": len(d["counts"]), "max_count_diff": max_diff, "tie_set_pixels": n_tie,
            "mIoU_gpu": tg.miou(0), "mIoU_oracle": to.miou(0),
            "mIoU_gap_points": abs(tg.miou(0) - to.miou(0)) * 100,
            "FBIoU_gap_points": abs(tg.fb_iou(0) - to.fb_iou(0)) * 100}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    rest = []
    if "--" in argv:
        k = argv.index("--")
        argv, rest = argv[:k], argv[k + 1:]
    ap = argparse.ArgumentParser()
    ap.add_argument("--sample", type=int, default=8)
    a = ap.parse_args(argv)
    from few_shot_seg_cwt_b200 import sweep
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "counts.pt")
        out = sweep.main(rest + ["--dump-counts", dump, "--dump-sample", str(a.sample)])
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"oracle_check": check(dump)}), flush=True)
    return out


if __name__ == "__main__":
    main()
