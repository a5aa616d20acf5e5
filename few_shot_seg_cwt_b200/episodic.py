"""Episodic drivers: the batched per-episode head, the reference's ``validate_transformer`` loop
(src/test.py:103-254), the sharded multi-GPU sweep and the meta-training step
(src/train.py:233-267).

Episodes are independent (fresh classifier per episode, frozen backbone and transformer), so a
sweep is sharded by episode index — rank r of G takes episodes r, r+G, ... — with no data-path
collective; the only exchange is one integer all-reduce of the per-class intersection / union
table at the end (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import time
from collections import defaultdict
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .synthetic import EpisodeBatch, make_batch, shard_indices
from .hostformat import CompressedEpisodeBatch, CompressedMap, expand_map

IGNORE = 255


def transformer_params(transformer) -> Dict[str, torch.Tensor]:
    """The five tensors of the block, by the reference's state-dict names (module or dict)."""
    sd = transformer if isinstance(transformer, dict) else {
        "w_qkvs.weight": transformer.w_qkvs.weight, "fc.weight": transformer.fc.weight,
        "fc.bias": transformer.fc.bias, "layer_norm.weight": transformer.layer_norm.weight,
        "layer_norm.bias": transformer.layer_norm.bias}
    return {k: v.detach() for k, v in sd.items()}


@dataclass
class HeadOutput:
    w_fit: torch.Tensor        # [E,2,C]
    w_adapted: torch.Tensor    # [E,2,C]
    counts: torch.Tensor       # int64 [E,2,2,3]  variant (0 adapted / 1 baseline) x class x (I,U,T)
    ce: torch.Tensor           # float64 [E,2,2]  variant x (sum -log p[y], #valid)
    logits60: Optional[torch.Tensor]   # [E,2,2,h,w] variant x class
    status: Optional[torch.Tensor] = None   # int32 [E] deferred error word of the fit (ops.fit_status), still on the device


def episode_head(f_s, s_label, f_q, q_label, w0, params: Dict[str, torch.Tensor], n_head: int, lr: float,
                 n_iter: int, return_logits: bool = False, fit_algo: int = L.FIT_AUTO,
                 attn_algo: int = L.ATTN_REASSOC) -> HeadOutput:
    """The head of E evaluation episodes (src/test.py:162-234), fully on the device, no host sync.

    f_s [E,S,C,h,w]; s_label [E,S,H,W]; f_q [E,C,h,w] (un-normalised backbone features);
    q_label [E,H,W]; w0 [E,2,C]."""
    # no host sync: the reference's ZeroDivisionError (empty support mask, :174) / bad labels / non-finite weights travel as
    # a per-episode status word that the caller reads back with the counts (ops.raise_for_status)
    w_fit, status = ops.fit_classifier(f_s, s_label, w0, lr, n_iter, check=False, algo=fit_algo, return_status=True)   # :164-187
    w_ad = ops.transformer_forward(w_fit, f_q, params["w_qkvs.weight"], params["fc.weight"], params["fc.bias"],
                                   params["layer_norm.weight"], params["layer_norm.bias"], n_head,
                                   normalize_k=True, algo=attn_algo)                            # :194-197
    weights = torch.stack([w_ad, w_fit], dim=1)                                                 # [E,2,2,C]
    counts, ce, logits = ops.logits_iou(weights, f_q, q_label, normalize_mask=0b01,
                                        return_logits=return_logits)                            # :192,200-223
    return HeadOutput(w_fit, w_ad, counts, ce, logits, status)


class HeadPipeline:
    """Software-pipelined :func:`episode_head` for a SEQUENCE of batches.

    The resident fit is one cooperative launch on 144 of the 148 SMs; the stage behind it (transformer + fused
    logits / up-sample / IoU, 0.41 ms per 64 episodes) only needs the fit's result. ``submit`` therefore queues the fit of
    batch i on the caller's stream and the post stage of batch i on a side stream, so that it runs while the fit of batch
    i+1 already occupies the chip (measured on B200, 64 episodes per batch: 16.87 ms per batch instead of 17.32 —
    ``tools/overlap_post_stage.py`` — i.e. the post stage disappears behind the next fit; results are identical).

    The outputs of ``submit`` live on the side stream: call :meth:`finish` (or use ``done`` events) before reading them
    from another stream, and do not overwrite a submitted batch before its ``done`` event."""

    def __init__(self, device, params: Dict[str, torch.Tensor], n_head: int, lr: float, n_iter: int,
                 fit_algo: int = L.FIT_AUTO, attn_algo: int = L.ATTN_REASSOC, table: Optional["IoUTable"] = None,
                 reduce_every_step: bool = False):
        self.device = torch.device(device)
        self.params, self.n_head, self.lr, self.n_iter = params, n_head, lr, n_iter
        self.fit_algo, self.attn_algo = fit_algo, attn_algo
        self.table, self.reduce_every_step = table, reduce_every_step
        self.side = torch.cuda.Stream(self.device)

    def submit(self, f_s, s_label, f_q, q_label, w0, subcls=None, reduce: Optional[bool] = None, after=None):
        """Queue one batch. ``after(out)`` (optional) runs on the side stream right behind the post stage (e.g. the
        asynchronous copy of the counts to the host). Returns (HeadOutput, done event)."""
        main = torch.cuda.current_stream(self.device)
        w_fit, status = ops.fit_classifier(f_s, s_label, w0, self.lr, self.n_iter, check=False, algo=self.fit_algo,
                                           return_status=True)
        self.side.wait_stream(main)
        p = self.params
        with torch.cuda.stream(self.side):
            for t in (w_fit, status, f_q, q_label) + ((subcls,) if subcls is not None else ()):
                t.record_stream(self.side)
            w_ad = ops.transformer_forward(w_fit, f_q, p["w_qkvs.weight"], p["fc.weight"], p["fc.bias"],
                                           p["layer_norm.weight"], p["layer_norm.bias"], self.n_head,
                                           normalize_k=True, algo=self.attn_algo)
            counts, ce, _ = ops.logits_iou(torch.stack([w_ad, w_fit], dim=1), f_q, q_label, normalize_mask=0b01,
                                           return_logits=False)
            out = HeadOutput(w_fit, w_ad, counts, ce, None, status)
            if self.table is not None and subcls is not None:
                self.table.update(counts, subcls, ce, status)
                if self.reduce_every_step if reduce is None else reduce:
                    self.table.all_reduce()
            if after is not None:
                after(out)
            done = torch.cuda.Event()
            done.record(self.side)
        return out, done

    def finish(self) -> None:
        """Make the caller's stream wait for everything submitted so far."""
        torch.cuda.current_stream(self.device).wait_stream(self.side)


class IoUTable:
    """Device-resident int64 accumulator of the sweep metrics:
    ``cls[c, v, 0/1]`` = foreground intersection / union of class c (1..num_classes) for variant v
    (0 adapted, 1 baseline) — what src/test.py:225-230 keeps in dicts of float32 tensors — and
    ``fb[v, k, 0/1]`` = class-agnostic background / foreground I, U for FB-IoU."""

    def __init__(self, num_classes: int, device, n_variants: int = 2):
        self.num_classes = num_classes
        self.cls = torch.zeros(num_classes + 1, n_variants, 2, dtype=torch.int64, device=device)
        self.fb = torch.zeros(n_variants, 2, 2, dtype=torch.int64, device=device)
        self.ce = torch.zeros(n_variants, 2, dtype=torch.float64, device=device)
        self.n_episodes = torch.zeros(1, dtype=torch.int64, device=device)
        self.n_bad = torch.zeros(1, dtype=torch.int64, device=device)      # episodes whose fit status was non-zero
        # What this rank accumulated since the last all_reduce (``_local_*``) is kept apart from the table the metrics
        # read, so that all_reduce can be called any number of times (per step or once at the end): every call adds the
        # sum over ranks of the not-yet-exchanged DELTAS to every rank's table — nothing is ever counted twice.
        self._local = [torch.zeros_like(t) for t in self._fields()]
        self._distributed = False

    def _fields(self):
        return [self.cls, self.fb, self.n_episodes, self.n_bad, self.ce]

    def update(self, counts: torch.Tensor, subcls: torch.Tensor, ce: Optional[torch.Tensor] = None,
               status: Optional[torch.Tensor] = None) -> None:
        """counts int64 [E,V,2,3]; subcls int64 [E] in 1..num_classes; status int32 [E] (episodes with a non-zero fit
        status are counted in ``n_bad`` — their counts are still added, as the reference has no such episodes)."""
        tgt = self._local if self._distributed else self._fields()
        fg = counts[:, :, 1, :2]                                   # [E,V,(I,U)]  "do not count background"
        tgt[0].index_add_(0, subcls.to(self.cls.device), fg)
        tgt[1] += counts[:, :, :, :2].sum(0)
        tgt[2] += counts.shape[0]
        if status is not None:
            tgt[3] += (status != 0).sum()
        if ce is not None:
            tgt[4] += ce.sum(0)

    def all_reduce(self) -> None:
        """The one collective of a sharded sweep: integer sum over ranks (NVLink / NVSwitch via NCCL). Idempotent: only
        what was accumulated since the previous call is exchanged, so per-step and end-of-sweep reductions give the same
        table. Every rank must call it the same number of times."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if not self._distributed:
            # first call: everything accumulated so far is this rank's delta
            self._distributed = True
            for loc, t in zip(self._local, self._fields()):
                loc.copy_(t)
                t.zero_()
        ints = self._local[:4]
        flat = torch.cat([t.reshape(-1) for t in ints])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        dist.all_reduce(self._local[4], op=dist.ReduceOp.SUM)
        off = 0
        for loc, t in zip(ints, self._fields()[:4]):
            t += flat[off:off + t.numel()].view_as(t)
            off += t.numel()
            loc.zero_()
        self.ce += self._local[4]
        self._local[4].zero_()

    def miou(self, variant: int = 0) -> float:
        """mean over classes seen of I_c / (U_c + 1e-10)   (src/test.py:232-243)."""
        c = self.cls[1:, variant].cpu().double()
        seen = c[:, 1] > 0
        if not bool(seen.any()):
            return 0.0
        return float((c[seen, 0] / (c[seen, 1] + 1e-10)).mean())

    def class_iou(self, variant: int = 0) -> Dict[int, float]:
        c = self.cls[:, variant].cpu().double()
        return {i: float(c[i, 0] / (c[i, 1] + 1e-10)) for i in range(1, self.num_classes + 1) if c[i, 1] > 0}

    def fb_iou(self, variant: int = 0) -> float:
        """Standard FB-IoU: mean of the global background and foreground IoU."""
        f = self.fb[variant].cpu().double()
        return float((f[:, 0] / (f[:, 1] + 1e-10)).mean())

    def mean_loss(self, variant: int = 0) -> float:
        c = self.ce[variant].cpu()
        return float(c[0] / c[1].clamp_min(1.0))


# ----------------------------------------------------------------------------------------
# synthetic sharded sweep (BASELINE.json configs[3])
# ----------------------------------------------------------------------------------------
def run_sweep(n_episodes: int, params: Dict[str, torch.Tensor], n_head: int, lr: float, n_iter: int,
              device, rank: int = 0, world: int = 1, batch: int = 16, num_classes_val: int = 5,
              gen_kwargs: Optional[dict] = None, reduce: bool = True) -> IoUTable:
    """Evaluate synthetic episodes ``start..start+n_episodes`` sharded as i -> rank i mod world."""
    gen_kwargs = dict(gen_kwargs or {})
    gen_kwargs.setdefault("num_classes_val", num_classes_val)
    table = IoUTable(num_classes_val, device)
    mine = shard_indices(n_episodes, rank, world)
    p = {k: v.to(device) for k, v in params.items()}
    for i in range(0, len(mine), batch):
        b = make_batch(mine[i:i + batch], **gen_kwargs).to(device)
        out = episode_head(b.f_s, b.s_label, b.f_q, b.q_label, b.w0, p, n_head, lr, n_iter)
        table.update(out.counts, b.subcls, out.ce, out.status)        # episodes with a bad fit status are counted in table.n_bad
    if reduce:
        table.all_reduce()
    return table


# ----------------------------------------------------------------------------------------
# validate_transformer drop-in (src/test.py:103-254)
# ----------------------------------------------------------------------------------------
def validate_transformer(args, val_loader, model, transformer, verbose: bool = True,
                         overlap_backbone: bool = True, head_batch: Optional[int] = None) -> Tuple[float, float]:
    """Same arguments and return value as the reference: ``(mean mIoU over runs, mean loss over runs)``.

    ``args`` needs test_num, batch_size_val, image_size, n_runs, bottleneck_dim, num_classes_tr, cls_lr,
    adapt_iter (and optionally heads). Loader items are the reference's 7-tuples
    ``(qry_img, q_label, spprt_imgs, s_label, subcls, spprt_oris, qry_oris)`` with a leading batch
    dimension of 1 (src/dataset/dataset.py:326-327). The backbone ``model.extract_features`` stays the
    caller's PyTorch module; everything after it runs in the fused kernels, batch_size_val episodes
    per launch.

    Backbone -> head handoff. The reference calls ``extract_features`` twice per episode at batch
    size ``shot`` and 1 (src/test.py:177-178,190-191). Here the ``batch_size_val`` episodes of a batch go through the
    backbone as TWO calls (all support images, all query images; eval-mode BatchNorm makes the images independent),
    on a side stream: the backbone of batch i+1 is queued before the head of batch i, so its launches, its H2D copies
    and — whenever the cooperative fit leaves SMs free (post stage, launch gaps) — its kernels overlap the head; the
    head waits on an event, never on the host. Results (counts, CE, the fit's status word) are read back one batch
    late; an empty support mask still raises the reference's ``ZeroDivisionError`` (src/test.py:174), bad labels a
    ``ValueError``, then.

    ``head_batch`` (or ``args.head_batch``): episodes per head launch. The reference's scripts run ``batch_size_val 1``
    (scripts/test.sh:12), and one episode occupies only 36 of the 148 SMs (0.71 ms per fit whether 1 or 4 episodes are in
    it). Episodes are independent, so several loader batches are fused into ONE backbone call pair + ONE head launch of
    at least ``head_batch`` episodes (default 16; never across the end of a run); the per-batch loss (the reference logs
    the CE of each ``batch_size_val`` episodes), the per-class accumulation, the progress lines and the classifier-init RNG
    order stay per reference batch, so the results are the same numbers."""
    if verbose:
        print('==> Start testing')
    model.eval()
    transformer.eval()
    device = next(transformer.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("validate_transformer (cwt_b200): the transformer must live on a CUDA device")
    nb_episodes = int(args.test_num / args.batch_size_val)
    params = transformer_params(transformer)
    n_head = transformer.n_head
    C = args.bottleneck_dim
    B = int(args.batch_size_val)
    if head_batch is None:
        head_batch = int(getattr(args, "head_batch", 16) or 16)
    K = max(1, -(-int(head_batch) // max(B, 1)))        # reference batches fused into one head launch
    main = torch.cuda.current_stream(device)
    bb_stream = torch.cuda.Stream(device) if overlap_backbone else main

    runtimes = torch.zeros(args.n_runs)
    val_IoUs = np.zeros(args.n_runs)
    val_losses = np.zeros(args.n_runs)
    iter_loader = iter(val_loader)
    iter_num = 0

    def next_item():
        nonlocal iter_loader
        try:
            return next(iter_loader)
        except StopIteration:
            iter_loader = iter(val_loader)
            return next(iter_loader)

    def stage_batch(n_ref: int):
        """Load ``n_ref`` reference batches of ``B`` episodes, draw their classifier inits and queue the two backbone calls on
        the backbone stream."""
        nonlocal iter_num
        sp, sl, qi, ql, w0_l, classes = [], [], [], [], [], []
        for _ in range(n_ref * B):
            qry_img, q_label, spprt_imgs, s_label, subcls = next_item()[:5]
            iter_num += 1
            # fresh classifier init, drawn like the reference's nn.Conv2d(...) (a-1); the reference draws a second Conv2d per
            # episode (Pseudo_cls, src/test.py:200) whose init is overwritten — drawn too, to keep the RNG stream seed-for-seed
            w0_l.append(nn.Conv2d(C, args.num_classes_tr, kernel_size=1, bias=False).weight.detach().view(2, C))
            nn.Conv2d(C, args.num_classes_tr, kernel_size=1, bias=False)
            sp.append(spprt_imgs.squeeze(0)); sl.append(s_label.squeeze(0))
            qi.append(qry_img); ql.append(q_label.squeeze(0))
            classes.append([int(c.item()) if torch.is_tensor(c) else int(c) for c in subcls])
        S = sp[0].shape[0]
        with torch.cuda.stream(bb_stream):
            spprt = torch.cat(sp).to(device, non_blocking=True)                  # [B*S,3,H,W]
            qry = torch.cat(qi).to(device, non_blocking=True)                    # [B,3,H,W]
            s_lab = torch.stack(sl).to(device, non_blocking=True)
            q_lab = torch.stack(ql).to(device, non_blocking=True)
            w0 = torch.stack(w0_l).to(device, non_blocking=True)
            with torch.no_grad():
                f_s, _ = model.extract_features(spprt)                           # one call for all support images
                f_q, _ = model.extract_features(qry)                             # one call for all query images
            f_s = f_s.float().reshape(n_ref * B, S, *f_s.shape[1:]).contiguous()
            f_q = f_q.float().contiguous()
            ready = torch.cuda.Event()
            ready.record(bb_stream)
        return (f_s, s_lab, f_q, q_lab, w0), ready, classes, iter_num

    for run in range(args.n_runs):
        cls_I: Dict[int, int] = defaultdict(int)
        cls_U: Dict[int, int] = defaultdict(int)
        cls_I0: Dict[int, int] = defaultdict(int)
        cls_U0: Dict[int, int] = defaultdict(int)
        IoU: Dict[int, float] = {}
        IoU0: Dict[int, float] = {}
        loss_sum, loss_cnt, runtime = 0.0, 0, 0.0
        pending = None                       # results of the previous batch, still on their way to the host

        def finish(p):
            """Account one head launch, reference batch by reference batch: the reference's per-class accumulation
            (src/test.py:225-234), its per-batch loss and its progress line."""
            nonlocal loss_sum, loss_cnt
            counts, ce, status, ev, classes, n_seen = p
            ev.synchronize()
            n_ep = len(classes)
            ops.raise_for_status(status[:n_ep], first_episode=n_seen - n_ep)      # the reference's errors, one launch late
            for g0 in range(0, n_ep, B):
                cg = ce[g0:g0 + B]
                loss = float(cg[:, 0, 0].sum() / cg[:, 0, 1].sum().clamp_min(1.0))  # CE over the reference batch's valid pixels
                loss_sum += loss; loss_cnt += 1
                for i in range(g0, g0 + B):
                    for j, class_ in enumerate(classes[i]):
                        cls_I[class_] += int(counts[i, 0, j + 1, 0]); cls_U[class_] += int(counts[i, 0, j + 1, 1])
                        cls_I0[class_] += int(counts[i, 1, j + 1, 0]); cls_U0[class_] += int(counts[i, 1, j + 1, 1])
                for class_ in cls_U:
                    IoU[class_] = cls_I[class_] / (cls_U[class_] + 1e-10)
                    IoU0[class_] = cls_I0[class_] / (cls_U0[class_] + 1e-10)
                seen_g = n_seen - n_ep + g0 + B
                if verbose and seen_g % 200 == 0:
                    print('Test: [{}/{}] mIoU {:.4f} mIoU0 {:.4f} Loss {:.4f} ({:.4f}) '.format(
                        seen_g, args.test_num, np.mean(list(IoU.values())), np.mean(list(IoU0.values())),
                        loss, loss_sum / loss_cnt))

        # pinned result buffers, two sets (the launch being accounted and the launch in flight), allocated once per run
        EB = K * B
        host = [(torch.empty((EB, 2, 2, 3), dtype=torch.int64, pin_memory=True),
                 torch.empty((EB, 2, 2), dtype=torch.float64, pin_memory=True),
                 torch.empty((EB,), dtype=torch.int32, pin_memory=True)) for _ in range(2)]
        t0 = time.time()
        plan = [min(K, nb_episodes - i) for i in range(0, nb_episodes, K)]       # reference batches per head launch
        staged = stage_batch(plan[0]) if plan else None
        for e, n_ref in enumerate(plan):
            tensors, ready, classes, n_seen = staged
            # queue the NEXT launch's backbone before this launch's head (it runs on the backbone stream)
            staged = stage_batch(plan[e + 1]) if e + 1 < len(plan) else None
            main.wait_event(ready)
            for t in tensors:
                t.record_stream(main)
            f_s, s_lab, f_q, q_lab, w0 = tensors
            out = episode_head(f_s, s_lab, f_q, q_lab, w0, params, n_head, args.cls_lr, args.adapt_iter)
            # No device sync per batch (the reference has one per EPISODE, src/test.py:169-171): the counts travel to
            # pinned host memory asynchronously and are accounted one launch later — the GPU never waits for Python.
            n_ep = n_ref * B
            counts, ce, status = host[e % 2]
            counts[:n_ep].copy_(out.counts, non_blocking=True)
            ce[:n_ep].copy_(out.ce, non_blocking=True)
            status[:n_ep].copy_(out.status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            if pending is not None:
                finish(pending)
            pending = (counts, ce, status, ev, classes, n_seen)
        if pending is not None:
            finish(pending)
        runtime += time.time() - t0
        runtimes[run] = runtime
        mIoU = float(np.mean(list(IoU.values()))) if IoU else 0.0
        if verbose:
            print('mIoU---Val result: mIoU {:.4f}.'.format(mIoU))
            for class_ in cls_U:
                print("Class {} : {:.4f}".format(class_, IoU[class_]))
        val_IoUs[run] = mIoU
        val_losses[run] = loss_sum / max(loss_cnt, 1)
    if verbose:
        print('Average mIoU over {} runs --- {:.4f}.'.format(args.n_runs, val_IoUs.mean()))
        print('Average runtime / run --- {:.4f}.'.format(runtimes.mean()))
    return val_IoUs.mean(), val_losses.mean()


# ----------------------------------------------------------------------------------------
# meta-training step (src/train.py:233-267)
# ----------------------------------------------------------------------------------------
class _QueryLoss(torch.autograd.Function):
    """loss(W') = weighted CE( up( W' . normalize(f_q) ), q_label ) per episode, SUMMED over the batch's episodes
    (the reference trains with batch_size 1, config_files/pascal.yaml:25, where sum and mean coincide; with E > 1 the
    transformer gradient is the sum of the per-episode gradients)."""

    @staticmethod
    def forward(ctx, w_ad, f_q, q_label):
        E, Lq, C = w_ad.shape
        h, w = f_q.shape[-2:]
        logits = ops.rows_times_feat(w_ad.detach(), f_q, normalize=True).view(E, 2, h, w)
        loss, dl = ops.query_loss_grad(logits, q_label)
        ctx.save_for_backward(dl, f_q)
        ctx.mark_non_differentiable(logits)
        return loss.sum(), logits

    @staticmethod
    def backward(ctx, g_loss, _g_logits):
        dl, f_q = ctx.saved_tensors
        E = dl.shape[0]
        d_w = ops.feat_times_rows(dl.view(E, 2, -1), f_q, normalize=True)
        return d_w * g_loss, None, None


def query_loss(w_adapted: torch.Tensor, f_q: torch.Tensor, q_label: torch.Tensor):
    """Returns (loss, logits60 [E,2,h,w]); differentiable w.r.t. ``w_adapted``."""
    return _QueryLoss.apply(w_adapted, f_q, q_label)


def meta_train_step(transformer, optimizer, f_s, s_label, f_q, q_label, w0, cls_lr: float, adapt_iter: int,
                    keep_attn=None, keep_out=None):
    """One iteration of src/train.py:do_epoch on pre-computed features (batch of E episodes; E = 1 in the
    reference). Fits the classifier, adapts it with the transformer in train mode, takes the weighted
    query CE and steps ``optimizer`` (SGD momentum .9, nesterov, wd 1e-4: src/optimizer.py:11-15).
    Returns dict(loss, w_fit, w_adapted, logits60)."""
    w_fit = ops.fit_classifier(f_s, s_label, w0, cls_lr, adapt_iter, check=False)
    transformer.train()
    prev = transformer.normalize_k
    transformer.normalize_k = True
    try:
        updated = transformer(w_fit.detach(), f_q, f_q, keep_attn=keep_attn, keep_out=keep_out)
    finally:
        transformer.normalize_k = prev
    loss, logits = query_loss(updated, f_q, q_label)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return {"loss": loss.detach(), "w_fit": w_fit, "w_adapted": updated.detach(), "logits60": logits}


def do_epoch(args, train_loader, model, transformer, optimizer_trans, epoch: int, iter_per_epoch: int, log_iter: int,
             verbose: bool = True):
    """Drop-in for the reference's training epoch (src/train.py:166-290), same arguments and return value
    ``(train_Ious, train_losses)`` (two ``[log_iter]`` tensors). ``args`` needs image_size, bottleneck_dim,
    num_classes_tr, cls_lr, adapt_iter, batch_size (1 in the reference, config_files/pascal.yaml:25).

    Per iteration: the backbone (the caller's PyTorch module, ``model.extract_features``) produces f_s in train mode
    and f_q in eval mode exactly as the reference does; the classifier fit, the transformer forward / backward, the
    weighted query CE and both IoU read-outs run in the fused kernels. In the 1-shot case the reference feeds the
    support image twice (``expand(2, ...)``, so that train-mode BatchNorm sees a batch) and fits on BOTH feature maps
    (S = 2, pooled loss): with the backbone in train mode (src/train.py:183) ``PSPNet.bottleneck`` ends in
    ``nn.Dropout2d(p=args.dropout)`` (dropout 0.1 in pascal.yaml), so the two copies carry different channel masks and are
    NOT identical. The fit therefore runs on both copies; only a backbone without any active dropout module (where the
    two copies are bit-identical and the pooled-mean loss of two identical shots equals the 1-shot loss) takes the
    single-copy shortcut onto the on-chip kernel."""
    device = next(transformer.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("do_epoch (cwt_b200): the transformer must live on a CUDA device")
    C = args.bottleneck_dim
    bs = int(getattr(args, "batch_size", 1))
    if bs != 1:
        raise NotImplementedError("do_epoch keeps the reference's batch size / episode of 1 (src/train.py:194)")
    train_losses = torch.zeros(log_iter)
    train_Ious = torch.zeros(log_iter)
    train_Ious0 = torch.zeros(log_iter)
    loss_sum, loss_n = 0.0, 0
    it = iter(train_loader)
    model.train()
    transformer.train()
    active_dropout = any(isinstance(m, nn.modules.dropout._DropoutNd) and m.p > 0 for m in model.modules())
    for i in range(iter_per_epoch):
        try:
            item = next(it)
        except StopIteration:
            it = iter(train_loader)
            item = next(it)
        qry_img, q_label, spprt_imgs, s_label, subcls = item[:5]
        spprt_imgs, s_label = spprt_imgs.to(device), s_label.to(device)
        q_label, qry_img = q_label.to(device), qry_img.to(device)
        one_shot = spprt_imgs.shape[1] == 1
        if one_shot:
            spprt = spprt_imgs.squeeze(0).expand(2, 3, args.image_size, args.image_size)
        else:
            spprt = spprt_imgs.squeeze(0)
        w0 = nn.Conv2d(C, args.num_classes_tr, kernel_size=1, bias=False).weight.detach().view(1, 2, C).to(device)
        with torch.no_grad():
            f_s, _ = model.extract_features(spprt)                     # [n_support, c, h, w], backbone in train mode
        if one_shot and not active_dropout:
            f_s = f_s[:1]                                              # two bit-identical copies: fit on one
        elif one_shot:
            s_label = s_label.expand(1, 2, *s_label.shape[-2:])        # [1,2,H,W]: the label of both copies (src/train.py:201)
        model.eval()
        with torch.no_grad():
            f_q, _ = model.extract_features(qry_img)                   # [1, c, h, w]
        f_s, f_q = f_s.float().unsqueeze(0).contiguous(), f_q.float().contiguous()
        out = meta_train_step(transformer, optimizer_trans, f_s, s_label, f_q, q_label, w0, args.cls_lr, args.adapt_iter)
        # read-outs (src/train.py:269-279): adapted logits on normalised features, baseline logits on raw features
        l0 = ops.rows_times_feat(out["w_fit"], f_q, normalize=False).view(1, 2, f_q.shape[-2], f_q.shape[-1])
        counts, _ = ops.upsample_argmax_iou(torch.cat([out["logits60"], l0]), torch.cat([q_label, q_label]))
        counts = counts.cpu().double()
        iou = counts[:, :, 0] / (counts[:, :, 1] + 1e-10)               # [2 read-outs, 2 classes]
        loss_sum += float(out["loss"]) / bs
        loss_n += 1
        if i < log_iter:
            train_losses[i] = loss_sum / loss_n
            train_Ious[i] = float(iou[0].mean())
            train_Ious0[i] = float(iou[1].mean())
        if verbose and ((epoch == 0 and i % 100 == 0) or i % 500 == 0):
            print('iter {} IoUf {:.2f}, IoUb {:.2f}, IoUf0 {:.2f}, IoUb0 {:.2f}'.format(
                i, float(iou[0, 1]), float(iou[0, 0]), float(iou[1, 1]), float(iou[1, 0])))
        model.train()
    if verbose:
        print('Epoch {}: The mIoU {:.2f}, loss {:.2f}, mIoU0 {:.2f}'.format(
            epoch + 1, train_Ious.mean(), train_losses.mean(), train_Ious0.mean()))
    return train_Ious, train_losses


# ----------------------------------------------------------------------------------------
# host-resident episodes: double-buffered H2D staging in front of the head
# ----------------------------------------------------------------------------------------
def bind_host_to_gpu(device) -> Optional[List[int]]:
    """Pin this process to the CPUs that are local to ``device`` (NVML's ideal CPU affinity) so that the pinned staging
    buffers it allocates afterwards live on the GPU's NUMA node: with one process per GPU the H2D streams of a sharded
    sweep then do not cross the socket interconnect. Returns the CPU list, or None when NVML / the CPUs are unavailable
    (e.g. a container that only sees a slice of the host); never raises."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        idx = torch.device(device).index or 0
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            tok = vis.split(",")[idx].strip()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(tok)) if tok.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(tok)
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class HostPipeline:
    """Evaluate batches of episodes that live in (pinned) HOST memory.

    The H2D copy of batch i+1 runs on a side stream while the head of batch i computes; the small int64
    count tensor of every batch is copied back asynchronously into pinned memory. This is the call the
    end-to-end benchmark times (``bench.py`` ``e2e``): every input byte crosses PCIe inside the timed region.
    """

    def __init__(self, device, params: Dict[str, torch.Tensor], n_head: int, lr: float, n_iter: int,
                 fit_algo: int = L.FIT_AUTO, attn_algo: int = L.ATTN_REASSOC, num_classes_val: int = 5, n_slots: int = 4,
                 sub_batch: int = 32, sub_batch_all: bool = True, expand_on_main: int = 2, ramp=(8, 8, 16)):
        self.device = torch.device(device)
        self.params = {k: v.to(self.device) for k, v in params.items()}
        self.n_head, self.lr, self.n_iter = n_head, lr, n_iter
        self.fit_algo, self.attn_algo = fit_algo, attn_algo
        self.copy_stream = torch.cuda.Stream(self.device)
        self.table = IoUTable(num_classes_val, self.device)
        # Host batches go through in SUB-BATCHES of ``sub_batch`` episodes (``sub_batch_all`` False: only the first batch of a run),
        # the first batch starting with the ``ramp`` sizes: the head can only start when its first piece has arrived, so the
        # exposed copy at the start of a run is that of 8 episodes. Sizing: a zero-compressed episode is 8.3 MB = 0.157 ms of
        # H2D at 53 GB/s, its fit 0.178 ms — the copy is barely faster than the fit, so a piece's copy only hides under the
        # previous piece's fit if the pieces stop growing: whole 64-episode batches behind a doubling ramp stall 8 ms per run
        # (round-2 default while the fit still took 0.27 ms per episode), uniform pieces do not.
        # Measured on B200 (bench.py e2e, 20 steps of 64 episodes, profiles/r2g_e2e_policies.txt): pieces of 32 after a ramp
        # 8, 8, 16: 12.10 ms per step (0.957 of the device-resident rate); whole batches after 8, 8, 16, 32: 12.44; pieces of
        # 16: 12.38; of 8: 13.05 (launch overheads); expansion on the copy stream instead of the head's: 13.4.
        # Device-resident staging slots are reused for the whole run: no allocator traffic (a cudaMalloc of ~1 GB
        # synchronises the device). At least three slots: the copy of sub-batch i+2 may start as soon as the copy of i+1 has
        # finished (its slot was released by head i-1 long ago), so the copy engine never waits for the head.
        # Zero-compressed host batches: where the expansion kernel runs (``expand_on_main``).
        # 1: on the head's stream, right in front of the fit (~0.25 ms per 64 episodes, exposed).
        # 0: on the copy stream, behind the copies it expands, i.e. concurrently with the previous piece's cooperative fit —
        #    whose 144 CTAs leave it the four free SMs: the expansion of 64 episodes then takes 5-7 ms instead of 0.25, the next
        #    piece's copy queues behind it and the copy engine idles (tools/e2e_timeline.py): 13.4 ms per step against 12.38.
        # 2 (default): on a THIRD stream behind the copy it expands — the copy stream carries nothing but copies, the expansion
        #    of piece i+1 runs on the SMs the fit of piece i leaves free, and with FOUR staging slots copy, expansion and fit of
        #    three consecutive pieces overlap. Measured on B200 (profiles/r2i_e2e_policies.txt, 20 steps of 64 episodes,
        #    device-resident rate 11.45 ms per step): 11.76 ms per step (0.974) against 11.98 (0.956) for mode 1 with three slots;
        #    pieces of 16 with five slots 11.71 (0.978).
        self.expand_on_main = int(expand_on_main)
        self.expand_stream = torch.cuda.Stream(self.device) if self.expand_on_main == 2 else None
        self.ramp = tuple(int(x) for x in (ramp or ()))     # sub-batch sizes at the very start of a run (first host batch only)
        self.sub_batch = max(0, int(sub_batch))
        self.sub_batch_all = bool(sub_batch_all)            # sub-batch every host batch, not only the first
        self._slots: List[Optional[EpisodeBatch]] = [None] * n_slots
        self._slot_free: List[Optional[torch.cuda.Event]] = [None] * n_slots
        self._results: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = []   # pinned (counts, status) per host batch
        self._cslots: List[Optional[dict]] = [None] * n_slots   # device staging of zero-compressed features (mask, prefix, values)
        self._head: Optional[HeadPipeline] = None

    @staticmethod
    def _fields(b: EpisodeBatch):
        return (b.f_s, b.s_label, b.f_q, b.q_label, b.w0, b.subcls, b.idx)

    def _stage(self, hb, lo: int, hi: int, slot: int, cap: int = 0):
        """H2D copy of episodes ``lo:hi`` of one host batch into staging slot ``slot`` (capacity >= ``cap`` episodes) on the
        copy stream. ``hb`` is an :class:`EpisodeBatch` (dense pinned tensors) or a :class:`CompressedEpisodeBatch`
        (zero-compressed features: mask / prefix / packed values are copied and the dense tensors are rebuilt on the device
        by ``cwt_expand_zero_compressed_f32``, on the copy stream too)."""
        n = hi - lo
        pending = []                  # expansions left for the head's stream (expand_on_main == 1)
        side_jobs = []                # expansions for the dedicated stream (expand_on_main == 2)
        compressed = isinstance(hb, CompressedEpisodeBatch)
        shapes = [(tuple(hb.f_s.shape[1:]), torch.float32), (tuple(hb.s_label.shape[1:]), hb.s_label.dtype),
                  (tuple(hb.f_q.shape[1:]), torch.float32), (tuple(hb.q_label.shape[1:]), hb.q_label.dtype),
                  (tuple(hb.w0.shape[1:]), hb.w0.dtype), (tuple(hb.subcls.shape[1:]), hb.subcls.dtype),
                  (tuple(hb.idx.shape[1:]), hb.idx.dtype)]
        with torch.cuda.stream(self.copy_stream):
            full = self._slots[slot]
            if full is None or not (all(tuple(x.shape[1:]) == sh and x.dtype == dt for x, (sh, dt) in zip(self._fields(full), shapes))
                                    and full.f_s.shape[0] >= n):
                cap = max(n, cap, 1)
                full = EpisodeBatch(*(torch.empty((cap,) + sh, dtype=dt, device=self.device) for sh, dt in shapes))
                for t in self._fields(full):
                    t.record_stream(self._main)                       # read by the head's stream
                self._slots[slot] = full
                self._cslots[slot] = None
            elif self._slot_free[slot] is not None:
                self.copy_stream.wait_event(self._slot_free[slot])      # the head that read this slot last has finished
            db = EpisodeBatch(*(t.narrow(0, 0, n) for t in self._fields(full)))
            if not compressed:
                for dst, src in zip(self._fields(db), self._fields(hb)):
                    dst.copy_(src[lo:hi], non_blocking=True)
            else:
                for dst, src in ((db.s_label, hb.s_label), (db.q_label, hb.q_label), (db.w0, hb.w0), (db.subcls, hb.subcls),
                                 (db.idx, hb.idx)):
                    dst.copy_(src[lo:hi], non_blocking=True)
                cs = self._cslots[slot] or {}
                for name, cm, dense in (("f_s", hb.f_s, db.f_s), ("f_q", hb.f_q, db.f_q)):
                    W = cm.mask.shape[1]
                    v0, v1 = cm.val_start[lo], cm.val_start[hi]
                    st = cs.get(name)
                    if st is None or st[0].shape[0] < n or st[0].shape[1] != W or st[2].numel() < v1 - v0:
                        capn = max(n, full.f_s.shape[0])
                        st = (torch.empty((capn, W), dtype=torch.int32, device=self.device),
                              torch.empty((capn, cm.woff.shape[1]), dtype=torch.int32, device=self.device),
                              torch.empty(max(int(1.25 * (v1 - v0)) + 1024, 1), dtype=torch.float32, device=self.device))
                        cs[name] = st
                    m, o, v = st[0][:n], st[1][:n], st[2][:max(v1 - v0, 1)]
                    m.copy_(cm.mask[lo:hi], non_blocking=True)
                    o.copy_(cm.woff[lo:hi], non_blocking=True)
                    if v1 > v0:
                        v[:v1 - v0].copy_(cm.vals[v0:v1], non_blocking=True)
                    if self.expand_on_main == 1:
                        pending.append((m, o, v, dense, v0))
                    elif self.expand_on_main == 2:
                        side_jobs.append((m, o, v, dense, v0))
                    else:
                        expand_map(m, o, v, dense, v0)
                self._cslots[slot] = cs
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        if side_jobs:
            # the expansions on their own stream, behind this piece's copies; the head waits for THEIR event
            self.expand_stream.wait_event(ev)
            with torch.cuda.stream(self.expand_stream):
                for m, o, v, dense, v0 in side_jobs:
                    for t in (m, o, v, dense):
                        t.record_stream(self.expand_stream)
                    expand_map(m, o, v, dense, v0)
                ev = torch.cuda.Event()
                ev.record(self.expand_stream)
        return db, ev, pending

    def _sub_batches(self, host_batches):
        """(batch number, first episode, end episode, is-last, host batch, E): slices of a pinned tensor along dim 0 stay pinned.
        The FIRST host batch of a run goes through in the ``ramp`` sizes (default 8, 8, 16, then ``sub_batch``): the head can
        only start when its first piece has arrived, so the exposed copy at the start of a run is that of 8 episodes."""
        for bi, hb in enumerate(host_batches):
            E = hb.n_episodes
            step = self.sub_batch if (self.sub_batch > 0 and (bi == 0 or self.sub_batch_all)) else max(E, 1)
            sizes = list(self.ramp) if (bi == 0 and self.ramp) else []
            if not sizes and 0 < step < E:
                # equal pieces: 36 episodes in pieces of at most 32 are 18 + 18, not 32 + 4 (a 4-episode launch leaves most
                # groups of the persistent fit idle)
                n_pieces = (E + step - 1) // step
                step = (E + n_pieces - 1) // n_pieces
            lo = 0
            while lo < max(E, 1):
                n = sizes.pop(0) if sizes else step
                hi = min(E, lo + max(1, n))
                yield bi, lo, hi, hi >= E, hb, E
                lo = max(hi, lo + 1)

    def _result_buffers(self, bi: int, E: int):
        """Pinned host buffers of host batch ``bi`` (counts int64 [E,2,2,3], fit status int32 [E]): owned by the pipeline and
        reused by every run — no pinned allocation (a cudaHostAlloc synchronises) inside the loop after the first run."""
        while len(self._results) <= bi:
            self._results.append(None)
        cur = self._results[bi]
        if cur is None or cur[0].shape[0] < E:
            cur = (torch.empty((max(E, 1), 2, 2, 3), dtype=torch.int64, pin_memory=True),
                   torch.empty((max(E, 1),), dtype=torch.int32, pin_memory=True))
            self._results[bi] = cur
        return cur[0][:E], cur[1][:E]

    def run(self, host_batches: Iterable[EpisodeBatch], reduce_every_step: bool = False, reduce: bool = True,
            check: bool = True) -> List[torch.Tensor]:
        """Returns the per-batch count tensors (CPU, int64 [E,2,2,3]; views of pipeline-owned pinned buffers, valid until the
        next ``run``); ``self.table`` accumulates the sweep metrics. ``reduce_every_step``: all-reduce the table after
        every host batch (every rank must then see the same number of batches); ``reduce``: all-reduce once at the end.
        ``check``: raise the reference's errors (empty support mask -> ZeroDivisionError, bad labels -> ValueError,
        non-finite fit -> FloatingPointError) from the status words that came back with the counts — after the run, so
        the pipeline itself never waits for the host."""
        main = torch.cuda.current_stream(self.device)
        self._main = main
        it = self._sub_batches(host_batches)
        results: List[torch.Tensor] = []
        statuses: List[torch.Tensor] = []
        n_slots = len(self._slots)
        queue = []                          # staged sub-batches, oldest first
        nslot = 0

        def prefetch():
            nonlocal nslot
            while len(queue) < n_slots - 1:                   # one slot is being read by the head
                item = next(it, None)
                if item is None:
                    return
                bi, lo, hi, last, hb, E = item
                db_, ev_, pend_ = self._stage(hb, lo, hi, nslot, E)
                queue.append((db_, ev_, pend_, nslot, bi, lo, last, E))
                nslot = (nslot + 1) % n_slots

        # ONE HeadPipeline (one side stream) for the life of this object: the library workspaces are keyed by stream, so a new
        # side stream per run() meant a fresh cudaMalloc of every post-stage workspace inside the run (measured with
        # tools/e2e_timeline.py: the first post stage of every run took 10 ms instead of 0.4)
        if self._head is None:
            self._head = HeadPipeline(self.device, self.params, self.n_head, self.lr, self.n_iter, self.fit_algo, self.attn_algo,
                                      table=self.table)
        head = self._head

        def read_back(out, dst, dst_status):
            dst.copy_(out.counts, non_blocking=True)
            dst_status.copy_(out.status, non_blocking=True)

        prefetch()
        while queue:
            db, ev, pend, cur, bi, lo, last, E = queue.pop(0)
            main.wait_event(ev)
            for m_, o_, v_, dense_, v0_ in pend:              # zero-compressed features: rebuild the dense tensors in front of the fit
                expand_map(m_, o_, v_, dense_, v0_)
            if bi == len(results):                            # first sub-batch of host batch bi: its pinned result buffers
                r, st = self._result_buffers(bi, E)
                results.append(r); statuses.append(st)
            dst = results[bi][lo:lo + db.n_episodes]
            dst_status = statuses[bi][lo:lo + db.n_episodes]
            # fit on the main stream; transformer + logits/IoU, the table update, the (optional) all-reduce and the D2H read
            # of the step's result on the head's side stream, overlapping the NEXT sub-batch's fit
            _, done = head.submit(db.f_s, db.s_label, db.f_q, db.q_label, db.w0, db.subcls,
                                  reduce=bool(last and reduce_every_step),
                                  after=lambda out, dst=dst, dst_status=dst_status: read_back(out, dst, dst_status))
            self._slot_free[cur] = done                       # the slot may be overwritten once its post stage has finished
            prefetch()                                        # H2D of the following sub-batches overlaps this and the next head
        head.finish()
        if reduce and not reduce_every_step:
            self.table.all_reduce()
        torch.cuda.current_stream(self.device).synchronize()
        if check:
            first = 0
            for st in statuses:
                ops.raise_for_status(st, first_episode=first)
                first += st.shape[0]
        return results
