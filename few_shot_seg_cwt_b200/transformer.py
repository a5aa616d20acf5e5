"""Drop-in for the reference's ``MultiHeadAttentionOne`` (src/model/transformer.py:33-83).

Same constructor, same parameter / state-dict names and shapes (``w_qkvs.weight [nH*C,C]``,
``layer_norm.weight/bias [C]``, ``fc.weight [C,nH*C]``, ``fc.bias [C]``) so
``load_state_dict(checkpoint['state_dict'])`` (src/test.py:87-88) and
``get_optimizer(args, transformer.parameters())`` (src/train.py:98) keep working; ``.train()`` /
``.eval()`` toggle the two dropouts; the forward is an autograd Function over the CUDA kernels
so the meta-training step of src/train.py:253-267 back-propagates into the five parameters.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class _MHAOneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, w_qkvs, fc_w, fc_b, ln_g, ln_b, n_head, normalize_k, keep_attn, keep_out,
                p_attn, p_out, algo):
        need = any(t.requires_grad for t in (w_qkvs, fc_w, fc_b, ln_g, ln_b))
        res = ops.transformer_forward(q.detach(), k.detach(), w_qkvs.detach(), fc_w.detach(), fc_b.detach(),
                                      ln_g.detach(), ln_b.detach(), n_head, normalize_k, keep_attn, keep_out,
                                      p_attn, p_out, need_saved=need, algo=algo)
        if need:
            out, saved = res
            ctx.save_for_backward(q, k, w_qkvs, fc_w, ln_g, saved,
                                  keep_attn if keep_attn is not None else torch.empty(0),
                                  keep_out if keep_out is not None else torch.empty(0))
            ctx.cfg = (n_head, normalize_k, keep_attn is not None, keep_out is not None, p_attn, p_out)
        else:
            out = res
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, w_qkvs, fc_w, ln_g, saved, ka, ko = ctx.saved_tensors
        n_head, normalize_k, has_ka, has_ko, p_attn, p_out = ctx.cfg
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            raise NotImplementedError(
                "MultiHeadAttentionOne (cwt_b200): gradients w.r.t. q / k are not provided — on the reference "
                "path the classifier weights enter as .data and the features under no_grad (src/train.py:246-253)")
        g = ops.transformer_backward(d_out, q.detach(), k.detach(), w_qkvs.detach(), fc_w.detach(), ln_g.detach(),
                                     n_head, saved, normalize_k, ka if has_ka else None, ko if has_ko else None,
                                     p_attn, p_out)
        d_wqkvs, d_fcw, d_fcb, d_g, d_b = g
        return (None, None, d_wqkvs, d_fcw, d_fcb, d_g, d_b, None, None, None, None, None, None, None)


class MultiHeadAttentionOne(nn.Module):
    """Multi-Head Attention module with shared projection (B200 kernels behind it)."""

    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1, attn_dropout=0.1, algo: int = L.ATTN_REASSOC):
        super().__init__()
        if not (d_model == d_k == d_v):
            raise NotImplementedError("cwt_b200 MultiHeadAttentionOne supports d_model == d_k == d_v "
                                      "(the reference always builds it with 512/512/512: src/test.py:57)")
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qkvs = nn.Linear(d_model, n_head * d_k, bias=False)
        nn.init.normal_(self.w_qkvs.weight, mean=0, std=np.sqrt(2.0 / (d_model + d_k)))      # transformer.py:45
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)                                               # transformer.py:51
        self.p_out = float(dropout)                # nn.Dropout(dropout) on fc output, transformer.py:52,80
        self.p_attn = float(attn_dropout)          # ScaledDotProductAttention attn_dropout=0.1, transformer.py:17-20
        self.algo = algo
        self.normalize_k = False                   # set True to fuse F.normalize(f_q, dim=1) (src/test.py:194)

    def draw_masks(self, B, Lq, HW, C, device, generator=None):
        """Bernoulli keep-masks for the two dropouts (train mode). Head-major layout
        [n_head*B, Lq, HW] like the reference's permute(2,0,1,3) batching (transformer.py:71-75)."""
        ka = (torch.rand(self.n_head * B, Lq, HW, device=device, generator=generator) >= self.p_attn).to(torch.uint8)
        ko = (torch.rand(B, Lq, C, device=device, generator=generator) >= self.p_out).to(torch.uint8)
        return ka, ko

    def forward(self, q, k, v=None, query_input=False, keep_attn=None, keep_out=None):
        """q [B,Lq,C]; k, v [B,C,h,w] -> [B,Lq,C].  ``v`` must be ``k`` (the only call pattern on the
        reference path: transformer(W, f_q, f_q))."""
        if v is not None and v is not k and not (v.data_ptr() == k.data_ptr() and v.shape == k.shape):
            # Distinct storage cannot be proven equal without a comparison pass and a host sync, which the hot path must not
            # pay: refuse it. (The reference only ever calls transformer(W, f_q, f_q), src/test.py:197, src/train.py:257.)
            raise NotImplementedError("cwt_b200 MultiHeadAttentionOne requires v to be k (shared K/V projection: pass the "
                                      "same tensor twice, as the reference does: transformer(W, f_q, f_q))")
        if self.training:
            # each dropout draws its own mask unless the caller supplied it (a caller that passes only one mask still gets
            # dropout on the other path, like nn.Dropout)
            B, Lq, C = q.shape
            HW = k.shape[2] * (k.shape[3] if k.dim() == 4 else 1)
            if keep_attn is None and self.p_attn > 0.0:
                keep_attn = (torch.rand(self.n_head * B, Lq, HW, device=q.device) >= self.p_attn).to(torch.uint8)
            if keep_out is None and self.p_out > 0.0:
                keep_out = (torch.rand(B, Lq, C, device=q.device) >= self.p_out).to(torch.uint8)
        return _MHAOneFunction.apply(q, k, self.w_qkvs.weight, self.fc.weight, self.fc.bias,
                                     self.layer_norm.weight, self.layer_norm.bias, self.n_head,
                                     self.normalize_k, keep_attn, keep_out, self.p_attn, self.p_out, self.algo)
