"""Construction of the DMV score tensors: the step of ``DiscriminativeNDMV._forward`` right before the
chart (/root/reference/src/model/ldndmv.py:184-209), fused with ``DMV1o.merge``.

The caller keeps its modules (``mid_ff``, the three ``DMVFactorizedBilinear`` scorers); it hands over the PROJECTED operands
of the attach scorer (the outputs of ``project1`` / ``project2``, dmv_spec.py:69-70) instead of calling the scorer's
einsum, so the [B, n, n_token, 2, 2] rule tensor and its log-softmax over the vocabulary are never materialised::

    x1 = self.attach_scorer.project1(h_parent)                       # [B, n, 2, 2, r]
    x2 = self.attach_scorer.project2(h_child)[0]                     # [n_token, 2, 2, r]
    dec_score = self.dec_scorer(h_parent, h_dec)                     # [B, n, 2, 2, 2]  (small: library einsum)
    root_score = self.root_scorer(h_root, h_child).sum([-1, -2]).reshape(-1)   # [n_token]
    merged_dec, merged_attach = dmv_scores(x1, x2, inputs["token"], dec_score, root_score, head_mask=in_mask)

``out["attach"]``, ``out["dec"]`` and ``out["root"]`` are views of the merged tensors (``split_merged``).  Differentiable
w.r.t. x1, x2, dec_score and root_score: the backward takes the gradients of the merged tensors -- the chart's marginals
-- and recomputes the softmax over the vocabulary in two streaming passes.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from ._lib import VlgaeError, check, lib
from .torch_struct.dmv import NOCHILD

INF = 1e20       # /root/reference/src/__init__.py:110
NEGINF = -1e12   # import-time default of DMV1o.merge's `zero` (torch_struct/semirings/semirings.py:16)

_ws = {}


def _workspace(dev, need):
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws[key] = ws
    return ws


class _DmvScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, dec_score, root_score, token, head_mask, one, zero, neg):
        dev = x1.device
        B, n, _, _, r = x1.shape
        T = x2.shape[0]
        N = n + 1
        f = [t.detach().to(torch.float32).contiguous() for t in (x1, x2, dec_score, root_score)]
        token = token.to(device=dev, dtype=torch.int64).contiguous()
        hm = head_mask.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8) if head_mask is not None else None
        md = torch.empty((B, N, 2, 2, 2), dtype=torch.float32, device=dev)
        ma = torch.empty((B, N, N, 2), dtype=torch.float32, device=dev)
        lse = torch.empty((B, n, 2, 2), dtype=torch.float32, device=dev)
        root_lse = torch.empty((1,), dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib().vlgae_dmv_scores_workspace_bytes(B, n))
        with torch.cuda.device(dev):
            check(lib().vlgae_dmv_scores(f[0].data_ptr(), f[1].data_ptr(), token.data_ptr(), hm.data_ptr() if hm is not None else None,
                                         f[2].data_ptr(), f[3].data_ptr(), B, n, T, r, float(one), float(zero), float(neg),
                                         md.data_ptr(), ma.data_ptr(), lse.data_ptr(), root_lse.data_ptr(), ws.data_ptr(),
                                         ws.numel(), torch.cuda.current_stream(dev).cuda_stream), "vlgae_dmv_scores")
        ctx.save_for_backward(*f, token, lse, root_lse, *([hm] if hm is not None else []))
        ctx.has_mask = hm is not None
        ctx.dtypes = (x1.dtype, x2.dtype, dec_score.dtype, root_score.dtype)
        return md, ma

    @staticmethod
    def backward(ctx, gmd, gma):
        saved = ctx.saved_tensors
        x1, x2, dec_score, root_score, token, lse, root_lse = saved[:7]
        hm = saved[7] if ctx.has_mask else None
        dev = x1.device
        B, n, _, _, r = x1.shape
        T = x2.shape[0]
        gmd = (gmd if gmd is not None else torch.zeros((B, n + 1, 2, 2, 2), device=dev)).to(torch.float32).contiguous()
        gma = (gma if gma is not None else torch.zeros((B, n + 1, n + 1, 2), device=dev)).to(torch.float32).contiguous()
        g1, g2, gd, gr = (torch.empty_like(t) for t in (x1, x2, dec_score, root_score))
        ws = _workspace(dev, lib().vlgae_dmv_scores_workspace_bytes(B, n))
        with torch.cuda.device(dev):
            check(lib().vlgae_dmv_scores_backward(
                x1.data_ptr(), x2.data_ptr(), token.data_ptr(), hm.data_ptr() if hm is not None else None, dec_score.data_ptr(),
                root_score.data_ptr(), lse.data_ptr(), root_lse.data_ptr(), gmd.data_ptr(), gma.data_ptr(), B, n, T, r,
                g1.data_ptr(), g2.data_ptr(), gd.data_ptr(), gr.data_ptr(), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(dev).cuda_stream), "vlgae_dmv_scores_backward")
        grads = tuple(g.to(dt) for g, dt in zip((g1, g2, gd, gr), ctx.dtypes))
        return grads + (None, None, None, None, None)


def dmv_scores(x1, x2, token, dec_score, root_score, head_mask=None, *, extended_valence=True, one=0.0, zero=NEGINF, neg=-INF):
    """(merged_dec [B, n+1, 2, 2, 2], merged_attach [B, n+1, n+1, 2]) of ldndmv.py:184-209 (see the module docstring).

    x1 [B, n, 2, 2, r], x2 [n_token, 2, 2, r] (or [1, n_token, 2, 2, r]), token [B, n] int64, dec_score [B, n, 2, 2, 2]
    (decision-major, as ``dec_scorer`` returns it), root_score [n_token]; ``head_mask`` [B, n] bool = the reference's
    ``in_mask`` (``cfg.function_mask``).  r in {4, 8, 16, 32}."""
    if not extended_valence:
        raise VlgaeError("dmv_scores: extended_valence=False (ldndmv.py:186-187) is not provided; both shipped configs set it true")
    if x1.device.type != "cuda":
        raise VlgaeError("vlgae_b200.scores needs CUDA tensors (there is no CPU fallback)")
    if x2.dim() == 5 and x2.shape[0] == 1:
        x2 = x2[0]
    root_score = root_score.reshape(-1)
    B, n = token.shape
    if x1.dim() != 5 or tuple(x1.shape[:4]) != (B, n, 2, 2) or x2.dim() != 4 or tuple(x2.shape[1:]) != (2, 2, x1.shape[4]) \
            or tuple(dec_score.shape) != (B, n, 2, 2, 2) or root_score.shape[0] != x2.shape[0]:
        raise VlgaeError("dmv_scores: x1 [B,n,2,2,r], x2 [T,2,2,r], token [B,n], dec_score [B,n,2,2,2], root_score [T] expected")
    if head_mask is not None:
        head_mask = head_mask.reshape(B, n)
    return _DmvScores.apply(x1, x2, dec_score, root_score, token, head_mask, one, zero, neg)


def split_merged(merged_dec, merged_attach):
    """(dec, attach, root) = the reference's ``out['dec']``, ``out['attach']``, ``out['root']`` as views of the merged
    tensors (the inverse of distributions.py:253-265)."""
    return merged_dec[:, 1:], merged_attach[:, 1:, 1:, :], merged_attach[:, 0, 1:, NOCHILD]
