"""The ``gather_logit`` implementation group of the reference's ``DependencyBoxRel``
(/root/reference/src/model/joint.py:406-432), computed by the tcgen05 kernel in ``csrc/align_kernels.cu``.

Two call styles:

* functional -- ``gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask)`` etc. on plain tensors;
* drop-in    -- ``gather_logit_simple_impl(self, inputs, vis, txt, vp)`` with the reference's method signature
  (``vis = (feat, mask, split)``, ``txt = (feat, mask, marginal)``, named tensors), suitable for
  ``JointModelBase.add_impl_to_group("gather_logit", "simple")`` (joint.py:110, base.py:118-142); see INTEGRATION.md.

The result carries the names ``("B", "A", "Q", "V")`` because the consumers use ``.max("V")``, ``.log_softmax("A")``
and ``align_to`` (joint.py:473-483, 520-524); it is a fresh, writable tensor (the loss mutates it in place).
"""
from __future__ import annotations

import warnings

import torch

from ._lib import VlgaeError, check, lib

INF = 1e20  # reference src/__init__.py:110, bound at import by joint.py:16 (quirk Q5)

_ws = {}


def _plain(t):
    return t.rename(None) if t is not None and any(n is not None for n in t.names) else t


def _workspace(dev, nbytes):
    # one scratch buffer per (device, stream): calls on different streams must not share the packed operand tiles
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _ws[key] = ws
    return ws


def _launch(vf, vm, tf, tm, split, neg, pad_rows):
    """One kernel call on prepared tensors (fp32 contiguous features, uint8 masks) -> the [B, A, Q, ldv] buffer with
    ldv = V rounded up to 8 floats when pad_rows (elements >= V of a row are not written); the caller slices ``[..., :V]``."""
    dev = vf.device
    A, V, D = vf.shape
    B, Q, _ = tf.shape
    ldv = (V + 7) // 8 * 8 if pad_rows else V
    out = torch.empty((B, A, Q, ldv), dtype=torch.float32, device=dev)
    need = lib().vlgae_align_workspace_bytes(A, V, B, Q, D)
    if need == 0 and out.numel() > 0:
        raise VlgaeError(f"gather_logit: unsupported shape (D = {D} > 128?)")
    ws = _workspace(dev, max(need, 1))
    with torch.cuda.device(dev):
        check(lib().vlgae_align_logits(vf.data_ptr(), vm.data_ptr(), tf.data_ptr(), tm.data_ptr(), A, V, B, Q, D,
                                       float(neg), int(split), out.data_ptr(), ldv, ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream(dev).cuda_stream), "vlgae_align_logits")
    return out


class _AlignLogits(torch.autograd.Function):
    """The reference's attmap is an autograd node (einsum + two masked_fill_, joint.py:413-418): the grounding losses
    back-propagate through it into both encoders.  Forward = the tcgen05 kernel; backward = the two transposed
    contractions, also on tcgen05 (vlgae_align_logits_backward, csrc/align_bwd_kernels.cu), with the masks folded in (a
    masked entry was overwritten, so it passes no gradient):
        d vis[a,v,:] = m_v[a,v] * sum_{b,q} g[b,a,q,v] * (m_q[b,q] txt[b,q,:])
        d txt[b,q,:] = m_q[b,q] * sum_{a,v} g[b,a,q,v] * (m_v[a,v] vis[a,v,:])"""

    @staticmethod
    def forward(ctx, vis_feat, txt_feat, vm, tm, split, neg, pad_rows):
        vf = vis_feat.detach().to(torch.float32).contiguous()
        tf = txt_feat.detach().to(torch.float32).contiguous()
        ctx.save_for_backward(vf, tf, vm, tm)
        ctx.in_dtypes = (vis_feat.dtype, txt_feat.dtype)
        ctx.split = int(split)
        # the padded buffer itself is the output: a ``[..., :V]`` view taken in here would be marked as a view created
        # inside a custom Function, and the reference's in-place write into attmap (joint.py:466-469, 548-551) would raise
        return _launch(vf, vm, tf, tm, split, neg, pad_rows)

    @staticmethod
    def backward(ctx, g):
        vf, tf, vm, tm = ctx.saved_tensors
        B, A, Q = g.shape[:3]
        V, D = vf.shape[1], vf.shape[2]
        dev = g.device
        g = g.to(torch.float32)
        if g.stride(-1) != 1 or g.stride(2) < V or g.stride(1) != Q * g.stride(2) or g.stride(0) != A * Q * g.stride(2):
            g = g.contiguous()  # rows may be padded (stride >= V); anything else is made dense
        gv = torch.empty((A, V, D), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        gt = torch.empty((B, Q, D), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        ws = _workspace(dev, max(lib().vlgae_align_workspace_bytes(A, V, B, Q, D), 1))
        with torch.cuda.device(dev):
            check(lib().vlgae_align_logits_backward(g.data_ptr(), g.stride(2), vf.data_ptr(), vm.data_ptr(), tf.data_ptr(),
                                                    tm.data_ptr(), A, V, B, Q, D, ctx.split,
                                                    gv.data_ptr() if gv is not None else None,
                                                    gt.data_ptr() if gt is not None else None, ws.data_ptr(), ws.numel(),
                                                    torch.cuda.current_stream(dev).cuda_stream),
                  "vlgae_align_logits_backward")
        if gv is not None:
            gv = gv.to(ctx.in_dtypes[0])
        if gt is not None:
            gt = gt.to(ctx.in_dtypes[1])
        return gv, gt, None, None, None, None, None


def gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask, *, split=3, neg=-INF, named=True, pad_rows=True):
    """attmap [B, A, Q, V] = <txt[b,q,:], vis[a,v,:]> with both masks applied (joint.py:406-419); differentiable with
    respect to both feature tensors, like the reference's einsum.

    pad_rows: allocate rows of ``ceil(V / 8) * 8`` floats and return the ``[..., :V]`` view (same shape, dtype and
    values; last-dim stride 1, not contiguous).  Every consumer in joint.py indexes / reduces / sorts, none calls
    ``.view``; sector-aligned rows roughly halve the time of this write-bound operator.  ``pad_rows=False`` gives the
    reference's dense layout.
    """
    vis_feat, vis_mask, txt_feat, txt_mask = map(_plain, (vis_feat, vis_mask, txt_feat, txt_mask))
    dev = vis_feat.device
    if dev.type != "cuda":
        raise VlgaeError("vlgae_b200.alignment needs CUDA tensors (there is no CPU fallback)")
    A, V, D = vis_feat.shape
    B, Q, D2 = txt_feat.shape
    if D != D2 or tuple(vis_mask.shape) != (A, V) or tuple(txt_mask.shape) != (B, Q):
        raise VlgaeError("gather_logit: vis [A,V,D], vis_mask [A,V], txt [B,Q,D], txt_mask [B,Q] expected")
    vm = vis_mask.to(torch.bool).contiguous().view(torch.uint8)
    tm = txt_mask.to(torch.bool).contiguous().view(torch.uint8)
    if torch.is_grad_enabled() and (vis_feat.requires_grad or txt_feat.requires_grad):
        out = _AlignLogits.apply(vis_feat, txt_feat, vm, tm, split, neg, pad_rows)
    else:
        out = _launch(vis_feat.detach().to(torch.float32).contiguous(), vm,
                      txt_feat.detach().to(torch.float32).contiguous(), tm, split, neg, pad_rows)
    if out.shape[-1] != V:
        out = out[..., :V]  # an ordinary view of the Function's output: in-place writes are allowed
    if named:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = out.refine_names("B", "A", "Q", "V")
    return out


def max_over_factors(vis_feat, vis_mask, txt_feat, txt_mask, *, split=3, neg=-INF, want_argmax=True):
    """(maxatt [B, A, Q], argmax_v [B, A, Q] int32) of the alignment logits without materialising them
    (vlgae_align_max_over_factors); maxatt is bit-identical to ``gather_logit_simple(...).max(-1).values``."""
    vf, vm, tf, tm = map(_plain, (vis_feat, vis_mask, txt_feat, txt_mask))
    dev = vf.device
    if dev.type != "cuda":
        raise VlgaeError("vlgae_b200.alignment needs CUDA tensors (there is no CPU fallback)")
    A, V, D = vf.shape
    B, Q, D2 = tf.shape
    if D != D2 or tuple(vm.shape) != (A, V) or tuple(tm.shape) != (B, Q):
        raise VlgaeError("gather_logit: vis [A,V,D], vis_mask [A,V], txt [B,Q,D], txt_mask [B,Q] expected")
    vf = vf.detach().to(torch.float32).contiguous()
    tf = tf.detach().to(torch.float32).contiguous()
    vm = vm.to(torch.bool).contiguous().view(torch.uint8)
    tm = tm.to(torch.bool).contiguous().view(torch.uint8)
    maxv = torch.empty((B, A, Q), dtype=torch.float32, device=dev)
    argv = torch.empty((B, A, Q), dtype=torch.int32, device=dev) if want_argmax else None
    need = lib().vlgae_align_reduce_workspace_bytes(A, V, B, Q, D)
    if need == 0 and maxv.numel() > 0:
        raise VlgaeError(f"gather_logit: unsupported shape (D = {D} > 128?)")
    ws = _workspace(dev, max(need, 1))
    with torch.cuda.device(dev):
        check(lib().vlgae_align_max_over_factors(vf.data_ptr(), vm.data_ptr(), tf.data_ptr(), tm.data_ptr(), A, V, B, Q, D,
                                                 float(neg), int(split), maxv.data_ptr(),
                                                 argv.data_ptr() if argv is not None else None, ws.data_ptr(), ws.numel(),
                                                 torch.cuda.current_stream(dev).cuda_stream),
              "vlgae_align_max_over_factors")
    return maxv, argv


class _MaxOverFactors(torch.autograd.Function):
    """max over V of the alignment logits as one autograd node.  torch's backward of ``attmap.max(-1)`` routes the
    gradient of (b, a, q) to the single arg-max factor; through the einsum that is
        d txt[b,q,:] += g[b,a,q] * vis[a, argv[b,a,q], :]        d vis[a, argv[b,a,q], :] += g[b,a,q] * txt[b,q,:]
    and nothing where the arg-max entry was a masked one (masked_fill_ cut the graph there)."""

    @staticmethod
    def forward(ctx, vis_feat, txt_feat, vis_mask, txt_mask, split, neg):
        maxv, argv = max_over_factors(vis_feat, vis_mask, txt_feat, txt_mask, split=split, neg=neg)
        ctx.save_for_backward(vis_feat.detach(), txt_feat.detach(), vis_mask, txt_mask, argv)
        ctx.mark_non_differentiable(argv)
        return maxv, argv

    @staticmethod
    def backward(ctx, g, _g_arg):
        vis, txt, vm, tm, argv = ctx.saved_tensors
        B, A, Q = g.shape
        V, D = vis.shape[1], vis.shape[2]
        arg = argv.long()
        a_idx = torch.arange(A, device=g.device).view(1, A, 1).expand(B, A, Q)
        keep = vm.bool()[a_idx, arg] & tm.bool().view(B, 1, Q)
        g = (g * keep).to(torch.float32)
        gv = gt = None
        if ctx.needs_input_grad[1]:
            sel = vis.to(torch.float32)[a_idx, arg]                       # [B, A, Q, D]
            gt = torch.einsum("baq,baqd->bqd", g, sel).to(txt.dtype)
        if ctx.needs_input_grad[0]:
            src = (g.unsqueeze(-1) * txt.to(torch.float32).unsqueeze(1)).reshape(-1, D)   # [B*A*Q, D]
            flat = (a_idx * V + arg).reshape(-1)
            gv = torch.zeros(A * V, D, dtype=torch.float32, device=g.device).index_add_(0, flat, src)
            gv = gv.view(A, V, D).to(vis.dtype)
        return gv, gt, None, None, None, None


def gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal, *, split=3):
    """max over V, marginal-weighted mean over Q -> [B, A] (joint.py:421-432).  The [B, A, Q, V] tensor is never
    written: the kernel's epilogue reduces over the factors; differentiable w.r.t. both feature tensors and the
    marginal, like the reference."""
    vis_feat, vis_mask, txt_feat, txt_mask = map(_plain, (vis_feat, vis_mask, txt_feat, txt_mask))
    maxatt, _ = _MaxOverFactors.apply(vis_feat, txt_feat, vis_mask, txt_mask, split, -INF)
    tm = _plain(txt_marginal)
    return torch.sum(maxatt * tm.unsqueeze(1), dim=-1) / tm.sum(1, keepdim=True)


def word_factor_attention(vis_feat, txt_feat, vis_mid):
    """Per-sample word -> factor attention of ``DependencyBoxRel._forward`` (joint.py:668-673):
    ``softmax_v(<txt[b,q,:], vis[b,v,:]>) @ vis_mid[b]`` -> [B, Q, H].  Two plain batched GEMMs and a softmax (0.3 % of the
    alignment contraction's flops at the cfg2 shape): run as library calls, differentiable through torch."""
    vis_feat, txt_feat, vis_mid = map(_plain, (vis_feat, txt_feat, vis_mid))
    if vis_feat.device.type != "cuda":
        raise VlgaeError("vlgae_b200.alignment needs CUDA tensors (there is no CPU fallback)")
    att = torch.bmm(txt_feat, vis_feat.transpose(1, 2)).softmax(2)
    return torch.bmm(att, vis_mid)


# ---- drop-in methods for the reference's implementation-group registry -------------------------------------------
def gather_logit_simple_impl(self, inputs, vis, txt, vp):
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, _txt_marginal = txt
    return gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask)


def gather_logit_reduced_impl(self, inputs, vis, txt, vp):
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, txt_marginal = txt
    return gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal)
