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
        dev = g.device
        vf, tf = vis.to(torch.float32).contiguous(), txt.to(torch.float32).contiguous()
        vmu = vm.to(torch.bool).contiguous().view(torch.uint8)
        tmu = tm.to(torch.bool).contiguous().view(torch.uint8)
        g = g.to(torch.float32).contiguous()
        gv = torch.empty((A, V, D), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        gt = torch.empty((B, Q, D), dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(dev):
            check(lib().vlgae_align_max_over_factors_backward(
                g.data_ptr(), argv.data_ptr(), vf.data_ptr(), vmu.data_ptr(), tf.data_ptr(), tmu.data_ptr(), A, V, B, Q, D,
                gv.data_ptr() if gv is not None else None, gt.data_ptr() if gt is not None else None,
                torch.cuda.current_stream(dev).cuda_stream), "vlgae_align_max_over_factors_backward")
        if gv is not None:
            gv = gv.to(vis.dtype)
        if gt is not None:
            gt = gt.to(txt.dtype)
        return gv, gt, None, None, None, None


def gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal, *, split=3):
    """max over V, marginal-weighted mean over Q -> [B, A] (joint.py:421-432).  The [B, A, Q, V] tensor is never
    written: the kernel's epilogue reduces over the factors; differentiable w.r.t. both feature tensors and the
    marginal, like the reference."""
    vis_feat, vis_mask, txt_feat, txt_mask = map(_plain, (vis_feat, vis_mask, txt_feat, txt_mask))
    maxatt, _ = _MaxOverFactors.apply(vis_feat, txt_feat, vis_mask, txt_mask, split, -INF)
    tm = _plain(txt_marginal)
    return torch.sum(maxatt * tm.unsqueeze(1), dim=-1) / tm.sum(1, keepdim=True)


# ---- fused grounding consumers: loss and decode without the [B, A, Q, V] tensor ------------------
def _prep4(vis_feat, vis_mask, txt_feat, txt_mask):
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
    return dev, A, V, B, Q, D, vf, vm, tf, tm


def fused_maxima(vis_feat, vis_mask, txt_feat, txt_mask, *, split=3, neg=-INF):
    """(maxv [B,A,Q], argv, maxq [B,A,V], argq): ``attmap.max("V")`` and ``attmap.max("Q")`` (joint.py:473, 480, 520) in ONE
    pass of the tcgen05 kernel (vlgae_align_maxima); values bit-identical to the maxima of the materialised logits."""
    dev, A, V, B, Q, D, vf, vm, tf, tm = _prep4(vis_feat, vis_mask, txt_feat, txt_mask)
    maxv = torch.empty((B, A, Q), dtype=torch.float32, device=dev)
    argv = torch.empty((B, A, Q), dtype=torch.int32, device=dev)
    maxq = torch.empty((B, A, V), dtype=torch.float32, device=dev)
    argq = torch.empty((B, A, V), dtype=torch.int32, device=dev)
    ws = _workspace(dev, max(lib().vlgae_align_reduce_workspace_bytes(A, V, B, Q, D), 1))
    with torch.cuda.device(dev):
        check(lib().vlgae_align_maxima(vf.data_ptr(), vm.data_ptr(), tf.data_ptr(), tm.data_ptr(), A, V, B, Q, D, float(neg),
                                       int(split), maxv.data_ptr(), argv.data_ptr(), maxq.data_ptr(), argq.data_ptr(),
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream),
              "vlgae_align_maxima")
    return maxv, argv, maxq, argq


def diagonal_slab(vis_feat, vis_mask, txt_feat, txt_mask, *, neg=-INF):
    """attmap[b, b] for every caption -> [B, Q, V] (exact fp32; joint.py:466-469, 522-524).  Needs A == B."""
    dev, A, V, B, Q, D, vf, vm, tf, tm = _prep4(vis_feat, vis_mask, txt_feat, txt_mask)
    if A != B:
        raise VlgaeError("diagonal_slab: one image per caption expected (A == B)")
    out = torch.empty((B, Q, V), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().vlgae_align_diagonal(vf.data_ptr(), vm.data_ptr(), tf.data_ptr(), tm.data_ptr(), B, V, Q, D, float(neg),
                                         out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "vlgae_align_diagonal")
    return out


def grounding_ce(maxv, maxq, txt_marginal, vis_mask):
    """(txt2vis, vis2txt) of loss_grounding_factor_ce (joint.py:473-483) from the two maxima; vis2txt is None without maxq."""
    dev = maxv.device
    B, A, Q = maxv.shape
    if A != B:
        raise VlgaeError("grounding_ce: A == B expected")
    maxv = maxv.to(torch.float32).contiguous()
    marg = _plain(txt_marginal).to(device=dev, dtype=torch.float32).contiguous()
    V = 0
    vm = None
    if maxq is not None:
        maxq = maxq.to(torch.float32).contiguous()
        V = maxq.shape[2]
        vm = _plain(vis_mask).to(torch.bool).contiguous().view(torch.uint8)
    out2 = torch.empty(2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().vlgae_grounding_ce(maxv.data_ptr(), maxq.data_ptr() if maxq is not None else None, marg.data_ptr(),
                                       vm.data_ptr() if vm is not None else None, B, Q, V, out2.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "vlgae_grounding_ce")
    return out2[0], (out2[1] if maxq is not None else None)


def topk_rows(x, k=5):
    """Indices of the k largest entries along the last dim, descending (``x.argsort(-1, descending=True)[..., :k]``,
    joint.py:594); ties: smaller index first."""
    if x.device.type != "cuda":
        raise VlgaeError("vlgae_b200.alignment needs CUDA tensors (there is no CPU fallback)")
    x = x.to(torch.float32).contiguous()
    V = x.shape[-1]
    rows = x.numel() // max(V, 1)
    idx = torch.empty(x.shape[:-1] + (k,), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().vlgae_topk_rows(x.data_ptr(), rows, V, k, idx.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream),
              "vlgae_topk_rows")
    return idx.long()


class _Max:
    def __init__(self, values, indices):
        self.values, self.indices = values, indices


class FusedMatch:
    """What ``inputs["match_logit"]`` holds under ``gather_logit_mode: b200_fused``: the two maxima and the diagonal slab
    instead of the [B, A, Q, V] tensor.  It answers the three accesses the reference's ``decode_grounding_on_factor``
    makes (joint.py:520-524) -- ``.max("V")``, ``.max("Q")`` and ``.diagonal()`` -- so that method runs on it UNMODIFIED;
    ``loss_grounding_factor_ce_fused_impl`` is the matching loss.  Forward only (validation / test / bulk decode): training
    keeps the materialised, differentiable ``gather_logit_mode: b200``."""

    names = ("B", "A", "Q", "V")

    def __init__(self, vis_feat, vis_mask, txt_feat, txt_mask):
        self.maxv, self.argv, self.maxq, self.argq = fused_maxima(vis_feat, vis_mask, txt_feat, txt_mask)
        self.diag = diagonal_slab(vis_feat, vis_mask, txt_feat, txt_mask)
        B, A, Q = self.maxv.shape
        self.shape = (B, A, Q, self.maxq.shape[2])

    def max(self, dim):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if dim in ("V", -1, 3):
                return _Max(self.maxv.refine_names("B", "A", "Q"), self.argv.long().refine_names("B", "A", "Q"))
            if dim in ("Q", 2):
                return _Max(self.maxq.refine_names("B", "A", "V"), self.argq.long().refine_names("B", "A", "V"))
        raise VlgaeError(f"FusedMatch.max: only the V and Q axes are reduced by the fused kernel, got {dim!r}")

    def diagonal(self):
        return self.diag.permute(1, 2, 0)  # [Q, V, B], as torch's diagonal() of a [B, A, Q, V] tensor

    def __len__(self):
        return self.shape[0]


def grounding_loss_fused(match, txt_marginal, vis_mask, num_token, *, prior=None, vis2txt=1.0):
    """The two cross-entropy terms of ``loss_grounding_factor_ce`` (joint.py:439-491) from a FusedMatch.
    prior: optional list of ``(mask [B, T, 1] bool, lo, hi)``: the POS prior lowers, on the diagonal a = b only, the scores of
    the words flagged by ``mask`` outside the factor group ``[lo, hi)`` by 100 (joint.py:446-470)."""
    maxv, maxq = match.maxv, match.maxq
    if prior:
        diag = match.diag.clone()
        for mask, lo, hi in prior:
            T = mask.shape[1]
            m = mask.to(diag.dtype) * 100
            diag[:, 1:T + 1, :lo] -= m
            diag[:, 1:T + 1, hi:] -= m
        ar = torch.arange(diag.shape[0], device=diag.device)
        maxv, maxq = maxv.clone(), maxq.clone()
        maxv[ar, ar] = diag.max(-1).values   # [B, Q]
        maxq[ar, ar] = diag.max(1).values    # [B, V]
    t2v, v2t = grounding_ce(maxv, maxq if vis2txt > 0 else None, txt_marginal, vis_mask)
    loss = {"txt2vis": t2v / (t2v.detach() + 1e-6) * num_token}
    if vis2txt > 0:
        loss["mt_vis2txt"] = vis2txt * v2t / (v2t.detach() + 1e-6) * num_token
    return sum(loss.values()), loss, (t2v, v2t)


class _WordAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vis, txt, mid):
        dev = vis.device
        B, V, D = vis.shape
        n, H = txt.shape[1], mid.shape[2]
        vf, tf, mf = (x.detach().to(torch.float32).contiguous() for x in (vis, txt, mid))
        out = torch.empty((B, n, H), dtype=torch.float32, device=dev)
        lse = torch.empty((B, n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib().vlgae_word_attention(vf.data_ptr(), tf.data_ptr(), mf.data_ptr(), B, V, n, D, H, out.data_ptr(),
                                             lse.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "vlgae_word_attention")
        ctx.save_for_backward(vf, tf, mf, out, lse)
        ctx.dtypes = (vis.dtype, txt.dtype, mid.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        vf, tf, mf, out, lse = ctx.saved_tensors
        dev = g.device
        B, V, D = vf.shape
        n, H = tf.shape[1], mf.shape[2]
        g = g.to(torch.float32).contiguous()
        need = ctx.needs_input_grad
        gv = torch.empty_like(vf) if need[0] else None
        gt = torch.empty_like(tf) if need[1] else None
        gm = torch.empty_like(mf) if need[2] else None
        if n > 0:
            with torch.cuda.device(dev):
                check(lib().vlgae_word_attention_backward(
                    vf.data_ptr(), tf.data_ptr(), mf.data_ptr(), out.data_ptr(), lse.data_ptr(), g.data_ptr(), B, V, n, D, H,
                    gv.data_ptr() if gv is not None else None, gt.data_ptr() if gt is not None else None,
                    gm.data_ptr() if gm is not None else None, torch.cuda.current_stream(dev).cuda_stream),
                    "vlgae_word_attention_backward")
        else:
            gv, gm = (x.zero_() if x is not None else None for x in (gv, gm))
        return tuple(x.to(dt) if x is not None else None for x, dt in zip((gv, gt, gm), ctx.dtypes))


def word_factor_attention(vis_feat, txt_feat, vis_mid):
    """Per-sample word -> factor attention of ``DependencyBoxRel._forward`` (joint.py:668-673):
    ``softmax_v(<txt[b,q,:], vis[b,v,:]>) @ vis_mid[b]`` -> [B, n, H] in one kernel (``vlgae_word_attention``: online softmax
    over tiles of factors, the [B, n, V] map is never written); differentiable w.r.t. all three inputs through
    ``vlgae_word_attention_backward``.  n <= 64 words, H <= 256."""
    vis_feat, txt_feat, vis_mid = map(_plain, (vis_feat, txt_feat, vis_mid))
    if vis_feat.device.type != "cuda":
        raise VlgaeError("vlgae_b200.alignment needs CUDA tensors (there is no CPU fallback)")
    if vis_feat.dim() != 3 or txt_feat.dim() != 3 or vis_mid.dim() != 3 or vis_feat.shape[0] != txt_feat.shape[0] \
            or vis_mid.shape[:2] != vis_feat.shape[:2] or vis_feat.shape[2] != txt_feat.shape[2]:
        raise VlgaeError("word_factor_attention: vis [B,V,D], txt [B,n,D], vis_mid [B,V,H] expected")
    return _WordAttention.apply(vis_feat, txt_feat, vis_mid)


# ---- drop-in methods for the reference's implementation-group registry -------------------------------------------
def gather_logit_simple_impl(self, inputs, vis, txt, vp):
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, _txt_marginal = txt
    return gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask)


def gather_logit_reduced_impl(self, inputs, vis, txt, vp):
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, txt_marginal = txt
    return gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal)


def gather_logit_fused_impl(self, inputs, vis, txt, vp):
    """``gather_logit_mode: b200_fused`` -- the consumers' view of attmap without the tensor (forward only)."""
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, _txt_marginal = txt
    return FusedMatch(vis_feat, vis_mask, txt_feat, txt_mask)


def loss_grounding_factor_ce_fused_impl(self, inputs, vp):
    """``loss_grounding_mode: factor|ce`` on a FusedMatch (mirrors joint.py:439-491, POS prior included)."""
    match = inputs["match_logit"]
    _txt_feat, _txt_mask, txt_marginal = inputs["txt_packed"]
    _vis_feat, vis_mask, vis_split = inputs["vis_packed"]
    prior = None
    if self.cfg.loss_grounding_args.use_pos_prior:
        prior, offset = [], 0
        pos = {"obj": getattr(self, "pos_for_obj", None), "rel": getattr(self, "pos_for_rel", None),
               "attr": getattr(self, "pos_for_attr", None)}
        for name, width in zip(self.vis_factor_names, vis_split):
            if name in pos:
                mask = vp.tag.unsqueeze(-1).eq(pos[name]).any(-1, keepdim=True)
                prior.append((mask, offset, offset + width))
            offset += width
    total, loss, _ = grounding_loss_fused(match, txt_marginal, vis_mask, vp.num_token, prior=prior,
                                          vis2txt=self.cfg.loss_grounding_args.vis2txt)
    return total, loss
