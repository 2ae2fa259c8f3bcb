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
    ws = _ws.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _ws[dev] = ws
    return ws


def gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask, *, split=3, neg=-INF, named=True, pad_rows=True):
    """attmap [B, A, Q, V] = <txt[b,q,:], vis[a,v,:]> with both masks applied (joint.py:406-419).

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
    vf = vis_feat.detach().to(torch.float32).contiguous()
    tf = txt_feat.detach().to(torch.float32).contiguous()
    vm = vis_mask.to(torch.bool).contiguous().view(torch.uint8)
    tm = txt_mask.to(torch.bool).contiguous().view(torch.uint8)
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
    if ldv != V:
        out = out[..., :V]
    if named:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = out.refine_names("B", "A", "Q", "V")
    return out


def gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal, *, split=3):
    """max over V, marginal-weighted mean over Q -> [B, A] (joint.py:421-432)."""
    att = gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask, split=split, named=False)
    maxatt = att.max(dim=-1).values
    tm = _plain(txt_marginal)
    return torch.sum(maxatt * tm.unsqueeze(1), dim=-1) / tm.sum(1, keepdim=True)


# ---- drop-in methods for the reference's implementation-group registry -------------------------------------------
def gather_logit_simple_impl(self, inputs, vis, txt, vp):
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, _txt_marginal = txt
    return gather_logit_simple(vis_feat, vis_mask, txt_feat, txt_mask)


def gather_logit_reduced_impl(self, inputs, vis, txt, vp):
    vis_feat, vis_mask, _ = vis
    txt_feat, txt_mask, txt_marginal = txt
    return gather_logit_reduced(vis_feat, vis_mask, txt_feat, txt_mask, txt_marginal)
