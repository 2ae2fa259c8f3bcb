"""Visual factor features with the relation MLP collapsed.

Mirrors ``VisBoxRelSimpleEncoder.forward`` (/root/reference/src/model/vis_encoder/box_rel.py:31-54) followed by
``DependencyBoxRel.vis_feat_unprune`` (/root/reference/src/model/joint.py:140-179).  The reference runs
``rel_fc = LeakyReLU(Linear(.))`` on the n^2 pairwise means of the box inputs; the Linear is affine, so the pre-activation of
pair (i, j) is the mean of the per-box pre-activations -- one pass of n rows through the 4096 x 256 matrix instead of n^2,
and the [B, n, n, 4096] pair tensor is never formed.  The three per-box Linears stay plain library GEMMs (they ARE plain
GEMMs); the pairwise expansion + LeakyReLU + concatenation + factor mask is one kernel (``vlgae_vis_factors``), its backward
folds the n^2 pair gradients back onto the n boxes (``vlgae_vis_factors_backward``).  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from ._lib import VlgaeError, check, lib


class _VisFactors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u_box, u_rel, u_attr, box_mask, has_img, slope):
        dev = u_box.device
        B, n, H = u_box.shape
        ub, ur = u_box.detach().to(torch.float32).contiguous(), u_rel.detach().to(torch.float32).contiguous()
        ua = u_attr.detach().to(torch.float32).contiguous() if u_attr is not None else None
        bm = box_mask.to(device=dev, dtype=torch.bool).contiguous().view(torch.uint8)
        V = n + n * n + (n if ua is not None else 0) + (1 if has_img else 0)
        mid = torch.empty((B, V, H), dtype=torch.float32, device=dev)
        mask = torch.empty((B, V), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().vlgae_vis_factors(ub.data_ptr(), ur.data_ptr(), ua.data_ptr() if ua is not None else None, bm.data_ptr(),
                                          B, n, H, int(has_img), float(slope), mid.data_ptr(), mask.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream), "vlgae_vis_factors")
        ctx.save_for_backward(ub, ur, *([ua] if ua is not None else []))
        ctx.has_attr, ctx.has_img, ctx.slope = ua is not None, bool(has_img), float(slope)
        ctx.dtypes = (u_box.dtype, u_rel.dtype, u_attr.dtype if u_attr is not None else None)
        ctx.mark_non_differentiable(mask)
        return mid, mask

    @staticmethod
    def backward(ctx, g_mid, _g_mask):
        saved = ctx.saved_tensors
        ub, ur = saved[0], saved[1]
        ua = saved[2] if ctx.has_attr else None
        dev = ub.device
        B, n, H = ub.shape
        g = g_mid.to(torch.float32).contiguous()
        gb, gr = torch.empty_like(ub), torch.empty_like(ur)
        ga = torch.empty_like(ua) if ua is not None else None
        with torch.cuda.device(dev):
            check(lib().vlgae_vis_factors_backward(ub.data_ptr(), ur.data_ptr(), ua.data_ptr() if ua is not None else None,
                                                   g.data_ptr(), B, n, H, int(ctx.has_img), ctx.slope, gb.data_ptr(), gr.data_ptr(),
                                                   ga.data_ptr() if ga is not None else None,
                                                   torch.cuda.current_stream(dev).cuda_stream), "vlgae_vis_factors_backward")
        return (gb.to(ctx.dtypes[0]), gr.to(ctx.dtypes[1]), ga.to(ctx.dtypes[2]) if ga is not None else None, None, None, None)


def pairwise_factors(u_box, u_rel, u_attr, box_mask, *, add_image=True, slope=0.01):
    """(mid [B, V, H], vis_mask [B, V] bool, split) from the per-box PRE-activations ``u_* = mlp.linear(inputs)`` [B, n, H]
    (``u_attr`` may be None).  Factor order and mask as ``vis_feat_unprune``: box (n) | rel (n^2, pair (i, j) at i n + j,
    mask = outer product of the box mask, strictly upper triangle) | attr (n) | img (mean of the box features, mask 1)."""
    if u_box.device.type != "cuda":
        raise VlgaeError("vlgae_b200.vis_factors needs CUDA tensors (there is no CPU fallback)")
    if u_box.dim() != 3 or u_rel.shape != u_box.shape or (u_attr is not None and u_attr.shape != u_box.shape) \
            or tuple(box_mask.shape) != tuple(u_box.shape[:2]):
        raise VlgaeError("pairwise_factors: u_box, u_rel, u_attr [B, n, H] and box_mask [B, n] expected")
    n = u_box.shape[1]
    mid, mask = _VisFactors.apply(u_box, u_rel, u_attr, box_mask, bool(add_image), float(slope))
    split = [n, n * n] + ([n] if u_attr is not None else []) + ([1] if add_image else [])
    return mid, mask.view(torch.bool), split


def vis_feat_unprune_collapsed(vis_encoder, vis_mlp_pre_matching, vis_box_feat, vis_box_mask, *, add_image=True, return_mid=False):
    """Drop-in for ``self.vis_encoder(x, ctx)`` + ``vis_feat_unprune`` with ``add_rel`` (and ``add_attr`` iff the encoder has
    ``attr_fc``): takes the reference's own modules -- ``box_fc`` / ``rel_fc`` / ``attr_fc`` MLPs (nn/common.py:23-51: Linear,
    LeakyReLU, dropout 0 in config/model/vlgae.yaml:31) and the bias-free ``vis_mlp_pre_matching`` -- and returns
    ``(vis [A, V, Y], vis_mask [A, V], split[, mid])`` as named tensors like the reference."""
    feat = vis_box_feat
    if getattr(vis_encoder, "img_feat", False):  # box_rel.py:35-40
        inputs = torch.cat([feat, feat.mean(1, keepdim=True).expand(-1, feat.shape[1], -1)], dim=-1)
    else:
        inputs = feat
    for name in ("box_fc", "rel_fc"):
        mlp = getattr(vis_encoder, name)
        if not isinstance(mlp.dropout, torch.nn.Identity) and mlp.training:
            raise VlgaeError("vis_feat_unprune_collapsed: dropout inside the visual MLPs is not reproduced (config has dropout 0)")
    slope = float(getattr(vis_encoder.rel_fc.activation, "negative_slope", 1.0))  # nn.Identity when activate: false
    u_box, u_rel = vis_encoder.box_fc.linear(inputs), vis_encoder.rel_fc.linear(inputs)
    u_attr = vis_encoder.attr_fc.linear(inputs) if getattr(vis_encoder, "use_attr", False) else None
    mid, mask, split = pairwise_factors(u_box, u_rel, u_attr, vis_box_mask, add_image=add_image, slope=slope)
    vis = vis_mlp_pre_matching(mid).refine_names("A", "V", "Y")
    mask = mask.refine_names("A", "V")
    return (vis, mask, split, mid) if return_mid else (vis, mask, split)
