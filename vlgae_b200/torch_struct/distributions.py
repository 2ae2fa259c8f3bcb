"""``DMV1o`` / ``DependencyCRF`` with the reference's operator API
(/root/reference/src/model/torch_struct/distributions.py:25-299), computed by libvlgae_b200.so.

API contract kept from the reference:
  * ``DMV1o.merge(dec, attach, root, one=0, zero=NEGINF)`` -> float32 ``(dec_wroot, attach_wroot)``, differentiable;
  * ``DMV1o([dec, attach], lengths)`` with lazily cached ``.partition`` / ``.max`` of shape ``[B, 1]`` that are
    differentiable w.r.t. BOTH inputs, and ``.argmax`` / ``.marginals`` of shape ``[B, N, N, 2]``;
  * ``DependencyCRF(arc [B,N,N], lengths).partition/.max -> [B]``, ``.argmax/.marginals -> [B, N, N]``.

Unlike the reference, ``partition`` does not record an autograd graph through the chart: when an input requires
grad the kernel runs the explicit reverse sweep in the same launch and the backward is a row scaling
(d sum_b g_b Z_b / d theta = g_b * dZ_b/d theta).  Double backward through the chart (``create_graph=True`` in
helpers.py:152) is not supported; no caller in src/model uses it.
"""
from __future__ import annotations

import torch
from torch.distributions.utils import lazy_property

from .. import ops
from .dmv import NOCHILD
from .semirings import LogSemiring, MaxSemiring
from .semirings.semirings import NEGINF


class _Merge(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, attach, root, one, zero):
        ctx.in_dtypes = (dec.dtype, attach.dtype, root.dtype)
        return ops.dmv_merge(dec, attach, root, one, zero)

    @staticmethod
    def backward(ctx, gdec_w, gattach_w):
        dd, da, dr = ctx.in_dtypes
        gdec = gdec_w[:, 1:].to(dd) if ctx.needs_input_grad[0] else None
        gatt = gattach_w[:, 1:, 1:, :].to(da) if ctx.needs_input_grad[1] else None
        groot = gattach_w[:, 0, 1:, NOCHILD].to(dr) if ctx.needs_input_grad[2] else None
        return gdec, gatt, groot, None, None


class _Partition(torch.autograd.Function):
    """log Z; saves dZ/d(dec, attach) computed by the in-kernel reverse sweep."""

    @staticmethod
    def forward(ctx, dec, attach, lengths):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        Z, gdec, gatt = ops.dmv_inside_outside(dec, attach, lengths, want_grad=need)
        if need:
            ctx.save_for_backward(gdec, gatt)
        ctx.in_dtypes = (dec.dtype, attach.dtype)
        return Z.unsqueeze(-1)

    @staticmethod
    def backward(ctx, gZ):
        gdec, gatt = ctx.saved_tensors
        dd, da = ctx.in_dtypes
        out_dec = ops.scale_rows(gdec, gZ).to(dd) if ctx.needs_input_grad[0] else None
        out_att = ops.scale_rows(gatt, gZ).to(da) if ctx.needs_input_grad[1] else None
        return out_dec, out_att, None


class _Max(torch.autograd.Function):
    """best tree score; saves the 0/1 arc indicator and decision counts of the best tree."""

    @staticmethod
    def forward(ctx, dec, attach, lengths):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        best, _, arcs, gdec = ops.dmv_viterbi(dec, attach, lengths, want_heads=False, want_arcs=need, want_gdec=need)
        if need:
            ctx.save_for_backward(gdec, arcs)
        ctx.in_dtypes = (dec.dtype, attach.dtype)
        return best.unsqueeze(-1)

    @staticmethod
    def backward(ctx, g):
        gdec, arcs = ctx.saved_tensors
        dd, da = ctx.in_dtypes
        out_dec = ops.scale_rows(gdec, g).to(dd) if ctx.needs_input_grad[0] else None
        out_att = ops.scale_rows(arcs, g).to(da) if ctx.needs_input_grad[1] else None
        return out_dec, out_att, None


class StructDistribution:
    """Base class: holds potentials + lengths, exposes the lazily cached quantities the models read."""

    struct = None

    def __init__(self, log_potentials, lengths=None, args={}):
        self.log_potentials = log_potentials
        self.lengths = lengths
        self.args = args
        anchor = log_potentials[0] if isinstance(log_potentials, (list, tuple)) else log_potentials
        self.batch_shape = anchor.shape[:1]
        self.event_shape = anchor.shape[1:]

    def _unsupported(self, what):
        raise NotImplementedError(
            f"{type(self).__name__}.{what}: not on VLGAE's hot path (no caller in src/model); "
            "only partition / max / argmax / marginals are provided by vlgae_b200")

    def entropy(self):
        self._unsupported("entropy")

    def sample(self, sample_shape=torch.Size()):
        self._unsupported("sample")

    def kmax(self, k):
        self._unsupported("kmax")

    def topk(self, k):
        self._unsupported("topk")


class DMV1o(StructDistribution):
    """First-order DMV over merged (ROOT-prefixed) ``[dec, attach]`` score tensors."""

    def __init__(self, log_potentials, lengths, args={}):
        super().__init__(log_potentials[0], lengths=lengths, args=args)
        self.log_potentials = log_potentials

    @staticmethod
    def merge(dec, attach, root, one=0, zero=NEGINF):
        return _Merge.apply(dec, attach, root, float(one), float(zero))

    @lazy_property
    def partition(self):
        dec, attach = self.log_potentials
        return _Partition.apply(dec, attach, self.lengths)

    @lazy_property
    def max(self):
        dec, attach = self.log_potentials
        return _Max.apply(dec, attach, self.lengths)

    @lazy_property
    def argmax(self):
        dec, attach = self.log_potentials
        return ops.dmv_viterbi(dec, attach, self.lengths, want_heads=False, want_arcs=True)[2]

    @lazy_property
    def marginals(self):
        dec, attach = self.log_potentials
        return ops.dmv_inside_outside(dec, attach, self.lengths, want_grad=True)[2]

    # extras (not in the reference API): what callers otherwise rebuild from argmax.sum(-1).nonzero()
    @lazy_property
    def heads(self):
        """[B, N] int64: heads[b, c] = head of word c (ROOT = 0); column 0 and padding are 0."""
        dec, attach = self.log_potentials
        return ops.dmv_viterbi(dec, attach, self.lengths, want_heads=True, want_arcs=False)[1]


class DependencyCRF(StructDistribution):
    """Arc-factored projective CRF used for MBR decoding (reference distributions.py:269-299)."""

    def __init__(self, log_potentials, lengths=None, args={}, multiroot=False):
        assert not multiroot  # deptree.py:27
        if log_potentials.dim() not in (3, 4):
            raise ValueError("potentials must have dim of 3 (unlabeled) or 4 (labeled)")
        super().__init__(log_potentials, lengths, args)

    def _run(self, semiring, want_marg):
        from .. import deptree

        if torch.is_grad_enabled() and self.log_potentials.requires_grad:
            # the reference's partition / max are autograd-connected; here they are decode-only (the one caller,
            # ldndmv.py:294-299, runs under no_grad on detached marginals) -- refuse rather than detach silently
            raise NotImplementedError(
                "DependencyCRF: potentials that require grad are not supported (MBR decoding only); "
                "detach them or call under torch.no_grad()")
        return deptree.run(self.log_potentials, self.lengths, semiring, want_marg)

    @lazy_property
    def partition(self):
        return self._run(LogSemiring, False)[0]

    @lazy_property
    def max(self):
        return self._run(MaxSemiring, False)[0]

    @lazy_property
    def argmax(self):
        return self._run(MaxSemiring, True)[1]

    @lazy_property
    def marginals(self):
        return self._run(LogSemiring, True)[1]
