"""Tensor-level wrappers over the C ABI: validate, allocate outputs, pass device pointers + stream.

PyTorch is plumbing here (device memory, streams); all arithmetic happens in libvlgae_b200.so.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import VlgaeError, check, lib

# class attribute `zero` of the reference's _BaseLog, frozen at import (semirings.py:16,128): the value the
# single-root mask writes (dmv.py:63) even after src.setup_inf() rebinds the module global.
MASK_ZERO = -1e12

_workspaces = {}


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    if dev.type != "cuda":
        raise VlgaeError("vlgae_b200 operators need CUDA tensors (there is no CPU fallback)")
    for t in tensors:
        if t.device != dev:
            raise VlgaeError("all tensors must be on the same CUDA device")
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _workspace(dev: torch.device, B: int, N: int) -> Tuple[Optional[torch.Tensor], int]:
    need = lib().vlgae_dmv_workspace_bytes(B, N)
    if need == 0:
        return None, 0
    # one scratch buffer per (device, stream): launches on different streams must not share chart storage
    key = (dev, _stream(dev))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws, ws.numel()


def _prep(dec: torch.Tensor, attach: torch.Tensor, lengths: torch.Tensor):
    dev = _require_cuda(dec, attach)
    if dec.dim() != 5 or tuple(dec.shape[2:]) != (2, 2, 2):
        raise VlgaeError(f"dec must be [B, N, 2, 2, 2], got {tuple(dec.shape)}")
    B, N = dec.shape[:2]
    if tuple(attach.shape) != (B, N, N, 2):
        raise VlgaeError(f"attach must be [B, N, N, 2] = {(B, N, N, 2)}, got {tuple(attach.shape)}")
    dec = dec.detach().to(torch.float32).contiguous()
    attach = attach.detach().to(torch.float32).contiguous()
    lengths = torch.as_tensor(lengths).to(device=dev, dtype=torch.int64).contiguous()
    if tuple(lengths.shape) != (B,):
        raise VlgaeError(f"lengths must be [B] = {(B,)}, got {tuple(lengths.shape)}")
    return dev, B, N, dec, attach, lengths


def dmv_inside_outside(dec, attach, lengths, *, want_grad=True, gZ=None, mask_zero=MASK_ZERO):
    """Z [B] and (optionally) d(sum gZ*Z)/d dec, d(...)/d attach.  See vlgae_dmv_inside_outside."""
    dev, B, N, dec, attach, lengths = _prep(dec, attach, lengths)
    Z = torch.empty(B, dtype=torch.float32, device=dev)
    gdec = torch.empty((B, N, 2, 2, 2), dtype=torch.float32, device=dev) if want_grad else None
    gatt = torch.empty((B, N, N, 2), dtype=torch.float32, device=dev) if want_grad else None
    if gZ is not None:
        gZ = gZ.detach().to(device=dev, dtype=torch.float32).reshape(B).contiguous()
    ws, ws_bytes = _workspace(dev, B, N)
    with torch.cuda.device(dev):
        check(lib().vlgae_dmv_inside_outside(dec.data_ptr(), attach.data_ptr(), lengths.data_ptr(), B, N, mask_zero,
                                             _ptr(gZ), Z.data_ptr(), _ptr(gdec), _ptr(gatt), _ptr(ws), ws_bytes,
                                             _stream(dev)), "vlgae_dmv_inside_outside")
    return Z, gdec, gatt


def dmv_viterbi(dec, attach, lengths, *, want_heads=True, want_arcs=True, want_gdec=False, mask_zero=MASK_ZERO):
    """best [B], heads [B,N] int64, arcs [B,N,N,2], gdec [B,N,2,2,2].  See vlgae_dmv_viterbi."""
    dev, B, N, dec, attach, lengths = _prep(dec, attach, lengths)
    best = torch.empty(B, dtype=torch.float32, device=dev)
    heads = torch.empty((B, N), dtype=torch.int64, device=dev) if want_heads else None
    arcs = torch.empty((B, N, N, 2), dtype=torch.float32, device=dev) if want_arcs else None
    gdec = torch.empty((B, N, 2, 2, 2), dtype=torch.float32, device=dev) if want_gdec else None
    ws, ws_bytes = _workspace(dev, B, N)
    with torch.cuda.device(dev):
        check(lib().vlgae_dmv_viterbi(dec.data_ptr(), attach.data_ptr(), lengths.data_ptr(), B, N, mask_zero,
                                      best.data_ptr(), _ptr(heads), _ptr(arcs), _ptr(gdec), _ptr(ws), ws_bytes,
                                      _stream(dev)), "vlgae_dmv_viterbi")
    return best, heads, arcs, gdec


class ParseBuffers:
    """Pre-allocated outputs for repeated `dmv_parse` calls on same-shaped batches (no allocator traffic)."""

    def __init__(self, B, N, device, want_arcs=False, want_vgdec=False):
        f32 = dict(dtype=torch.float32, device=device)
        self.Z = torch.empty(B, **f32)
        self.gdec = torch.empty((B, N, 2, 2, 2), **f32)
        self.gattach = torch.empty((B, N, N, 2), **f32)
        self.best = torch.empty(B, **f32)
        self.heads = torch.empty((B, N), dtype=torch.int64, device=device)
        self.arcs = torch.empty((B, N, N, 2), **f32) if want_arcs else None
        self.vgdec = torch.empty((B, N, 2, 2, 2), **f32) if want_vgdec else None


def dmv_parse(dec, attach, lengths, *, out: Optional[ParseBuffers] = None, want_arcs=False, want_vgdec=False,
              gZ=None, mask_zero=MASK_ZERO, prepared=False):
    """inside + outside + Viterbi in one launch (vlgae_dmv_parse).  Returns a ParseBuffers."""
    if prepared:
        dev, (B, N) = dec.device, dec.shape[:2]
    else:
        dev, B, N, dec, attach, lengths = _prep(dec, attach, lengths)
    if out is None:
        out = ParseBuffers(B, N, dev, want_arcs, want_vgdec)
    ws, ws_bytes = _workspace(dev, B, N)
    with torch.cuda.device(dev):
        check(lib().vlgae_dmv_parse(dec.data_ptr(), attach.data_ptr(), lengths.data_ptr(), B, N, mask_zero, _ptr(gZ),
                                    out.Z.data_ptr(), out.gdec.data_ptr(), out.gattach.data_ptr(),
                                    out.best.data_ptr(), out.heads.data_ptr(), _ptr(out.arcs), _ptr(out.vgdec),
                                    _ptr(ws), ws_bytes, _stream(dev)), "vlgae_dmv_parse")
    return out


def dmv_merge(dec, attach, root, one=0.0, zero=-1e12):
    """DMV1o.merge forward (always float32, reference quirk Q3)."""
    dev = _require_cuda(dec, attach, root)
    B, n = dec.shape[:2]
    if tuple(dec.shape) != (B, n, 2, 2, 2) or tuple(attach.shape) != (B, n, n, 2) or tuple(root.shape) != (B, n):
        raise VlgaeError("merge: dec [B,n,2,2,2], attach [B,n,n,2], root [B,n] expected")
    dec = dec.detach().to(torch.float32).contiguous()
    attach = attach.detach().to(torch.float32).contiguous()
    root = root.detach().to(torch.float32).contiguous()
    dec_w = torch.empty((B, n + 1, 2, 2, 2), dtype=torch.float32, device=dev)
    att_w = torch.empty((B, n + 1, n + 1, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().vlgae_dmv_merge(dec.data_ptr(), attach.data_ptr(), root.data_ptr(), B, n, float(one), float(zero),
                                    dec_w.data_ptr(), att_w.data_ptr(), _stream(dev)), "vlgae_dmv_merge")
    return dec_w, att_w


def scale_rows(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """out[b, ...] = g[b] * x[b, ...]."""
    dev = _require_cuda(x, g)
    B = x.shape[0]
    x = x.contiguous()
    g = g.detach().to(torch.float32).reshape(B).contiguous()
    out = torch.empty_like(x)
    inner = x.numel() // max(B, 1)
    with torch.cuda.device(dev):
        check(lib().vlgae_scale_rows(x.data_ptr(), g.data_ptr(), B, inner, out.data_ptr(), _stream(dev)),
              "vlgae_scale_rows")
    return out
